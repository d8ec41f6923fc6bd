// Input / output steps either side of the generator, HBM-bound, one pass each:
//   depth_masks   getDepthMask (codes/data/LQGTker_Depth_dataset.py:204-226): per-image depth range -> K bins ->
//                 u8 label map (what K-DYN and K-LOSS consume) and, optionally, the reference's fp32 one-hot planes
//   tensor2img    clamp / x255 / round-half-even / uint8 / RGB->BGR / CHW->HWC (codes/utils/util.py:566-590)
#include "dasr_internal.h"

namespace dasr {

// one block per image: min/max of the depth map, then the reference's fp32 bin edges
//   interval = (max - min) / K ; start_i = min + interval * i ; end_i = min + interval * (i + 1)
//   mask_i = (d >= start_i) & (d < end_i)          (fp32, no fused multiply-add: same roundings as torch)
__global__ void __launch_bounds__(256) depth_masks_kernel(const float* __restrict__ depth, uint8_t* __restrict__ labels,
                                                          float* __restrict__ masks, float* __restrict__ range_out,
                                                          int K, int HW, int fixed_range) {
    __shared__ float smin[8], smax[8];
    __shared__ float edges[DASR_LOSS_KMAX + 1];
    const int b = blockIdx.x;
    const float* dp = depth + (size_t)b * HW;
    float mn = INFINITY, mx = -INFINITY;
    if (!fixed_range) {
        for (int i = threadIdx.x; i < HW; i += blockDim.x) {
            const float d = __ldg(dp + i);
            mn = fminf(mn, d);
            mx = fmaxf(mx, d);
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        }
        if ((threadIdx.x & 31) == 0) {
            smin[threadIdx.x >> 5] = mn;
            smax[threadIdx.x >> 5] = mx;
        }
    }
    __syncthreads();
    if (threadIdx.x <= K) {
        const int i = threadIdx.x;
        if (fixed_range) {
            // python scalars: interval = (1 - 0) / K and the edges are evaluated in double, then compared as fp32
            edges[i] = (float)(0.0 + (1.0 / (double)K) * (double)i);
        } else {
            float a = smin[0], z = smax[0];
            for (int w = 1; w < 8; w++) {
                a = fminf(a, smin[w]);
                z = fmaxf(z, smax[w]);
            }
            const float interval = __fdiv_rn(__fsub_rn(z, a), (float)K);
            edges[i] = __fadd_rn(a, __fmul_rn(interval, (float)i));
            if (range_out && i == 0) {
                range_out[2 * b] = a;
                range_out[2 * b + 1] = z;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float d = __ldg(dp + i);
        int lab = 255;
        for (int k = 0; k < K; k++)
            if (d >= edges[k] && d < edges[k + 1]) lab = k;      // bins are disjoint: at most one hit
        labels[(size_t)b * HW + i] = (uint8_t)lab;
        if (masks)
            for (int k = 0; k < K; k++) masks[((size_t)b * K + k) * HW + i] = (k == lab) ? 1.f : 0.f;
    }
}

// sr NCHW fp32 [B,3,H,W] (RGB) -> img u8 [B,H,W,3] (BGR): round_half_even(clamp(x, lo, hi) - lo) / (hi - lo) * 255)
__global__ void tensor2img_kernel(const float* __restrict__ sr, uint8_t* __restrict__ img, int B, int HW, float lo,
                                  float hi) {
    const size_t total = (size_t)B * HW;
    const float inv = hi - lo;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / HW, p = i - b * HW;
        const float* s = sr + b * 3 * (size_t)HW + p;
        uint8_t o[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float v = fminf(fmaxf(__ldg(s + (size_t)c * HW), lo), hi);
            v = __fdiv_rn(__fsub_rn(v, lo), inv);
            o[2 - c] = (uint8_t)__float2int_rn(__fmul_rn(v, 255.0f));     // BGR order, numpy round = half to even
        }
        uint8_t* d = img + i * 3;
        d[0] = o[0];
        d[1] = o[1];
        d[2] = o[2];
    }
}

// F.interpolate(x, size=(Ho, Wo), mode='nearest') for integer ratios (normalization.py:58-59: a SEAN instance above LR
// resolution resizes the depth map and the depth masks to its feature map): out[b,c,Y,X] = in[b,c,Y/ry,X/rx]
__global__ void __launch_bounds__(256) nearest_up_kernel(const float* __restrict__ in, float* __restrict__ out, int H,
                                                         int W, int ry, int rx, size_t total) {
    const int Wo = W * rx, Ho = H * ry;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(i % Wo);
        const size_t r = i / Wo;
        const int Y = (int)(r % Ho);
        const size_t plane = r / Ho;
        out[i] = __ldg(in + (plane * H + Y / ry) * W + X / rx);
    }
}

}  // namespace dasr

using namespace dasr;

extern "C" int dasr_depth_masks(const float* depth, uint8_t* labels, float* masks, float* range_out, int B, int K, int H,
                                int W, int fixed_range, void* stream) {
    DASR_REQUIRE(depth && labels && B > 0 && H > 0 && W > 0, "bad arguments");
    DASR_REQUIRE(K >= 1 && K <= DASR_LOSS_KMAX, "depth masks: 1..%d bins (got %d)", DASR_LOSS_KMAX, K);
    depth_masks_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(depth, labels, masks, range_out, K, H * W, fixed_range);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_nearest_up(const float* in, float* out, int planes, int H, int W, int Ho, int Wo, void* stream) {
    DASR_REQUIRE(in && out && planes > 0 && H > 0 && W > 0, "bad arguments");
    DASR_REQUIRE(Ho >= H && Wo >= W && Ho % H == 0 && Wo % W == 0, "nearest resize: integer ratios only (%dx%d -> %dx%d)", H, W,
                 Ho, Wo);
    const size_t total = (size_t)planes * Ho * Wo;
    size_t grid = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    if (grid > cap) grid = cap;
    nearest_up_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(in, out, H, W, Ho / H, Wo / W, total);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_tensor2img(const float* sr, uint8_t* img, int B, int H, int W, float lo, float hi, void* stream) {
    DASR_REQUIRE(sr && img && B > 0 && H > 0 && W > 0 && hi > lo, "bad arguments");
    const size_t total = (size_t)B * H * W;
    size_t grid = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    if (grid > cap) grid = cap;
    tensor2img_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(sr, img, B, H * W, lo, hi);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

// ------------------------------------------------------------------------------------------------ validation metrics
// PSNR (codes/utils/util.py:646-653 as called from codes/train.py:251-257) and
// pytorch_ssim.ssim (codes/pytorch_ssim/__init__.py:17-38,65-72) of the validation loop, on the device.
namespace dasr {

// exact integer sum of squared differences of two uint8 frames [H,W,C] inside a `crop`-pixel border, per frame
__global__ void __launch_bounds__(256) sqdiff_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                        unsigned long long* __restrict__ out, int H, int W, int C,
                                                        int crop) {
    const int f = blockIdx.y;
    const int Hc = H - 2 * crop, Wc = W - 2 * crop;
    const size_t n = (size_t)Hc * Wc * C;
    const uint8_t* pa = a + (size_t)f * H * W * C;
    const uint8_t* pb = b + (size_t)f * H * W * C;
    unsigned long long s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t p = i / C;
        const int x = (int)(p % Wc) + crop, y = (int)(p / Wc) + crop;
        const size_t idx = ((size_t)y * W + x) * C + c;
        const int d = (int)pa[idx] - (int)pb[idx];
        s += (unsigned long long)(d * d);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out + f, s);      // integer atomics: exact and order-independent
}

// psnr[f] = 20 log10(255 / sqrt(ssd[f] / n))  (double; +inf for identical frames, like util.calculate_psnr)
__global__ void psnr_finalize_kernel(const unsigned long long* __restrict__ ssd, double* __restrict__ out, int F, double n) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double mse = (double)ssd[f] / n;
    out[f] = ssd[f] == 0ull ? (double)INFINITY : 20.0 * log10(255.0 / sqrt(mse));
}

// SSIM map of one 16x16 output tile per block: separable 11-tap Gaussian (sigma 1.5, zero padding) of x, y, x^2, y^2,
// xy in shared memory, then the SSIM formula and a block sum.  part[frame][channel][tile] partial sums.
constexpr int kSsimT = 16, kSsimR = 5, kSsimP = kSsimT + 2 * kSsimR;    // tile, radius, padded tile (26)
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                                                   float* __restrict__ part, int H, int W, int tiles_x, int tiles_y) {
    __shared__ float s1[kSsimP][kSsimP + 1], s2[kSsimP][kSsimP + 1];
    __shared__ float hz[5][kSsimP][kSsimT + 1];       // horizontally filtered rows of the 5 quantities
    __shared__ float g[11];
    __shared__ float red[8];
    const int plane = blockIdx.y;                      // frame * C + channel
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const float* p1 = img1 + (size_t)plane * H * W;
    const float* p2 = img2 + (size_t)plane * H * W;
    if (threadIdx.x < 11) {
        // gaussian(11, 1.5) normalised like the reference (fp32 exp, fp32 sum)
        float w[11], sum = 0.f;
        for (int i = 0; i < 11; i++) {
            w[i] = expf(-(float)((i - 5) * (i - 5)) / (2.f * 1.5f * 1.5f));
            sum += w[i];
        }
        g[threadIdx.x] = w[threadIdx.x] / sum;
    }
    const int y0 = ty * kSsimT - kSsimR, x0 = tx * kSsimT - kSsimR;
    for (int i = threadIdx.x; i < kSsimP * kSsimP; i += 256) {
        const int r = i / kSsimP, c = i - r * kSsimP;
        const int y = y0 + r, x = x0 + c;
        const bool in = (y >= 0 && y < H && x >= 0 && x < W);
        s1[r][c] = in ? __ldg(p1 + (size_t)y * W + x) : 0.f;
        s2[r][c] = in ? __ldg(p2 + (size_t)y * W + x) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSsimP * kSsimT; i += 256) {
        const int r = i / kSsimT, c = i - r * kSsimT;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float u = s1[r][c + k], v = s2[r][c + k], wk = g[k];
            a = fmaf(wk, u, a);
            b = fmaf(wk, v, b);
            aa = fmaf(wk, u * u, aa);
            bb = fmaf(wk, v * v, bb);
            ab = fmaf(wk, u * v, ab);
        }
        hz[0][r][c] = a; hz[1][r][c] = b; hz[2][r][c] = aa; hz[3][r][c] = bb; hz[4][r][c] = ab;
    }
    __syncthreads();
    float acc = 0.f;
    {
        const int r = threadIdx.x / kSsimT, c = threadIdx.x - r * kSsimT;       // 256 threads = 16 x 16 outputs
        const int y = ty * kSsimT + r, x = tx * kSsimT + c;
        if (y < H && x < W) {
            float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 11; k++)
#pragma unroll
                for (int q = 0; q < 5; q++) m[q] = fmaf(g[k], hz[q][r + k][c], m[q]);
            const float mu1 = m[0], mu2 = m[1];
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float sg1 = m[2] - mu1_sq, sg2 = m[3] - mu2_sq, sg12 = m[4] - mu12;
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
            acc = ((2.f * mu12 + C1) * (2.f * sg12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sg1 + sg2 + C2));
        }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; i++) t += red[i];
        part[(size_t)plane * tiles_x * tiles_y + blockIdx.x] = t;
    }
}

// out[frame] = mean over the frame's channels and pixels of the SSIM map (tiles added in index order, in double)
__global__ void ssim_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int per_frame, double inv_n) {
    const int f = blockIdx.x;
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < per_frame; i += blockDim.x) s += (double)part[(size_t)f * per_frame + i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[f] = (float)(red[0] * inv_n);
}

}  // namespace dasr

extern "C" int dasr_sqdiff_u8(const uint8_t* a, const uint8_t* b, unsigned long long* out, int F, int H, int W, int C,
                              int crop, void* stream) {
    DASR_REQUIRE(a && b && out && F > 0 && crop >= 0 && H > 2 * crop && W > 2 * crop, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    DASR_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)F * sizeof(unsigned long long), st));
    const size_t n = (size_t)(H - 2 * crop) * (W - 2 * crop) * C;
    size_t gx = (n + 255) / 256;
    if (gx > (size_t)num_sms() * 8) gx = (size_t)num_sms() * 8;
    sqdiff_u8_kernel<<<dim3((unsigned)gx, F), 256, 0, st>>>(a, b, out, H, W, C, crop);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_psnr_u8(const uint8_t* a, const uint8_t* b, unsigned long long* ssd, double* psnr, int F, int H, int W,
                            int C, int crop, void* stream) {
    DASR_REQUIRE(psnr != nullptr, "null pointer");
    int rc = dasr_sqdiff_u8(a, b, ssd, F, H, W, C, crop, stream);
    if (rc) return rc;
    psnr_finalize_kernel<<<(F + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ssd, psnr, F,
                                                                           (double)(H - 2 * crop) * (W - 2 * crop) * C);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_ssim_tiles(int H, int W) { return ((H + kSsimT - 1) / kSsimT) * ((W + kSsimT - 1) / kSsimT); }

extern "C" int dasr_ssim(const float* img1, const float* img2, float* part, float* out, int F, int C, int H, int W,
                         void* stream) {
    DASR_REQUIRE(img1 && img2 && part && out && F > 0 && C > 0, "bad arguments");
    const int tx = (W + kSsimT - 1) / kSsimT, ty = (H + kSsimT - 1) / kSsimT;
    cudaStream_t st = (cudaStream_t)stream;
    ssim_kernel<<<dim3(tx * ty, F * C), 256, 0, st>>>(img1, img2, part, H, W, tx, ty);
    DASR_LAUNCH_OK();
    ssim_reduce_kernel<<<F, 256, 0, st>>>(part, out, C * tx * ty, 1.0 / ((double)C * H * W));
    DASR_LAUNCH_OK();
    return DASR_OK;
}

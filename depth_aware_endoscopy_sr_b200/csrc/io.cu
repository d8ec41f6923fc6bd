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

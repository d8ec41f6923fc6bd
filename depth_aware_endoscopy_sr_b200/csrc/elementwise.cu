// Memory-bound kernels of the DepthNet hot path (CUDA cores, vectorised, coalesced):
//   conv_first      encoder.layer1 (Cin=3) + LeakyReLU, NCHW fp32 -> NHWC bf16   (sftmd_arch.py:743,772,783)
//   zero_insert2    zero-stuffed input of the transposed conv                    (sftmd_arch.py:748)
//   add             feat_add1                                                    (sftmd_arch.py:931)
//   region_pool     RegionWiseAvgPooling                                         (sftmd_arch.py:714-733)
//   mask_labels     one-hot depth masks -> u8 label map (getDepthMask output, LQGTker_Depth_dataset.py:204-226)
//   actv            SEAN mlp_mask: ReLU(conv3x3(depth, 1->C))                    (normalization.py:37-40,61)
//   style_mix       SEAN A_i_j label mixing st' = A st + a                       (normalization.py:27,80)
//   dynconv         K-DYN: depth-guided dynamic 3x3 convolution, apply step      (normalization.py:81-85 restated)
//   instats_finalize double InstanceNorm closed form                             (sftmd_arch.py:813 + normalization.py:56)
#include "dasr_internal.h"
#include <stdlib.h>
#include <string.h>

namespace dasr {

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }

__device__ __forceinline__ uint4 pack8f(const float* f) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}
__device__ __forceinline__ void unpack8f(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

// ------------------------------------------------------------------------------------ conv_first
__global__ void __launch_bounds__(128) conv_first_kernel(const float* __restrict__ x, const float* __restrict__ v,
                                                         const float* __restrict__ g, const float* __restrict__ bias,
                                                         __nv_bfloat16* __restrict__ out, int B, int H, int W, int npl) {
    __shared__ float ws[27][32];  // [ci*9 + tap][co]
    __shared__ float bs[32];
    if (threadIdx.x < 32) {
        const int co = threadIdx.x;
        float ss = 0.f;
        for (int i = 0; i < 27; i++) ss += v[co * 27 + i] * v[co * 27 + i];
        const float sc = g ? g[co] / sqrtf(ss) : 1.f;
        for (int i = 0; i < 27; i++) ws[i][co] = v[co * 27 + i] * sc;
        bs[co] = bias[co];
    }
    __syncthreads();
    const size_t npix = (size_t)B * H * W;
    for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (size_t)gridDim.x * blockDim.x) {
        const int w = pix % W;
        const int h = (pix / W) % H;
        const int b = pix / ((size_t)W * H);
        float in[27];
#pragma unroll
        for (int ci = 0; ci < 3; ci++)
#pragma unroll
            for (int t = 0; t < 3; t++)
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const int hh = h + t - 1, ww = w + u - 1;
                    in[ci * 9 + t * 3 + u] = (hh >= 0 && hh < H && ww >= 0 && ww < W)
                                                 ? __ldg(x + (((size_t)b * 3 + ci) * H + hh) * W + ww)
                                                 : 0.f;
                }
        uint4* op = reinterpret_cast<uint4*>(out + pix * 32);
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 8) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = bs[c0 + j];
#pragma unroll
            for (int i = 0; i < 27; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] = fmaf(in[i], ws[i][c0 + j], acc[j]);
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = lrelu02(acc[j]);
            if (npl == 1) op[c0 / 8] = pack8f(acc);
            else pl_store8(reinterpret_cast<uint4*>(out), pix * 4 + c0 / 8, npix * 4, npl, acc);
        }
    }
}

// ------------------------------------------------------------------------------------ zero_insert2 / add
__global__ void zero_insert2_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B, int H, int W, int C8) {
    const int Ho = 2 * H - 1, Wo = 2 * W - 1;
    const size_t total = (size_t)B * Ho * Wo * C8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = i % C8;
        size_t pix = i / C8;
        const int wo = pix % Wo;
        const int ho = (pix / Wo) % Ho;
        const int b = pix / ((size_t)Wo * Ho);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (!(ho & 1) && !(wo & 1)) v = __ldg(x + (((size_t)b * H + (ho >> 1)) * W + (wo >> 1)) * C8 + c);
        out[i] = v;
    }
}

__global__ void add_kernel(const uint4* __restrict__ a, const float4* __restrict__ a32, const uint4* __restrict__ b,
                           uint4* __restrict__ out, size_t n8, int npl) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        float fa[8], fb[8];
        if (npl > 1) {       // fp32-split planes: plane sums in, split out
            pl_load8(a, i, n8, npl, fa);
            pl_load8(b, i, n8, npl, fb);
#pragma unroll
            for (int j = 0; j < 8; j++) fa[j] += fb[j];
            pl_store8(out, i, n8, npl, fa);
            continue;
        }
        if (a32) {
            const float4 lo = __ldg(a32 + 2 * i), hi = __ldg(a32 + 2 * i + 1);
            fa[0] = lo.x; fa[1] = lo.y; fa[2] = lo.z; fa[3] = lo.w;
            fa[4] = hi.x; fa[5] = hi.y; fa[6] = hi.z; fa[7] = hi.w;
        } else {
            unpack8f(__ldg(a + i), fa);
        }
        unpack8f(__ldg(b + i), fb);
#pragma unroll
        for (int j = 0; j < 8; j++) fa[j] += fb[j];
        out[i] = pack8f(fa);
    }
}

// ------------------------------------------------------------------------------------ region pooling
// one block per (b, k): threshold the bilinearly (align_corners=True) resized mask, then masked mean.
constexpr int kPoolMaxPos = 8192;
__global__ void __launch_bounds__(256) region_pool_kernel(const __nv_bfloat16* __restrict__ e5,
                                                          const float* __restrict__ masks, float* __restrict__ vec,
                                                          float* __restrict__ msel_out, float* __restrict__ cnt_out,
                                                          int hf, int wf, int C, int K, int H, int W, int npl,
                                                          size_t ps) {
    __shared__ float msel[kPoolMaxPos];
    __shared__ float cnt_s;
    const int b = blockIdx.x / K, k = blockIdx.x % K;
    const float* mp = masks + ((size_t)b * K + k) * H * W;
    const int P = hf * wf;
    const bool resize = (H != hf) || (W != wf);
    const float sh = (hf > 1) ? (float)(H - 1) / (float)(hf - 1) : 0.f;
    const float sw = (wf > 1) ? (float)(W - 1) / (float)(wf - 1) : 0.f;
    float local = 0.f;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        const int i = p / wf, j = p - i * wf;
        float m;
        if (!resize) {
            m = mp[p];
        } else {
            // torch upsample_bilinear2d, align_corners=True: src = scale*dst, 2-tap lerp per axis
            const float hr = __fmul_rn(sh, (float)i), wr = __fmul_rn(sw, (float)j);
            const int h1 = (int)hr, w1 = (int)wr;
            const int h1p = (h1 < H - 1) ? 1 : 0, w1p = (w1 < W - 1) ? 1 : 0;
            const float h1l = __fsub_rn(hr, (float)h1), h0l = __fsub_rn(1.f, h1l);
            const float w1l = __fsub_rn(wr, (float)w1), w0l = __fsub_rn(1.f, w1l);
            const float v00 = mp[h1 * W + w1], v01 = mp[h1 * W + w1 + w1p];
            const float v10 = mp[(h1 + h1p) * W + w1], v11 = mp[(h1 + h1p) * W + w1 + w1p];
            const float t0 = __fadd_rn(__fmul_rn(w0l, v00), __fmul_rn(w1l, v01));
            const float t1 = __fadd_rn(__fmul_rn(w0l, v10), __fmul_rn(w1l, v11));
            const float val = __fadd_rn(__fmul_rn(h0l, t0), __fmul_rn(h1l, t1));
            m = (val >= 0.5f) ? 1.f : 0.f;
        }
        msel[p] = m;
        if (msel_out) msel_out[(size_t)blockIdx.x * P + p] = m;   // saved for the backward pass
        local += m;
    }
    __shared__ float red[8];
    for (int off = 16; off; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[i];
        cnt_s = t;
        if (cnt_out) cnt_out[blockIdx.x] = t;
    }
    __syncthreads();
    const float inv = 1.f / (cnt_s + 1e-10f);
    const __nv_bfloat16* ep = e5 + (size_t)b * P * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < P; p++) {
            const float m = msel[p];
            if (m != 0.f) s = fmaf(m, pl_load1(ep, (size_t)p * C + c, ps, npl), s);
        }
        vec[((size_t)b * K + k) * C + c] = s * inv;
    }
}

// ------------------------------------------------------------------------------------ mask -> labels
__global__ void mask_labels_kernel(const float* __restrict__ masks, uint8_t* __restrict__ labels,
                                   int* __restrict__ flag, int B, int K, int HW) {
    const size_t total = (size_t)B * HW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = i / HW, p = i - (size_t)b * HW;
        int lab = 255, nz = 0;
        bool exact = true;
        for (int k = 0; k < K; k++) {
            const float m = __ldg(masks + ((size_t)b * K + k) * HW + p);
            if (m != 0.f) {
                nz++;
                lab = k;
                if (m != 1.f) exact = false;
            }
        }
        if (nz > 1 || !exact) {
            *flag = 1;
            lab = 255;
        }
        labels[i] = (uint8_t)lab;
    }
}

// ------------------------------------------------------------------------------------ auxiliary input tensor
// aux NHWC bf16 [B,H,W,32]: ch 0..K-1 one-hot depth mask (from the label map), ch 16/17 depth split into two bf16
// parts (hi + lo), ch 18 = 1, others 0.  It is the X operand of the tensor-core weight-gradient kernels that
// differentiate the two input-driven convolutions of SEAN (K-DYN and mlp_mask); built once per forward.
__global__ void build_aux_kernel(const uint8_t* __restrict__ labels, const float* __restrict__ depth,
                                 uint4* __restrict__ aux, size_t npix) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const int lab = labels[i];
        const float d = __ldg(depth + i);
        const __nv_bfloat16 hi = __float2bfloat16(d);
        const __nv_bfloat16 lo = __float2bfloat16(d - __bfloat162float(hi));
        const __nv_bfloat16 lo2 = __float2bfloat16((d - __bfloat162float(hi)) - __bfloat162float(lo));
        __align__(16) __nv_bfloat16 v[32];
#pragma unroll
        for (int c = 0; c < 32; c++) v[c] = __float2bfloat16(c == lab ? 1.f : 0.f);
        v[DASR_AUX_DEPTH_HI] = hi;
        v[DASR_AUX_DEPTH_LO] = lo;
        v[DASR_AUX_DEPTH_LO2] = lo2;
        v[DASR_AUX_ONE] = __float2bfloat16(1.f);
        const uint4* src = reinterpret_cast<const uint4*>(v);
#pragma unroll
        for (int j = 0; j < 4; j++) aux[i * 4 + j] = src[j];
    }
}

// ------------------------------------------------------------------------------------ actv
// Store-bandwidth kernel (2*C bytes per pixel out, 4 bytes in).  Persistent blocks over (image, 2-row band) items; the
// zero-padded depth halo of the band sits in shared memory.  A pixel is produced by C/4 consecutive lanes (a full
// warp at C = 128): lane = 4 channels, whose 36 weights + 4 biases live in registers as float2 pairs, every tap is
// two packed FFMA2 (fma.rn.f32x2), and the lanes of a pixel write one contiguous 2*C-byte row.  A slot walks a run
// of pixels of one row two at a time with a sliding 3x4 window (6 broadcast shared-memory loads per pixel pair, 8
// independent FMA chains).  v1: 242 instructions per (pixel, 8 channels), 64 us; v2 (8 channels per lane, 125
// registers, 16 warps/SM): 33 us; v3 (4 channels per lane, 2 pixels in flight): see profiles/.
constexpr int kActvRows = 2;
// MINB = resident blocks per SM the register budget is compiled for: 3 (<= 85 registers) when the kernel has the
// device to itself; 4 (64 registers) for the "background" launches of Engine._ActvPrefetch, which run ONE block per
// SM next to a convolution kernel that leaves 16 K registers free (conv_igemm.cu, LEAN)
template <int C, int MINB, bool PL>
__global__ void __launch_bounds__(256, MINB) actv_kernel(const float* __restrict__ depth, const float* __restrict__ w,
                                                   const float* __restrict__ bias, uint2* __restrict__ out, int H,
                                                   int W, int n_items, int npl, size_t ps2) {
    constexpr int LPP = C / 4;                     // lanes per pixel
    constexpr int SLOTS = 256 / LPP;
    // lets the SEAN conv that follows in the stream (launched with programmatic stream serialization) set up its
    // barriers / TMEM while this grid drains; it still waits for this grid's completion before reading actv
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ float dsm[];                 // (rows + 2) x (W + 4), zero padded
    const int bands = (H + kActvRows - 1) / kActvRows;
    const int LW = W + 4;
    const int g = threadIdx.x % LPP, slot = threadIdx.x / LPP;
    float2 wr[9][2], br[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        br[j] = make_float2(__ldg(bias + g * 4 + 2 * j), __ldg(bias + g * 4 + 2 * j + 1));
#pragma unroll
        for (int t = 0; t < 9; t++)
            wr[t][j] = make_float2(__ldg(w + (g * 4 + 2 * j) * 9 + t), __ldg(w + (g * 4 + 2 * j + 1) * 9 + t));
    }
    const int seg = 16;
    const int segs_per_row = (W + seg - 1) / seg;
    // persistent over (image, band) items: the weights (identical for every image) are loaded once per block
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / bands, band = item - b * bands;
        const int h0 = band * kActvRows;
        const int rows = min(kActvRows, H - h0);
        const float* dp = depth + (size_t)b * H * W;
        __syncthreads();
        for (int i = threadIdx.x; i < (rows + 2) * LW; i += 256) {
            const int r = i / LW, c = i - r * LW;
            const int hh = h0 + r - 1, ww = c - 1;
            dsm[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(dp + hh * W + ww) : 0.f;
        }
        __syncthreads();
        const int nruns = rows * segs_per_row;
        uint2* op = out + ((size_t)b * H + h0) * W * LPP;
        for (int run = slot; run < nruns; run += SLOTS) {
            const int r = run / segs_per_row;
            const int x0 = (run - r * segs_per_row) * seg;
            const int x1 = min(x0 + seg, W);
            const float* d0 = dsm + r * LW + x0;       // window row 0, image column x0 - 1 (padded coordinates)
            float win[3][4];
#pragma unroll
            for (int t = 0; t < 3; t++) {
                win[t][2] = d0[t * LW];
                win[t][3] = d0[t * LW + 1];
            }
            for (int x = x0; x < x1; x += 2) {
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    win[t][0] = win[t][2];
                    win[t][1] = win[t][3];
                    win[t][2] = d0[t * LW + (x - x0) + 2];
                    win[t][3] = d0[t * LW + (x - x0) + 3];
                }
                float2 acc[2][2];
#pragma unroll
                for (int q = 0; q < 2; q++)
#pragma unroll
                    for (int j = 0; j < 2; j++) acc[q][j] = br[j];
#pragma unroll
                for (int t = 0; t < 3; t++)
#pragma unroll
                    for (int u = 0; u < 3; u++)
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const float2 d2 = make_float2(win[t][u + q], win[t][u + q]);
#pragma unroll
                            for (int j = 0; j < 2; j++) acc[q][j] = __ffma2_rn(d2, wr[t * 3 + u][j], acc[q][j]);
                        }
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    if (x + q >= x1) break;
                    uint2 o;
                    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                    for (int j = 0; j < 2; j++)
                        oh[j] = __floats2bfloat162_rn(fmaxf(acc[q][j].x, 0.f), fmaxf(acc[q][j].y, 0.f));
                    op[((size_t)r * W + x + q) * LPP + g] = o;
                    if (PL) {          // fp32-split planes: the remainders of the fp32 value
                        float2 rem[2];
#pragma unroll
                        for (int j = 0; j < 2; j++) rem[j] = make_float2(fmaxf(acc[q][j].x, 0.f), fmaxf(acc[q][j].y, 0.f));
                        for (int k = 1; k < npl; k++) {
#pragma unroll
                            for (int j = 0; j < 2; j++) {
                                const float2 t = __bfloat1622float2(oh[j]);
                                rem[j].x -= t.x;
                                rem[j].y -= t.y;
                                oh[j] = __floats2bfloat162_rn(rem[j].x, rem[j].y);
                            }
                            op[(size_t)k * ps2 + ((size_t)r * W + x + q) * LPP + g] = o;
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------ style mix
// stp[b][j][c] = sum_i A[j][i] st[b][i][c] + a[j]   -> bf16 (GEMM A operand of the table GEMM)
__global__ void style_mix_kernel(const float* __restrict__ st, const float* __restrict__ A, const float* __restrict__ a,
                                 __nv_bfloat16* __restrict__ stp, int B, int K, int L, int npl) {
    const size_t total = (size_t)B * K * L;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = i % L;
        const int j = (i / L) % K;
        const int b = i / ((size_t)L * K);
        float s = a[j];
        for (int q = 0; q < K; q++) s = fmaf(A[j * K + q], st[((size_t)b * K + q) * L + c], s);
        pl_store1(stp, i, total, npl, s);
    }
}

// all SEAN instances of the network in one launch: stp[s][b][j][c] with per-instance A_i_j weights (pointer tables)
__global__ void style_mix_batched_kernel(const float* __restrict__ st, const float* const* __restrict__ A_ptrs,
                                         const float* const* __restrict__ a_ptrs, __nv_bfloat16* __restrict__ stp,
                                         int B, int K, int L, int npl) {
    const int sidx = blockIdx.y;
    const float* A = A_ptrs[sidx];
    const float* a = a_ptrs[sidx];
    const size_t total = (size_t)B * K * L;
    __nv_bfloat16* dst = stp + (size_t)sidx * total;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = i % L;
        const int j = (i / L) % K;
        const int b = i / ((size_t)L * K);
        float s = __ldg(a + j);
        for (int q = 0; q < K; q++) s = fmaf(__ldg(A + j * K + q), st[((size_t)b * K + q) * L + c], s);
        pl_store1(dst, i, total * gridDim.y, npl, s);
    }
}

// ------------------------------------------------------------------------------------ K-DYN apply
// Block = one (image, band of rows); table T[b] ([K][9][C2] bf16, plus one all-zero row for pixels in no mask) and
// the label halo are staged in smem.  thread item = (pixel, 8-channel group): 9 label lookups + 9 16-byte smem
// reads + one 16-byte store; the bf16 table entries are accumulated in fp32 with the mixed-precision add of sm_100
// (add.rn.f32.bf16 -> FHADD.BF16 with a free .H0/.H1 operand select: no unpacking instructions).
__device__ __forceinline__ void acc_bf16x8(float* acc, const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        unsigned short lo, hi;
        asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(w[j]));
        asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(acc[2 * j]) : "h"(lo));
        asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(acc[2 * j + 1]) : "h"(hi));
    }
}

__global__ void __launch_bounds__(256) dynconv_labels_kernel(const __nv_bfloat16* __restrict__ table,
                                                             const uint8_t* __restrict__ labels,
                                                             const float* __restrict__ masks, const int* __restrict__ flag,
                                                             uint4* __restrict__ out, int K, int H, int W, int C2,
                                                             int rows_per_block, int n_items) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint4* ts = reinterpret_cast<uint4*>(smraw);  // (K + 1)*9*C2/8 uint4; row K is all zero
    const int G = C2 / 8;
    const int tsz = K * 9 * G;
    uint8_t* ls = smraw + (size_t)(tsz + 9 * G) * 16;  // (rows+2) x (W+2) labels
    const int bands = (H + rows_per_block - 1) / rows_per_block;
    // persistent over a CONTIGUOUS range of (image, band) items: the table is restaged only when the image changes
    const int item0 = (int)((long long)blockIdx.x * n_items / gridDim.x);
    const int item1 = (int)((long long)(blockIdx.x + 1) * n_items / gridDim.x);
    const bool general = (flag != nullptr) && (*flag != 0) && (masks != nullptr);
    int staged_img = -1;
    for (int item = item0; item < item1; item++) {
    const int b = item / bands, band = item - b * bands;
    const int h0 = band * rows_per_block;
    const int rows = min(rows_per_block, H - h0);
    __syncthreads();
    if (b != staged_img) {
        const uint4* tg = reinterpret_cast<const uint4*>(table + (size_t)b * K * 9 * C2);
        for (int i = threadIdx.x; i < tsz; i += blockDim.x) ts[i] = __ldg(tg + i);
        for (int i = threadIdx.x; i < 9 * G; i += blockDim.x) ts[tsz + i] = make_uint4(0, 0, 0, 0);
        staged_img = b;
    }
    const int LW = W + 2;
    for (int i = threadIdx.x; i < (rows + 2) * LW; i += blockDim.x) {
        const int r = i / LW, c = i - r * LW;
        const int hh = h0 + r - 1, ww = c - 1;
        int lab = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? labels[((size_t)b * H + hh) * W + ww] : K;
        ls[i] = (uint8_t)(lab < K ? lab : K);      // 255 (pixel in no mask) and the zero padding -> the zero row
    }
    __syncthreads();
    const int items = rows * W * G;
    if (general) {
        // masks are not one-hot: exact linear form  sum_{k,tap} mask * T  (slow path, same smem table)
        for (int it = threadIdx.x; it < items; it += blockDim.x) {
            const int gch = it % G;
            const int pl = it / G;
            const int r = pl / W, c = pl - r * W;
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = 0.f;
            for (int k = 0; k < K; k++)
                for (int t = 0; t < 3; t++)
                    for (int u = 0; u < 3; u++) {
                        const int yy = h0 + r + t - 1, xx = c + u - 1;
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                        const float m = __ldg(masks + (((size_t)b * K + k) * H + yy) * W + xx);
                        if (m != 0.f) {
                            float f[8];
                            unpack8f(ts[(k * 9 + t * 3 + u) * G + gch], f);
#pragma unroll
                            for (int j = 0; j < 8; j++) acc[j] = fmaf(m, f[j], acc[j]);
                        }
                    }
            out[(((size_t)b * H + h0 + r) * W + c) * G + gch] = pack8f(acc);
        }
        continue;
    }
    // one-hot fast path.  A thread keeps its channel group and walks pixels with a fixed stride (no divisions in
    // the loop); the tap offsets into the table are compile-time multiples of G.
    if (G == 16) {
        const int gch = threadIdx.x & 15;
        const uint4* tp = ts + gch;
        const int npix = rows * W;
        int pl = threadIdx.x >> 4;
        int r = pl / W, c = pl - r * W;
        const int step = blockDim.x >> 4;          // pixels advanced per iteration (< W is not required)
        uint4* op = out + ((size_t)b * H + h0) * W * 16 + gch;
        for (; pl < npix; pl += step) {
            const uint8_t* lp = ls + r * LW + c;
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
            for (int t = 0; t < 3; t++)
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const int lab = lp[t * LW + u];
                    acc_bf16x8(acc, tp[lab * (9 * 16) + (t * 3 + u) * 16]);
                }
            op[(size_t)pl * 16] = pack8f(acc);
            c += step;
            while (c >= W) { c -= W; r++; }
        }
        continue;
    }
    const int G9 = 9 * G;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int gch = it % G;
        const int pl = it / G;
        const int r = pl / W, c = pl - r * W;
        const uint8_t* lp = ls + r * LW + c;
        const uint4* tp = ts + gch;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
        for (int t = 0; t < 3; t++)
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int lab = lp[t * LW + u];
                acc_bf16x8(acc, tp[lab * G9 + (t * 3 + u) * G]);
            }
        out[(((size_t)b * H + h0 + r) * W + c) * G + gch] = pack8f(acc);
    }
    }
}

// masks NCHW fp32 [B,K,H,W] -> NHWC bf16 [B,H,W,16]: the A operand of the K-DYN extension of the SEAN GEMM
__global__ void build_mask16_kernel(const float* __restrict__ masks, uint4* __restrict__ out, int K, int HW, size_t npix) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / HW, p = i - b * HW;
        __align__(16) __nv_bfloat16 v[16];
#pragma unroll
        for (int k = 0; k < 16; k++)
            v[k] = __float2bfloat16(k < K ? __ldg(masks + (b * K + k) * (size_t)HW + p) : 0.f);
        const uint4* src = reinterpret_cast<const uint4*>(v);
        out[2 * i] = src[0];
        out[2 * i + 1] = src[1];
    }
}

// table T[n][k][tap][c] (bf16) -> GEMM-B weights of the dynamic convolution: wdyn[n][c][tap*16 + k] (k >= K zero),
// n = (SEAN instance, image).  One block per n.
__global__ void __launch_bounds__(256) table_to_dynweights_kernel(const __nv_bfloat16* __restrict__ table,
                                                                  __nv_bfloat16* __restrict__ wdyn, int K, int C2,
                                                                  int group, int npl) {
    extern __shared__ __align__(16) uint8_t smraw[];
    __nv_bfloat16* ts = reinterpret_cast<__nv_bfloat16*>(smraw);       // [K][9][C2]
    // fp32-split planes: blockIdx.y = plane; table planes are outermost ([npl][n]...), the planes of wdyn sit inside
    // every group of `group` images ([n / group][npl][group]...: one SEAN instance = one weight matrix with planes)
    const size_t n = blockIdx.x, pl = blockIdx.y;
    const int tsz8 = K * 9 * C2 / 8;
    const uint4* tg = reinterpret_cast<const uint4*>(table + (pl * gridDim.x + n) * (size_t)K * 9 * C2);
    for (int i = threadIdx.x; i < tsz8; i += blockDim.x) reinterpret_cast<uint4*>(ts)[i] = __ldg(tg + i);
    __syncthreads();
    const size_t nd = ((n / group) * npl + pl) * group + n % group;
    uint4* dst = reinterpret_cast<uint4*>(wdyn + nd * (size_t)C2 * 9 * 16);
    const int items = C2 * 9 * 2;                                       // 16-byte pieces: (c, tap, 8 k's)
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int kg = it & 1;
        const int tap = (it >> 1) % 9;
        const int c = it / 18;
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = kg * 8 + j;
            v[j] = k < K ? ts[(k * 9 + tap) * C2 + c] : __float2bfloat16(0.f);
        }
        dst[it] = *reinterpret_cast<const uint4*>(v);
    }
}

// general (non one-hot) masks: exact linear form, slow path kept for API completeness
__global__ void dynconv_masks_kernel(const __nv_bfloat16* __restrict__ table, const float* __restrict__ masks,
                                     uint4* __restrict__ out, int B, int K, int H, int W, int C2) {
    const int G = C2 / 8;
    const size_t total = (size_t)B * H * W * G;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int gch = i % G;
        const size_t pix = i / G;
        const int x = pix % W;
        const int y = (pix / W) % H;
        const int b = pix / ((size_t)W * H);
        const uint4* tg = reinterpret_cast<const uint4*>(table + (size_t)b * K * 9 * C2);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = 0.f;
        for (int k = 0; k < K; k++)
            for (int t = 0; t < 3; t++)
                for (int u = 0; u < 3; u++) {
                    const int yy = y + t - 1, xx = x + u - 1;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    const float m = __ldg(masks + (((size_t)b * K + k) * H + yy) * W + xx);
                    if (m != 0.f) {
                        float f[8];
                        unpack8f(__ldg(tg + (k * 9 + t * 3 + u) * G + gch), f);
#pragma unroll
                        for (int j = 0; j < 8; j++) acc[j] = fmaf(m, f[j], acc[j]);
                    }
                }
        out[i] = pack8f(acc);
    }
}

// ------------------------------------------------------------------------------------ IN statistics
__global__ void instats_finalize_kernel(const float* __restrict__ stats, float* __restrict__ norm,
                                        float* __restrict__ normk, int n, int C, int nslots, float inv_hw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = i / C, c = i - b * C;
    const float2* sp = reinterpret_cast<const float2*>(stats) + (size_t)b * nslots * C + c;
    float s1 = 0.f, s2 = 0.f;
    for (int s = 0; s < nslots; s++) {   // fixed order: bit-reproducible
        const float2 v = __ldg(sp + (size_t)s * C);
        s1 += v.x;
        s2 += v.y;
    }
    const float mean = s1 * inv_hw;
    float var = s2 * inv_hw - mean * mean;
    var = fmaxf(var, 0.f);
    const float eps = 1e-5f;
    // IN(IN(y)) = (y - mean) * (var+eps)^-1/2 * (var/(var+eps) + eps)^-1/2
    const float r1 = rsqrtf(var + eps);
    const float r2 = rsqrtf(var * r1 * r1 + eps);
    norm[2 * i] = mean;
    norm[2 * i + 1] = r1 * r2;
    if (normk) {   // backward: k = -2 s'(v) / s = 1/a + eps / (a^2 r), a = v + eps, r = v/a + eps
        const float a = var + eps, r = var / a + eps;
        normk[i] = 1.f / a + eps / (a * a * r);
    }
}

static inline int grid_for(size_t n, int block, int cap = 148 * 16) {
    size_t g = (n + block - 1) / block;
    if (g > (size_t)cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace dasr

using namespace dasr;

extern "C" int dasr_conv_first(const float* x, const float* v, const float* g, const float* bias, void* out, int B,
                               int H, int W, void* stream) {
    DASR_REQUIRE(x && v && bias && out && B > 0 && H > 0 && W > 0, "bad arguments");
    const size_t npix = (size_t)B * H * W;
    conv_first_kernel<<<grid_for(npix, 128), 128, 0, (cudaStream_t)stream>>>(x, v, g, bias, (__nv_bfloat16*)out, B, H, W, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_zero_insert2(const void* x, void* out, int B, int H, int W, int C, void* stream) {
    DASR_REQUIRE(x && out && C % 8 == 0, "bad arguments (C must be a multiple of 8)");
    B *= planes();          // fp32-split planes are extra images of a copy kernel
    const size_t total = (size_t)B * (2 * H - 1) * (2 * W - 1) * (C / 8);
    zero_insert2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)out, B, H, W, C / 8);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_add(const void* a, const float* a32, const void* b, void* out, int64_t n, void* stream) {
    DASR_REQUIRE((a || a32) && b && out && n % 8 == 0, "bad arguments (n must be a multiple of 8)");
    DASR_REQUIRE(planes() == 1 || (a && !a32), "fp32-split planes: no fp32 residual operand");
    add_kernel<<<grid_for((size_t)n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)a, (const float4*)a32, (const uint4*)b, (uint4*)out, (size_t)n / 8, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_region_pool_fwd(const void* e5, const float* masks, float* depth_vec, float* msel, float* cnt,
                                    int B, int hf, int wf, int C, int K, int H, int W, void* stream) {
    DASR_REQUIRE(e5 && masks && depth_vec, "null pointer");
    DASR_REQUIRE(hf * wf <= kPoolMaxPos, "feature map %dx%d too large for region pooling (max %d positions)", hf, wf, kPoolMaxPos);
    region_pool_kernel<<<B * K, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)e5, masks, depth_vec, msel, cnt, hf, wf, C, K, H, W, planes(), (size_t)B * hf * wf * C);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_mask_labels(const float* masks, uint8_t* labels, int32_t* flag, int B, int K, int H, int W,
                                void* stream) {
    DASR_REQUIRE(masks && labels && flag, "null pointer");
    mask_labels_kernel<<<grid_for((size_t)B * H * W, 256), 256, 0, (cudaStream_t)stream>>>(masks, labels, flag, B, K, H * W);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_build_aux(const uint8_t* labels, const float* depth, void* aux, int B, int K, int H, int W,
                              void* stream) {
    DASR_REQUIRE(labels && depth && aux, "null pointer");
    DASR_REQUIRE(K >= 1 && K <= 16, "aux tensor: at most 16 depth masks (got %d)", K);
    const size_t npix = (size_t)B * H * W;
    build_aux_kernel<<<grid_for(npix, 256), 256, 0, (cudaStream_t)stream>>>(labels, depth, (uint4*)aux, npix);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

template <int C, int MINB, bool PL = false>
static int actv_launch(const float* depth, const float* w, const float* bias, void* out, int H, int W, int n_items,
                       size_t smem, int ctas_per_sm, cudaStream_t stream, size_t ps2 = 0) {
    // persistent grid = the blocks that are resident at once (asked from the occupancy calculator), or fewer
    static int per_sm = 0;
    if (per_sm == 0) {
        int nb = 0;
        DASR_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, actv_kernel<C, MINB, PL>, 256, smem));
        per_sm = nb > 0 ? nb : 1;
    }
    const int ctas = (ctas_per_sm > 0 && ctas_per_sm < per_sm) ? ctas_per_sm : per_sm;
    const int cap = ctas * num_sms();
    const int grid = n_items < cap ? n_items : cap;
    actv_kernel<C, MINB, PL><<<grid, 256, smem, stream>>>(depth, w, bias, (uint2*)out, H, W, n_items, planes(), ps2);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_actv_fwd(const float* depth, const float* w, const float* bias, void* out, int B, int H, int W,
                             int C, int ctas_per_sm, void* stream) {
    DASR_REQUIRE(depth && w && bias && out && C % 8 == 0, "bad arguments");
    DASR_REQUIRE(C == 128 || C == 64, "actv: C (= 2*nf) must be 64 or 128 (got %d)", C);
    DASR_REQUIRE((size_t)H * W < 0x7fffffffull, "frame too large");
    DASR_REQUIRE(ctas_per_sm >= 0, "ctas_per_sm must be >= 0 (0 = as many as fit)");
    const int bands = (H + kActvRows - 1) / kActvRows;
    const size_t smem = (size_t)(kActvRows + 2) * (W + 4) * sizeof(float);
    DASR_REQUIRE(smem <= 48 * 1024, "actv: frame too wide (%d)", W);
    const int n_items = B * bands;
    cudaStream_t st = (cudaStream_t)stream;
    if (planes() > 1) {       // fp32-split planes: the same kernel with the plane-split store compiled in
        const size_t ps2 = (size_t)B * H * W * (C / 4);
        return C == 128 ? actv_launch<128, 3, true>(depth, w, bias, out, H, W, n_items, smem, ctas_per_sm, st, ps2)
                        : actv_launch<64, 3, true>(depth, w, bias, out, H, W, n_items, smem, ctas_per_sm, st, ps2);
    }
    if (ctas_per_sm == 1)
        return C == 128 ? actv_launch<128, 4>(depth, w, bias, out, H, W, n_items, smem, 1, st)
                        : actv_launch<64, 4>(depth, w, bias, out, H, W, n_items, smem, 1, st);
    return C == 128 ? actv_launch<128, 3>(depth, w, bias, out, H, W, n_items, smem, ctas_per_sm, st)
                    : actv_launch<64, 3>(depth, w, bias, out, H, W, n_items, smem, ctas_per_sm, st);
}

extern "C" int dasr_style_mix(const float* depth_vec, const float* A, const float* a, void* stp, int B, int K, int L,
                              void* stream) {
    DASR_REQUIRE(depth_vec && A && a && stp, "null pointer");
    style_mix_kernel<<<grid_for((size_t)B * K * L, 256), 256, 0, (cudaStream_t)stream>>>(depth_vec, A, a, (__nv_bfloat16*)stp, B, K, L, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_style_mix_batched(const float* depth_vec, const void* A_ptrs, const void* a_ptrs, void* stp, int nS,
                                      int B, int K, int L, void* stream) {
    DASR_REQUIRE(depth_vec && A_ptrs && a_ptrs && stp && nS > 0, "bad arguments");
    const int gx = grid_for((size_t)B * K * L, 256, 64);
    style_mix_batched_kernel<<<dim3(gx, nS), 256, 0, (cudaStream_t)stream>>>(
        depth_vec, (const float* const*)A_ptrs, (const float* const*)a_ptrs, (__nv_bfloat16*)stp, B, K, L, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_build_mask16(const float* masks, void* mask16, int B, int K, int H, int W, void* stream) {
    DASR_REQUIRE(masks && mask16 && B > 0, "bad arguments");
    DASR_REQUIRE(K >= 1 && K <= 16, "mask image: at most 16 depth masks (got %d)", K);
    const size_t npix = (size_t)B * H * W;
    build_mask16_kernel<<<grid_for(npix, 256), 256, 0, (cudaStream_t)stream>>>(masks, (uint4*)mask16, K, H * W, npix);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_table_to_dynweights(const void* table, void* wdyn, int n, int K, int nf2, int group, void* stream) {
    DASR_REQUIRE(table && wdyn && n > 0, "bad arguments");
    if (group <= 0) group = n;
    DASR_REQUIRE(n % group == 0, "table_to_dynweights: n (%d) must be a multiple of the group size (%d)", n, group);
    DASR_REQUIRE(K >= 1 && K <= 16 && nf2 % 8 == 0, "dynamic-conv weights: K <= 16, 2*nf multiple of 8");
    const size_t smem = (size_t)K * 9 * nf2 * 2;
    DASR_REQUIRE(smem <= 48 * 1024, "table too large");
    table_to_dynweights_kernel<<<dim3(n, planes()), 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)table, (__nv_bfloat16*)wdyn, K, nf2, group, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_dynconv_fwd(const void* table, const uint8_t* labels, const float* masks, const int32_t* flag,
                                void* out, int B, int K, int H, int W, int nf2, void* stream) {
    DASR_REQUIRE(table && out && (labels || masks), "null pointer");
    DASR_REQUIRE(nf2 % 8 == 0, "2*nf must be a multiple of 8");
    DASR_REQUIRE(planes() == 1, "the stand-alone K-DYN kernel has no fp32-split form (the network folds K-DYN into the SEAN GEMM)");
    if (labels) {
        const int rows = 2;
        const size_t smem = (size_t)(K + 1) * 9 * nf2 * 2 + (size_t)(rows + 2) * (W + 2);
        DASR_REQUIRE(smem <= 200 * 1024, "image too wide for the label tile");
        static bool configured[64] = {false};
        int dev = 0;
        DASR_CUDA_OK(cudaGetDevice(&dev));
        if (!configured[dev & 63]) {
            DASR_CUDA_OK(cudaFuncSetAttribute(dynconv_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            configured[dev & 63] = true;
        }
        const int bands = (H + rows - 1) / rows;
        const int n_items = B * bands;
        const int grid = n_items < 4 * num_sms() ? n_items : 4 * num_sms();      // 4 resident blocks per SM
        dynconv_labels_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)table, labels, masks, flag, (uint4*)out, K, H, W, nf2, rows, n_items);
    } else {
        const size_t total = (size_t)B * H * W * (nf2 / 8);
        dynconv_masks_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)table, masks, (uint4*)out, B, K, H, W, nf2);
    }
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_instats_finalize(const float* stats, float* norm, float* normk, int B, int C, int HW, int nslots,
                                     void* stream) {
    DASR_REQUIRE(stats && norm && HW > 0 && nslots > 0, "bad arguments");
    const int n = B * C;
    instats_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, norm, normk, n, C, nslots, 1.f / (float)HW);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

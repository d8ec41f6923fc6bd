// Training-step kernels that sit behind the generator (HBM-bound, CUDA cores, vectorised):
//   K-LOSS   l_pix = w_pix * mean|SR-HR|  (nn.L1Loss, codes/models/F_model_depthCond.py:52,164) and the dynamic
//            depth-mask loss  l_dyn = l_w * sum_k softmax(w)_k * [ sum SmoothL1(m_k*SR, m_k*HR) / sum(m_k x3) ]
//            (dynamic_weight_mask_loss.forward, codes/models/modules/mask_loss.py:64-90) in ONE pass over SR/HR,
//            plus the backward pass that writes d(total)/d(SR).
//   K-ADAM   torch.optim.Adam step (F_model_depthCond.py:99-101,192) over one flat fp32 buffer.
//
// The per-mask sums are accumulated in registers (one predicated add per label), reduced per block in a fixed
// order and written to one partial-sum row per block; dasr_loss_reduce adds the rows in row order, so the loss is
// bit-reproducible run to run (no floating-point atomics on the one-hot path).
#include "dasr_internal.h"

namespace dasr {

constexpr int kLossK = DASR_LOSS_KMAX;        // label slots kept in registers
constexpr int kLossRow = DASR_LOSS_ROW;       // floats per partial-sum row: l1, num[16], cnt[16], pad -> 36

__device__ __forceinline__ float smooth_l1(float d) {
    const float a = fabsf(d);
    return a < 1.f ? 0.5f * d * d : a - 0.5f;
}
__device__ __forceinline__ float sgn(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

// legacy `mode='nearest'` source index of F.interpolate (mask_loss.py:73): floor(dst * in/out), clamped
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
    const int s = (int)floorf((float)dst * scale);
    return s < in_size - 1 ? s : in_size - 1;
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    // fixed-order reduction over 256 threads; every thread returns the total
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += red[i];
    return t;
}

// item = (b, y, x4): 4 horizontally adjacent HR pixels, all C channels (float4 loads from the NCHW planes)
__global__ void __launch_bounds__(256) loss_fwd_kernel(const float* __restrict__ sr, const float* __restrict__ hr,
                                                       const uint8_t* __restrict__ labels,
                                                       const float* __restrict__ masks, const int* __restrict__ flag,
                                                       float* __restrict__ part, int B, int C, int K, int h, int w,
                                                       int Ho, int Wo, float sy, float sx) {
    __shared__ float red[8];
    __shared__ float gen_s[2 * kLossK];
    const bool general = (labels == nullptr) || (flag != nullptr && *flag != 0);
    const int W4 = Wo >> 2;
    const size_t items = (size_t)B * Ho * W4;
    const size_t plane = (size_t)Ho * Wo;
    float l1 = 0.f;
    float num[kLossK], cnt[kLossK];
#pragma unroll
    for (int k = 0; k < kLossK; k++) num[k] = cnt[k] = 0.f;
    if (threadIdx.x < 2 * kLossK) gen_s[threadIdx.x] = 0.f;
    __syncthreads();

    for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (size_t)gridDim.x * blockDim.x) {
        const int x4 = (int)(it % W4);
        const size_t r = it / W4;
        const int y = (int)(r % Ho);
        const int b = (int)(r / Ho);
        const int ly = nearest_src(y, sy, h);
        const size_t base = (size_t)b * C * plane + (size_t)y * Wo + 4 * x4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};      // sum over channels of SmoothL1(d), per pixel (one-hot path)
        int lx[4];
#pragma unroll
        for (int i = 0; i < 4; i++) lx[i] = nearest_src(4 * x4 + i, sx, w);
        if (!general) {
            for (int c = 0; c < C; c++) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(sr + base + c * plane));
                const float4 g = __ldg(reinterpret_cast<const float4*>(hr + base + c * plane));
                const float d[4] = {a.x - g.x, a.y - g.y, a.z - g.z, a.w - g.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    l1 += fabsf(d[i]);
                    v[i] += smooth_l1(d[i]);
                }
            }
            const uint8_t* lp = labels + ((size_t)b * h + ly) * w;
            const int l0 = lp[lx[0]], l1b = lp[lx[1]], l2 = lp[lx[2]], l3 = lp[lx[3]];
            if (l0 == l1b && l1b == l2 && l2 == l3) {
                const float vs = (v[0] + v[1]) + (v[2] + v[3]);
#pragma unroll
                for (int k = 0; k < kLossK; k++) {
                    const bool hit = (l0 == k);
                    num[k] += hit ? vs : 0.f;
                    cnt[k] += hit ? 4.f : 0.f;
                }
            } else {
                const int ll[4] = {l0, l1b, l2, l3};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int k = 0; k < kLossK; k++) {
                        const bool hit = (ll[i] == k);
                        num[k] += hit ? v[i] : 0.f;
                        cnt[k] += hit ? 1.f : 0.f;
                    }
            }
        } else {
            // masks with arbitrary values: the reference's literal arithmetic m*SR - m*HR per mask (slow path;
            // per-mask sums go through shared-memory atomics)
            float4 a[4], g[4];
            for (int c = 0; c < C && c < 4; c++) {
                a[c] = __ldg(reinterpret_cast<const float4*>(sr + base + c * plane));
                g[c] = __ldg(reinterpret_cast<const float4*>(hr + base + c * plane));
                l1 += fabsf(a[c].x - g[c].x) + fabsf(a[c].y - g[c].y) + fabsf(a[c].z - g[c].z) + fabsf(a[c].w - g[c].w);
            }
            for (int k = 0; k < K; k++) {
                const float* mp = masks + (((size_t)b * K + k) * h + ly) * w;
                float nk = 0.f, ck = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float m = __ldg(mp + lx[i]);
                    if (m == 0.f) continue;
                    ck += m;
                    for (int c = 0; c < C && c < 4; c++) {
                        const float s = i == 0 ? a[c].x : i == 1 ? a[c].y : i == 2 ? a[c].z : a[c].w;
                        const float t = i == 0 ? g[c].x : i == 1 ? g[c].y : i == 2 ? g[c].z : g[c].w;
                        nk += smooth_l1(m * s - m * t);
                    }
                }
                if (ck != 0.f) {
                    atomicAdd(&gen_s[k], nk);
                    atomicAdd(&gen_s[kLossK + k], ck);
                }
            }
        }
    }
    float* row = part + (size_t)blockIdx.x * kLossRow;
    const float t = block_sum_256(l1, red);
    if (threadIdx.x == 0) row[0] = t;
    if (!general) {
#pragma unroll
        for (int k = 0; k < kLossK; k++) {
            const float a = block_sum_256(num[k], red);
            const float c = block_sum_256(cnt[k], red);
            if (threadIdx.x == 0) {
                row[1 + k] = a;
                row[1 + kLossK + k] = c;
            }
        }
    } else {
        __syncthreads();
        if (threadIdx.x < 2 * kLossK) row[1 + threadIdx.x] = gen_s[threadIdx.x];
    }
}

// sums[j] = sum over rows of part[row][j], rows added in row order (deterministic). One warp per column group.
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ part, float* __restrict__ sums,
                                                          int rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < 1 + 2 * kLossK; j += 8) {
        float s = 0.f;
        for (int r = lane; r < rows; r += 32) s += part[(size_t)r * kLossRow + j];
#pragma unroll
        for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) sums[j] = s;
    }
}

// sums (possibly all-reduced over the data-parallel ranks) -> loss values, backward coefficients, d l_dyn / d w
__global__ void loss_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ wdyn,
                                     float* __restrict__ out, int K, int C, double n_elems, float w_pix, float w_dyn) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float sw[kLossK], lk[kLossK];
    float mx = -INFINITY;
    for (int k = 0; k < K; k++) mx = fmaxf(mx, wdyn ? wdyn[k] : 0.f);
    float z = 0.f;
    for (int k = 0; k < K; k++) {
        sw[k] = expf((wdyn ? wdyn[k] : 0.f) - mx);
        z += sw[k];
    }
    float dyn = 0.f;
    for (int k = 0; k < K; k++) {
        sw[k] /= z;
        const float den = (float)C * sums[1 + kLossK + k];     // sum of the mask replicated over the C channels
        lk[k] = sums[1 + k] / den;                             // 0/0 = NaN for an empty mask, like the reference
        dyn += sw[k] * lk[k];
        out[4 + 0 * kLossK + k] = lk[k];
        out[4 + 1 * kLossK + k] = sw[k];
        out[4 + 2 * kLossK + k] = w_dyn * sw[k] / den;
    }
    for (int k = 0; k < K; k++) out[4 + 3 * kLossK + k] = w_dyn * sw[k] * (lk[k] - dyn);
    const float l_pix = w_pix * (float)((double)sums[0] / n_elems);
    const float l_dyn = w_dyn * dyn;
    out[0] = l_pix + l_dyn;
    out[1] = l_pix;
    out[2] = l_dyn;
    out[3] = (float)((double)w_pix / n_elems);
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(const float* __restrict__ sr, const float* __restrict__ hr,
                                                       const uint8_t* __restrict__ labels,
                                                       const float* __restrict__ masks, const int* __restrict__ flag,
                                                       const float* __restrict__ coef, const float* __restrict__ gup,
                                                       float use_pix, float use_dyn, float* __restrict__ dsr,
                                                       float* __restrict__ dwdyn, int B, int C, int K, int h, int w,
                                                       int Ho, int Wo, float sy, float sx) {
    __shared__ float ck_s[kLossK + 1];
    const bool general = (labels == nullptr) || (flag != nullptr && *flag != 0);
    // upstream gradients of (total, l_pix, l_dyn): the three leading entries of the forward's output vector
    const float gp = use_pix * (gup ? gup[0] + gup[1] : 1.f);
    const float gd = use_dyn * (gup ? gup[0] + gup[2] : 1.f);
    if (threadIdx.x < kLossK) ck_s[threadIdx.x] = threadIdx.x < K ? gd * coef[4 + 2 * kLossK + threadIdx.x] : 0.f;
    if (threadIdx.x == kLossK) ck_s[kLossK] = 0.f;       // label 255: pixel in no mask
    if (dwdyn && blockIdx.x == 0 && threadIdx.x < K) dwdyn[threadIdx.x] = gd * coef[4 + 3 * kLossK + threadIdx.x];
    __syncthreads();
    const float cp = gp * coef[3];
    const int W4 = Wo >> 2;
    const size_t items = (size_t)B * Ho * W4;
    const size_t plane = (size_t)Ho * Wo;
    for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (size_t)gridDim.x * blockDim.x) {
        const int x4 = (int)(it % W4);
        const size_t r = it / W4;
        const int y = (int)(r % Ho);
        const int b = (int)(r / Ho);
        const int ly = nearest_src(y, sy, h);
        const size_t base = (size_t)b * C * plane + (size_t)y * Wo + 4 * x4;
        int lx[4];
#pragma unroll
        for (int i = 0; i < 4; i++) lx[i] = nearest_src(4 * x4 + i, sx, w);
        float ck[4];
        if (!general) {
            const uint8_t* lp = labels + ((size_t)b * h + ly) * w;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int l = lp[lx[i]];
                ck[i] = ck_s[l < kLossK ? l : kLossK];
            }
        }
        for (int c = 0; c < C; c++) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(sr + base + c * plane));
            const float4 t = __ldg(reinterpret_cast<const float4*>(hr + base + c * plane));
            const float s[4] = {a.x, a.y, a.z, a.w}, q[4] = {t.x, t.y, t.z, t.w};
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float d = s[i] - q[i];
                float acc = cp * sgn(d);
                if (!general) {
                    acc += ck[i] * fminf(fmaxf(d, -1.f), 1.f);
                } else {
                    for (int k = 0; k < K; k++) {
                        const float m = __ldg(masks + (((size_t)b * K + k) * h + ly) * w + lx[i]);
                        if (m != 0.f) acc += ck_s[k] * m * fminf(fmaxf(m * s[i] - m * q[i], -1.f), 1.f);
                    }
                }
                o[i] = acc;
            }
            *reinterpret_cast<float4*>(dsr + base + c * plane) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ K-ADAM
// torch.optim.Adam (amsgrad=False, maximize=False):  g' = g + wd*p;  m = b1*m + (1-b1)*g';  v = b2*v + (1-b2)*g'^2;
// p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                   float4* __restrict__ m, float4* __restrict__ v, size_t n4,
                                                   float* __restrict__ pt, const float* __restrict__ gt,
                                                   float* __restrict__ mt, float* __restrict__ vt, int tail, float omb1,
                                                   float b2, float omb2, float eps, float wd, float step_size,
                                                   float sqrt_bc2, float gscale, const float* __restrict__ dev_scal) {
    if (dev_scal) {      // CUDA-graph replays: the step-dependent scalars live in device memory
        step_size = dev_scal[0];
        sqrt_bc2 = dev_scal[1];
    }
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg = gg * gscale;
        if (wd != 0.f) gg = fmaf(wd, pp, gg);
        mm = mm + omb1 * (gg - mm);                       // exp_avg.lerp_(grad, 1 - beta1)
        vv = __fadd_rn(__fmul_rn(b2, vv), __fmul_rn(__fmul_rn(omb2, gg), gg));   // mul_(beta2).addcmul_(g, g, 1 - beta2)
        const float denom = __fadd_rn(__fdiv_rn(sqrtf(vv), sqrt_bc2), eps);
        pp = __fsub_rn(pp, __fmul_rn(step_size, __fdiv_rn(mm, denom)));
    };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 P = p[i], M = m[i], V = v[i];
        const float4 G = __ldg(g + i);
        upd(P.x, G.x, M.x, V.x);
        upd(P.y, G.y, M.y, V.y);
        upd(P.z, G.z, M.z, V.z);
        upd(P.w, G.w, M.w, V.w);
        p[i] = P;
        m[i] = M;
        v[i] = V;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) upd(pt[threadIdx.x], gt[threadIdx.x], mt[threadIdx.x], vt[threadIdx.x]);
}

}  // namespace dasr

using namespace dasr;

static int loss_grid(size_t items) {
    size_t g = (items + 255) / 256;
    const size_t cap = (size_t)num_sms() * 4;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" int dasr_loss_rows(int B, int Ho, int Wo) {
    if (B <= 0 || Ho <= 0 || Wo <= 0 || (Wo & 3)) return fail(DASR_ERR_BAD_ARG, "loss: bad frame size %dx%d", Ho, Wo);
    return loss_grid((size_t)B * Ho * (Wo >> 2));
}

extern "C" int dasr_loss_fwd(const float* sr, const float* hr, const uint8_t* labels, const float* masks,
                             const int32_t* flag, float* part, float* sums, int B, int C, int K, int h, int w, int Ho,
                             int Wo, void* stream) {
    DASR_REQUIRE(sr && hr && part && sums && (labels || masks), "null pointer");
    DASR_REQUIRE(K >= 1 && K <= kLossK, "loss: number of depth masks must be 1..%d (got %d)", kLossK, K);
    DASR_REQUIRE(C >= 1 && C <= 4, "loss: 1..4 image channels (got %d)", C);
    DASR_REQUIRE(Wo % 4 == 0, "loss: HR width must be a multiple of 4 (got %d)", Wo);
    DASR_REQUIRE(labels == nullptr || masks != nullptr || flag == nullptr, "loss: the general-mask path needs masks");
    const int rows = loss_grid((size_t)B * Ho * (Wo >> 2));
    cudaStream_t st = (cudaStream_t)stream;
    loss_fwd_kernel<<<rows, 256, 0, st>>>(sr, hr, labels, masks, flag, part, B, C, K, h, w, Ho, Wo,
                                          (float)h / (float)Ho, (float)w / (float)Wo);
    DASR_LAUNCH_OK();
    loss_reduce_kernel<<<1, 256, 0, st>>>(part, sums, rows);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_loss_finalize(const float* sums, const float* wdyn, float* out, int K, int C, double n_elems,
                                  float w_pix, float w_dyn, void* stream) {
    DASR_REQUIRE(sums && out && K >= 1 && K <= kLossK && n_elems > 0, "bad arguments");
    loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, wdyn, out, K, C, n_elems, w_pix, w_dyn);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_loss_bwd(const float* sr, const float* hr, const uint8_t* labels, const float* masks,
                             const int32_t* flag, const float* out, const float* gup, float use_pix, float use_dyn,
                             float* dsr, float* dwdyn, int B, int C, int K, int h, int w, int Ho, int Wo,
                             void* stream) {
    DASR_REQUIRE(sr && hr && out && dsr && (labels || masks), "null pointer");
    DASR_REQUIRE(K >= 1 && K <= kLossK && C >= 1 && C <= 4 && Wo % 4 == 0, "loss: unsupported shape");
    const int grid = loss_grid((size_t)B * Ho * (Wo >> 2)) * 2;
    loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sr, hr, labels, masks, flag, out, gup, use_pix, use_dyn, dsr,
                                                            dwdyn, B, C, K, h, w, Ho, Wo, (float)h / (float)Ho,
                                                            (float)w / (float)Wo);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1,
                              double beta2, double eps, double weight_decay, int64_t step, double grad_scale,
                              const float* dev_scalars, void* stream) {
    DASR_REQUIRE(p && g && m && v && n > 0 && (step >= 1 || dev_scalars), "bad arguments");
    if (step < 1) step = 1;
    DASR_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0,
                 "adam: buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    const float step_size = (float)(lr / bc1);
    const float sqrt_bc2 = (float)sqrt(bc2);
    const size_t n4 = (size_t)n / 4;
    const int tail = (int)(n - (int64_t)n4 * 4);
    size_t grid = (n4 + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    adam_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, n4,
                                                             p + n4 * 4, g + n4 * 4, m + n4 * 4, v + n4 * 4, tail,
                                                             (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                             (float)eps, (float)weight_decay, step_size, sqrt_bc2,
                                                             (float)grad_scale, dev_scalars);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

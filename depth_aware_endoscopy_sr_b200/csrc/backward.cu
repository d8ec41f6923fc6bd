// Memory-bound kernels of the DepthNet BACKWARD pass (autograd of the reference graph, reached through
// total_loss.backward() at codes/models/F_model_depthCond.py:191).  All activations NHWC bf16, all
// reductions fp32 with a fixed summation order per block (cross-block sums use fp32 atomics where noted).
//
//   sean_bwd1 / sean_bwd2   SEAN modulate + double InstanceNorm backward (+ [gamma_o; beta_o] bias gradient)
//                                               (normalization.py:56,87-89; sftmd_arch.py:813,820,828,832-833)
//   colsum            bias gradients (sum over pixels of a gradient tensor)
//   dynconv_bwd       K-DYN backward: per-image table gradient dT (normalization.py:81-85 restated)
//   table_bwd_w/_s    style-table GEMM backward (dW_s, d st')
//   style_mix_bwd     A_i_j backward (normalization.py:27,80) + accumulation of d depthVec
//   region_pool_bwd   RegionWiseAvgPooling backward (sftmd_arch.py:714-733)
//   actv_bwd          mlp_mask (1 -> 2nf conv) weight / bias gradient (normalization.py:37-40)
//   unshuffle_actgrad PixelShuffle(2) + LeakyReLU backward as one gather (sftmd_arch.py:893-908)
//   out9_bwd_prep     clamp backward + im2row of the output gradient for the 9x9 conv (sftmd_arch.py:948-950)
//   nchw3_to_nhwc32   LR image as a 32-channel NHWC bf16 tensor (operand of encoder.layer1's weight gradient)
#include "dasr_internal.h"
#include <stdlib.h>

namespace dasr {

__device__ __forceinline__ uint4 bpack8(const float* f) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}
__device__ __forceinline__ void bunpack8(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

// ------------------------------------------------------------------------------------ SEAN backward, pass 1
// block = (image b, pixel chunk); thread = (8-channel group g, pixel lane).  G = nf / 8.
//   dz = dout * [act_out > 0]      (ReLU after the modulation / after the residual add)
//   n  = (y - mean) * scale ;  dgamma = dz * n ; dbeta = dz ; dn = dz * (1 + gamma)
//   part[b][slot][c] = (sum dn, sum dn * n) over the block's pixels
template <int G>
__global__ void __launch_bounds__(256) sean_bwd1_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ act_out,
                                                        const uint4* __restrict__ y, const float* __restrict__ norm,
                                                        const uint4* __restrict__ gamma, uint4* __restrict__ dgb,
                                                        uint4* __restrict__ dn_out, uint4* __restrict__ dskip,
                                                        float* __restrict__ part, int HW, int pix_per_block, int nslots,
                                                        int npl) {
    constexpr int NF = G * 8;
    constexpr int LANES = 256 / G;
    const int b = blockIdx.y, slot = blockIdx.x;
    const size_t ps8 = (size_t)gridDim.y * HW * G;          // plane stride of the [B,HW,nf] tensors (uint4 units)
    const int g = threadIdx.x % G, pl = threadIdx.x / G;
    float mean[8], scale[8], acc[4][8];          // acc: sum dn, sum dn*n, sum dz*n (dgamma), sum dz (dbeta)
#pragma unroll
    for (int j = 0; j < 8; j++) {
        mean[j] = __ldg(norm + ((size_t)b * NF + g * 8 + j) * 2);
        scale[j] = __ldg(norm + ((size_t)b * NF + g * 8 + j) * 2 + 1);
        acc[0][j] = acc[1][j] = acc[2][j] = acc[3][j] = 0.f;
    }
    const int p0 = slot * pix_per_block, p1 = min(p0 + pix_per_block, HW);
#pragma unroll 2
    for (int pix = p0 + pl; pix < p1; pix += LANES) {
        const size_t i = ((size_t)b * HW + pix) * G + g;
        float d[8], a[8], yv[8], gm[8], dg[8], dnv[8];
        bunpack8(__ldg(act_out + i), a);      // sign only: plane 0
        if (npl == 1) {
            bunpack8(__ldg(dout + i), d);
            bunpack8(__ldg(y + i), yv);
            bunpack8(__ldg(gamma + i), gm);
        } else {
            pl_load8(dout, i, ps8, npl, d);
            pl_load8(y, i, ps8, npl, yv);
            pl_load8(gamma, i, ps8, npl, gm);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float dz = a[j] > 0.f ? d[j] : 0.f;
            const float n = (yv[j] - mean[j]) * scale[j];
            d[j] = dz;
            dg[j] = dz * n;
            dnv[j] = dz * (1.f + gm[j]);
            acc[0][j] += dnv[j];
            acc[1][j] += dnv[j] * n;
            acc[2][j] += dg[j];
            acc[3][j] += dz;
        }
        const size_t o = ((size_t)b * HW + pix) * (2 * G);
        if (npl == 1) {
            dgb[o + g] = bpack8(dg);
            dgb[o + G + g] = bpack8(d);
            dn_out[i] = bpack8(dnv);
            if (dskip) dskip[i] = bpack8(d);
        } else {
            float t[8];
            pl_store8(dgb, o + g, 2 * ps8, npl, dg);
#pragma unroll
            for (int j = 0; j < 8; j++) t[j] = d[j];
            pl_store8(dgb, o + G + g, 2 * ps8, npl, t);
            pl_store8(dn_out, i, ps8, npl, dnv);
            if (dskip) pl_store8(dskip, i, ps8, npl, d);
        }
    }
    // lanes of a warp that share g (stride G) first, then the 8 warps through shared memory
    __shared__ float red[4][8][NF];
#pragma unroll
    for (int w = 0; w < 4; w++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float v = acc[w][j];
#pragma unroll
            for (int off = G; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            acc[w][j] = v;
        }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < G) {
#pragma unroll
        for (int w = 0; w < 4; w++)
#pragma unroll
            for (int j = 0; j < 8; j++) red[w][warp][g * 8 + j] = acc[w][j];
    }
    __syncthreads();
    if (threadIdx.x < 4 * NF) {
        const int which = threadIdx.x / NF, c = threadIdx.x - which * NF;
        float sum = 0.f;
#pragma unroll
        for (int l = 0; l < 8; l++) sum += red[which][l][c];
        part[(((size_t)b * nslots + slot) * NF + c) * 4 + which] = sum;
    }
}

// pass 2.  Prologue (per CTA, from the partial sums of pass 1):
//   coef[c] = (c1, c2):  dy = scale * (dn - c1) + c2 * n,  c1 = S1 / N,  c2 = -k * T2 / (N * scale)
// with k = 1/a + eps/(a^2 r), a = var + eps, r = var/a + eps   (norm[b][c] = (mean, scale); normk[b][c] = k);
// the first CTA of every image also adds the image's share of the [gamma_o; beta_o] bias gradient to dbias.
template <int G>
__global__ void __launch_bounds__(256) sean_bwd2_kernel(const uint4* __restrict__ dn, const uint4* __restrict__ y,
                                                        const float* __restrict__ norm, const float* __restrict__ normk,
                                                        const float* __restrict__ part, uint4* __restrict__ dy,
                                                        float* __restrict__ dbias, int HW, int pix_per_block,
                                                        int nslots, int npl) {
    constexpr int NF = G * 8;
    constexpr int LANES = 256 / G;
    const int b = blockIdx.y;
    const size_t ps8 = (size_t)gridDim.y * HW * G;
    const int g = threadIdx.x % G, pl = threadIdx.x / G;
    __shared__ float sums[2][NF];
    if (threadIdx.x < 4 * NF) {
        const int which = threadIdx.x / NF, c = threadIdx.x - which * NF;
        if (which < 2 || (dbias && blockIdx.x == 0)) {
            const float* sp = part + ((size_t)b * nslots * NF + c) * 4 + which;
            float s0 = 0.f, s1 = 0.f;
            int sl = 0;
            for (; sl + 1 < nslots; sl += 2) {
                s0 += __ldg(sp + (size_t)sl * NF * 4);
                s1 += __ldg(sp + (size_t)(sl + 1) * NF * 4);
            }
            if (sl < nslots) s0 += __ldg(sp + (size_t)sl * NF * 4);
            if (which < 2)
                sums[which][c] = s0 + s1;
            else
                atomicAdd(dbias + (which - 2) * NF + c, s0 + s1);
        }
    }
    __syncthreads();
    const float inv_hw = 1.f / (float)HW;
    float mean[8], scale[8], c1[8], c2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int c = g * 8 + j;
        const size_t k = (size_t)b * NF + c;
        mean[j] = __ldg(norm + k * 2);
        scale[j] = __ldg(norm + k * 2 + 1);
        c1[j] = sums[0][c] * inv_hw;
        c2[j] = -__ldg(normk + k) * sums[1][c] * inv_hw / scale[j];
    }
    const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, HW);
#pragma unroll 2
    for (int pix = p0 + pl; pix < p1; pix += LANES) {
        const size_t i = ((size_t)b * HW + pix) * G + g;
        float d[8], yv[8];
        if (npl == 1) {
            bunpack8(__ldg(dn + i), d);
            bunpack8(__ldg(y + i), yv);
        } else {
            pl_load8(dn, i, ps8, npl, d);
            pl_load8(y, i, ps8, npl, yv);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float n = (yv[j] - mean[j]) * scale[j];
            d[j] = scale[j] * (d[j] - c1[j]) + c2[j] * n;
        }
        if (npl == 1) dy[i] = bpack8(d);
        else pl_store8(dy, i, ps8, npl, d);
    }
}

// ------------------------------------------------------------------------------------ column sums
// out[c] += sum_rows x[row][c]   (x bf16 [rows][C], C multiple of 8, C <= 2048); fp32 atomics across blocks
__global__ void __launch_bounds__(256) colsum_kernel(const uint4* __restrict__ x, float* __restrict__ out, size_t rows,
                                                     int G, size_t rows_per_block) {
    const int LANES = 256 / G;
    const int g = threadIdx.x % G, pl = threadIdx.x / G;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t r0 = blockIdx.x * rows_per_block;
    const size_t r1 = min(r0 + rows_per_block, rows);
    if (pl < LANES)
        for (size_t r = r0 + pl; r < r1; r += 4 * (size_t)LANES) {      // four independent 16-byte loads in flight
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                v[u] = (r + u * (size_t)LANES < r1) ? __ldg(x + (r + u * (size_t)LANES) * G + g) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float f[8];
                bunpack8(v[u], f);
#pragma unroll
                for (int j = 0; j < 8; j++) s[j] += f[j];
            }
        }
    extern __shared__ float red[];   // [LANES][G*8 + 1]
    const int ld = G * 8 + 1;
    if (pl < LANES)
#pragma unroll
        for (int j = 0; j < 8; j++) red[pl * ld + g * 8 + j] = s[j];
    __syncthreads();
    for (int c = threadIdx.x; c < G * 8; c += blockDim.x) {
        float t = 0.f;
        for (int l = 0; l < LANES; l++) t += red[l * ld + c];
        atomicAdd(out + c, t);
    }
}

// ------------------------------------------------------------------------------------ K-DYN backward
// dT[b][k][tap][c] += sum_p dgb[b,p,c] * [label(p + tap - 1) == k]      (one-hot masks)
//                  += sum_p dgb[b,p,c] * mask[b,k,p + tap - 1]           (general masks)
// block = (image, band of rows); thread c owns column c of a shared [K*9][C2] fp32 accumulator (no atomics
// inside the block); bands are combined with fp32 atomics.
__global__ void __launch_bounds__(128) dynconv_bwd_kernel(const __nv_bfloat16* __restrict__ dgb,
                                                          const uint8_t* __restrict__ labels,
                                                          const float* __restrict__ masks, const int* __restrict__ flag,
                                                          float* __restrict__ dT, int K, int H, int W, int C2,
                                                          int rows_per_block, int npl, size_t ps) {
    extern __shared__ float acc[];   // [K*9][C2]
    const int bands = (H + rows_per_block - 1) / rows_per_block;
    const int b = blockIdx.x / bands, band = blockIdx.x % bands;
    const int h0 = band * rows_per_block, h1 = min(h0 + rows_per_block, H);
    const int c = threadIdx.x;
    if (c >= C2) return;
    // fallback role (labels == NULL, flag given): dasr_dynconv_bwd_tc already handled one-hot masks
    if (labels == nullptr && flag != nullptr && *flag == 0) return;
    for (int i = 0; i < K * 9; i++) acc[i * C2 + c] = 0.f;
    const bool general = (labels == nullptr) || (flag != nullptr && *flag != 0);
    for (int y = h0; y < h1; y++)
        for (int x = 0; x < W; x++) {
            const float v = pl_load1(dgb, (((size_t)b * H + y) * W + x) * C2 + c, ps, npl);
#pragma unroll
            for (int t = 0; t < 3; t++)
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const int yy = y + t - 1, xx = x + u - 1;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    if (!general) {
                        const int lab = labels[((size_t)b * H + yy) * W + xx];
                        if (lab != 255) acc[(lab * 9 + t * 3 + u) * C2 + c] += v;
                    } else {
                        for (int k = 0; k < K; k++) {
                            const float m = __ldg(masks + (((size_t)b * K + k) * H + yy) * W + xx);
                            if (m != 0.f) acc[(k * 9 + t * 3 + u) * C2 + c] += m * v;
                        }
                    }
                }
        }
    float* dst = dT + (size_t)b * K * 9 * C2;
    for (int i = 0; i < K * 9; i++) atomicAdd(dst + (size_t)i * C2 + c, acc[i * C2 + c]);
}

// ------------------------------------------------------------------------------------ style-table GEMM backward
// T[bk][n] = sum_c stp[bk][c] * Ws[n][c]  (n = tap*C2 + o2)
//   dWs[n][c]   = sum_bk dT[bk][n] * stp[bk][c]      (A = dT read transposed, B = stp)
//   dstp[bk][c] = sum_n  dT[bk][n] * Ws[n][c]        (A = dT, B = Ws; split-K with fp32 atomics)
// Both are small fp32 x bf16 GEMMs (94 MFLOP per instance at B = 16, 26 instances per step): one shared-memory tiled
// kernel, (16 TM) x 64 x 16 tiles, TM x 4 outputs per thread, operands read from shared memory as 128-bit vectors, the
// next K tile prefetched into registers while the current one is multiplied (two shared-memory buffers, one barrier
// per K tile).  C[m][n] (+)= sum_k A(m,k) * B[k][n];  A(m,k) = AT ? A[k*lda + m] : A[m*lda + k].
template <bool AT, int TM>
__global__ void __launch_bounds__(256) gemm_f32_bf16_kernel(const float* __restrict__ A, const __nv_bfloat16* __restrict__ Bm,
                                                            float* __restrict__ C, int M, int N, int K, int lda, int ldb,
                                                            int ldc, int k_per_split, int atomic, int splits,
                                                            long long sA, long long sB, long long sC, int npl,
                                                            size_t psB) {
    constexpr int BM = 16 * TM, BN = 64, BK = 16;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    // blockIdx.z = batch * splits + split (a batch of independent GEMMs with strides sA / sB / sC)
    const int batch = blockIdx.z / splits, split = blockIdx.z - batch * splits;
    A += (size_t)batch * sA;
    Bm += (size_t)batch * sB;
    C += (size_t)batch * sC;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = split * k_per_split, kend = min(K, kbeg + k_per_split);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    float ra[TM], rb[4];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < TM; i++) {
            const int idx = threadIdx.x + i * 256;
            int r, kk;
            if (AT) { kk = idx / BM; r = idx - kk * BM; } else { r = idx >> 4; kk = idx & 15; }
            const int m = m0 + r, k = k0 + kk;
            ra[i] = 0.f;
            if (m < M && k < kend) ra[i] = AT ? __ldg(A + (size_t)k * lda + m) : __ldg(A + (size_t)m * lda + k);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int idx = threadIdx.x + i * 256;
            const int kq = k0 + (idx >> 6), n = n0 + (idx & 63);
            rb[i] = (kq < kend && n < N) ? pl_load1(Bm, (size_t)kq * ldb + n, psB, npl) : 0.f;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < TM; i++) {
            const int idx = threadIdx.x + i * 256;
            int r, kk;
            if (AT) { kk = idx / BM; r = idx - kk * BM; } else { r = idx >> 4; kk = idx & 15; }
            As[buf][kk][r] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int idx = threadIdx.x + i * 256;
            Bs[buf][idx >> 6][idx & 63] = rb[i];
        }
    };
    if (kbeg < kend) {
        gload(kbeg);
        sstore(0);
    }
    __syncthreads();
    int cur = 0;
    for (int k0 = kbeg; k0 < kend; k0 += BK, cur ^= 1) {
        const bool more = k0 + BK < kend;
        if (more) gload(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[TM];
#pragma unroll
            for (int i4 = 0; i4 < TM; i4 += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&As[cur][kk][ty * TM + i4]);
                a[i4] = v.x; a[i4 + 1] = v.y; a[i4 + 2] = v.z; a[i4 + 3] = v.w;
            }
            const float4 b = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
#pragma unroll
            for (int i = 0; i < TM; i++) {
                acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
                acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
                acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
            }
        }
        if (more) sstore(cur ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            if (atomic) atomicAdd(C + (size_t)m * ldc + n, acc[i][j]);
            else C[(size_t)m * ldc + n] = acc[i][j];
        }
    }
}

// ------------------------------------------------------------------------------------ style mix backward
// stp[b][j][c] = sum_i A[j][i] vec[b][i][c] + a[j]
//   dA[j][i] += sum_{b,c} dstp[b][j][c] vec[b][i][c] ; da[j] += sum_{b,c} dstp[b][j][c]     block = (j, i)
//   dvec[b][i][c] += sum_j A[j][i] dstp[b][j][c]                                             (second kernel)
__global__ void __launch_bounds__(256) style_mix_bwd_a_kernel(const float* __restrict__ dstp, const float* __restrict__ vec,
                                                              float* __restrict__ dA, float* __restrict__ da, int B, int K,
                                                              int L) {
    const int j = blockIdx.x / K, i = blockIdx.x % K;
    float s = 0.f, sa = 0.f;
    for (int idx = threadIdx.x; idx < B * L; idx += blockDim.x) {
        const int b = idx / L, c = idx - b * L;
        const float d = dstp[((size_t)b * K + j) * L + c];
        s = fmaf(d, vec[((size_t)b * K + i) * L + c], s);
        sa += d;
    }
    __shared__ float r0[256], r1[256];
    r0[threadIdx.x] = s;
    r1[threadIdx.x] = sa;
    __syncthreads();
    for (int off = 128; off; off >>= 1) {
        if (threadIdx.x < off) {
            r0[threadIdx.x] += r0[threadIdx.x + off];
            r1[threadIdx.x] += r1[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        dA[j * K + i] += r0[0];
        if (i == 0) da[j] += r1[0];
    }
}
__global__ void style_mix_bwd_v_kernel(const float* __restrict__ dstp, const float* __restrict__ A,
                                       float* __restrict__ dvec, int B, int K, int L) {
    const size_t total = (size_t)B * K * L;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = idx % L;
        const int i = (idx / L) % K;
        const int b = idx / ((size_t)L * K);
        float s = 0.f;
        for (int j = 0; j < K; j++) s = fmaf(A[j * K + i], dstp[((size_t)b * K + j) * L + c], s);
        dvec[idx] += s;
    }
}

// the same for nS SEAN instances: A / dA / da are device pointer tables, dstp is [nS][B][K][L]; dvec sums the
// instances in index order (deterministic)
__global__ void __launch_bounds__(256) style_mix_bwd_a_batched_kernel(const float* __restrict__ dstp,
                                                                      const float* __restrict__ vec,
                                                                      float* const* __restrict__ dA_ptrs,
                                                                      float* const* __restrict__ da_ptrs, int B, int K,
                                                                      int L) {
    const int sidx = blockIdx.y;
    const int j = blockIdx.x / K, i = blockIdx.x % K;
    const float* ds = dstp + (size_t)sidx * B * K * L;
    float s = 0.f, sa = 0.f;
    for (int idx = threadIdx.x; idx < B * L; idx += blockDim.x) {
        const int b = idx / L, c = idx - b * L;
        const float d = ds[((size_t)b * K + j) * L + c];
        s = fmaf(d, vec[((size_t)b * K + i) * L + c], s);
        sa += d;
    }
    __shared__ float r0[256], r1[256];
    r0[threadIdx.x] = s;
    r1[threadIdx.x] = sa;
    __syncthreads();
    for (int off = 128; off; off >>= 1) {
        if (threadIdx.x < off) {
            r0[threadIdx.x] += r0[threadIdx.x + off];
            r1[threadIdx.x] += r1[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        dA_ptrs[sidx][j * K + i] += r0[0];
        if (i == 0) da_ptrs[sidx][j] += r1[0];
    }
}
__global__ void style_mix_bwd_v_batched_kernel(const float* __restrict__ dstp, const float* const* __restrict__ A_ptrs,
                                               float* __restrict__ dvec, int nS, int B, int K, int L) {
    const size_t total = (size_t)B * K * L;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = idx % L;
        const int i = (idx / L) % K;
        const int b = idx / ((size_t)L * K);
        float s = 0.f;
        for (int sidx = 0; sidx < nS; sidx++) {
            const float* A = A_ptrs[sidx];
            const float* ds = dstp + (size_t)sidx * total;
            for (int j = 0; j < K; j++) s = fmaf(__ldg(A + j * K + i), ds[((size_t)b * K + j) * L + c], s);
        }
        dvec[idx] += s;
    }
}

// ------------------------------------------------------------------------------------ region pooling backward
// de5[b][p][c] = sum_k msel[b][k][p] * dvec[b][k][c] / (cnt[b][k] + 1e-10)      (msel, cnt saved by the forward)
__global__ void region_pool_bwd_kernel(const float* __restrict__ dvec, const float* __restrict__ msel,
                                       const float* __restrict__ cnt, __nv_bfloat16* __restrict__ de5, int B, int P, int C,
                                       int K, int npl) {
    const size_t total = (size_t)B * P * C;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = idx % C;
        const int p = (idx / C) % P;
        const int b = idx / ((size_t)C * P);
        float s = 0.f;
        for (int k = 0; k < K; k++) {
            const float m = msel[((size_t)b * K + k) * P + p];
            if (m != 0.f) s += m * dvec[((size_t)b * K + k) * C + c] / (cnt[b * K + k] + 1e-10f);
        }
        pl_store1(de5, idx, total, npl, s);
    }
}

// ------------------------------------------------------------------------------------ actv (mlp_mask) backward
// dW[c][tap] += sum_{b,p} dA[b,p,c] * depth[b, p + tap - 1] ;  db[c] += sum dA        (dA already ReLU-masked)
__global__ void __launch_bounds__(128) actv_bwd_kernel(const __nv_bfloat16* __restrict__ dA, const float* __restrict__ depth,
                                                       float* __restrict__ dW, float* __restrict__ db, int H, int W, int C,
                                                       int rows_per_block) {
    const int bands = (H + rows_per_block - 1) / rows_per_block;
    const int b = blockIdx.x / bands, band = blockIdx.x % bands;
    const int h0 = band * rows_per_block, h1 = min(h0 + rows_per_block, H);
    const int c = threadIdx.x;
    if (c >= C) return;
    const float* dp = depth + (size_t)b * H * W;
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ab = 0.f;
    for (int y = h0; y < h1; y++)
        for (int x = 0; x < W; x++) {
            const float v = __bfloat162float(dA[(((size_t)b * H + y) * W + x) * C + c]);
            ab += v;
#pragma unroll
            for (int t = 0; t < 3; t++)
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const int yy = y + t - 1, xx = x + u - 1;
                    const float d = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(dp + yy * W + xx) : 0.f;
                    acc[t * 3 + u] = fmaf(v, d, acc[t * 3 + u]);
                }
        }
#pragma unroll
    for (int t = 0; t < 9; t++) atomicAdd(dW + c * 9 + t, acc[t]);
    atomicAdd(db + c, ab);
}

// ------------------------------------------------------------------------------------ PixelShuffle + LeakyReLU backward
// dconv[b,h,w,s*Cq + c] = dps[b,r*h+i,r*w+j,c] * (ps_out[b,r*h+i,r*w+j,c] > 0 ? 1 : slope),  s = r*i + j
// (the convolution in front of the shuffle is packed in this "shuffled" channel order, DASR_PACK_* shuffle_r)
__global__ void unshuffle_actgrad_kernel(const uint4* __restrict__ dps, const uint4* __restrict__ ps_out,
                                         uint4* __restrict__ dconv, int B, int H, int W, int Gq, float slope, int r,
                                         int npl) {
    const int r2 = r * r;
    const size_t total = (size_t)B * H * W * r2 * Gq;   // uint4 items of the output
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int g = idx % Gq;
        const int s = (idx / Gq) % r2;
        const size_t pix = idx / ((size_t)Gq * r2);
        const int w = pix % W;
        const int h = (pix / W) % H;
        const int b = pix / ((size_t)W * H);
        const int i = s / r, j = s - i * r;
        const size_t src = ((((size_t)b * r * H + r * h + i) * r * W) + r * w + j) * Gq + g;
        float d[8], o[8];
        if (npl == 1) bunpack8(__ldg(dps + src), d);
        else pl_load8(dps, src, total, npl, d);
        bunpack8(__ldg(ps_out + src), o);       // sign only: plane 0
#pragma unroll
        for (int k = 0; k < 8; k++) d[k] *= (o[k] > 0.f ? 1.f : slope);
        if (npl == 1) dconv[idx] = bpack8(d);
        else pl_store8(dconv, idx, total, npl, d);
    }
}

// PixelShuffle(r) of an NHWC tensor whose channels are in the shuffled order above (sftmd_arch.py:904-908, the x3
// tail): out[b, r*h+i, r*w+j, c] = in[b, h, w, (r*i+j)*Cq + c].  (r = 2 never runs this: it is the store addressing
// of the convolution's EPI_SHUFFLE2 epilogue.)
__global__ void pixel_shuffle_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int Gq,
                                     int r) {
    const int r2 = r * r;
    const size_t total = (size_t)B * H * W * r2 * Gq;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int g = idx % Gq;
        const int s = (idx / Gq) % r2;
        const size_t pix = idx / ((size_t)Gq * r2);
        const int w = pix % W;
        const int h = (pix / W) % H;
        const int b = pix / ((size_t)W * H);
        const int i = s / r, j = s - i * r;
        const size_t dst = ((((size_t)b * r * H + r * h + i) * r * W) + r * w + j) * Gq + g;
        out[dst] = __ldg(in + idx);
    }
}

// ------------------------------------------------------------------------------------ output conv backward prep
// g = dout * [0 < sr < 1]  (clamp backward);  A'[b,h,w,u*3 + co] = g[b,co,h,w - (u - 4)]  (zero outside), 27 of 32
// channels used;  dbias[co] += sum g
__global__ void __launch_bounds__(256) out9_bwd_prep_kernel(const float* __restrict__ dout, const float* __restrict__ sr,
                                                            __nv_bfloat16* __restrict__ aprime, float* __restrict__ dbias,
                                                            int B, int H, int W, int npl) {
    const size_t plane = (size_t)H * W;
    const size_t total = (size_t)B * plane;
    float sb[3] = {0.f, 0.f, 0.f};
    for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (size_t)gridDim.x * blockDim.x) {
        const int w = pix % W;
        const size_t bh = pix / W;          // b*H + h
        const size_t b = bh / H;
        const size_t row = b * 3 * plane + (bh - b * H) * W;
        float out[32];
#pragma unroll
        for (int i = 27; i < 32; i++) out[i] = 0.f;
#pragma unroll
        for (int u = 0; u < 9; u++) {
            const int ww = w - (u - 4);
            const bool in = ww >= 0 && ww < W;
#pragma unroll
            for (int co = 0; co < 3; co++) {
                float gv = 0.f;
                if (in) {
                    const size_t k = row + co * plane + ww;
                    const float s = __ldg(sr + k);
                    gv = (s > 0.f && s < 1.f) ? __ldg(dout + k) : 0.f;
                }
                out[u * 3 + co] = gv;
                if (u == 4) sb[co] += gv;
            }
        }
        uint4* op = reinterpret_cast<uint4*>(aprime + pix * 32);
        if (npl == 1) {
#pragma unroll
            for (int q = 0; q < 4; q++) op[q] = bpack8(out + 8 * q);
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) pl_store8(reinterpret_cast<uint4*>(aprime), pix * 4 + q, total * 4, npl, out + 8 * q);
        }
    }
    __shared__ float red[3][256];
    for (int co = 0; co < 3; co++) red[co][threadIdx.x] = sb[co];
    __syncthreads();
    for (int off = 128; off; off >>= 1) {
        if (threadIdx.x < off)
            for (int co = 0; co < 3; co++) red[co][threadIdx.x] += red[co][threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x < 3) atomicAdd(dbias + threadIdx.x, red[threadIdx.x][0]);
}

// x NCHW fp32 [B,3,H,W] -> NHWC bf16 [B,H,W,32] (channels 3..31 zero)
__global__ void nchw3_to_nhwc32_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int H, int W,
                                       int npl) {
    const size_t plane = (size_t)H * W;
    const size_t total = (size_t)B * plane;
    for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (size_t)gridDim.x * blockDim.x) {
        const size_t b = pix / plane, p = pix - b * plane;
        float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 3; c++) f[c] = __ldg(x + (b * 3 + c) * plane + p);
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int k = 0; k < npl; k++) {
            uint4* op = reinterpret_cast<uint4*>(out + ((size_t)k * total + pix) * 32);
            const uint4 u = bpack8(f);
            op[0] = u;
            op[1] = z;
            op[2] = z;
            op[3] = z;
            float t[8];
            bunpack8(u, t);
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] -= t[j];
        }
    }
}

// out = d * (act_out > 0 ? 1 : slope)      (ReLU / LeakyReLU backward, out of place)
__global__ void actgrad_kernel(const uint4* __restrict__ d, const uint4* __restrict__ act_out, uint4* __restrict__ out,
                               size_t n8, float slope, int npl) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        float g[8], a[8];
        if (npl == 1) bunpack8(__ldg(d + i), g);
        else pl_load8(d, i, n8, npl, g);
        bunpack8(__ldg(act_out + i), a);        // sign only: plane 0
#pragma unroll
        for (int j = 0; j < 8; j++) g[j] *= (a[j] > 0.f ? 1.f : slope);
        if (npl == 1) out[i] = bpack8(g);
        else pl_store8(out, i, n8, npl, g);
    }
}

// zero-stuffed copy with an explicit output size: out[b,ho,wo,:] = x[b,ho/2,wo/2,:] when ho, wo are even and inside
// the source, else 0  (gradient of a stride-2 convolution on the stride-1 grid)
__global__ void zero_insert2_to_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B, int H, int W, int C8,
                                       int Ho, int Wo) {
    const size_t total = (size_t)B * Ho * Wo * C8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = i % C8;
        size_t pix = i / C8;
        const int wo = pix % Wo;
        const int ho = (pix / Wo) % Ho;
        const int b = pix / ((size_t)Wo * Ho);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (!(ho & 1) && !(wo & 1) && (ho >> 1) < H && (wo >> 1) < W)
            v = __ldg(x + (((size_t)b * H + (ho >> 1)) * W + (wo >> 1)) * C8 + c);
        out[i] = v;
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

static int sean_ppb(int HW) {
    // pixels per block: ~16 blocks per image, at least 256 pixels each (measured at B = 16, 64x64: 16 blocks per
    // image 8.37 ms per training step, 32 -> 8.42, 64 -> 8.60: the per-block reduction is the fixed cost)
    int ppb = (HW + 15) / 16;
    if (ppb < 256) ppb = 256;
    return ppb;
}
extern "C" int dasr_sean_bwd_slots(int HW) {
    const int ppb = sean_ppb(HW);
    return (HW + ppb - 1) / ppb;
}

extern "C" int dasr_sean_bwd1(const void* dout, const void* act_out, const void* y, const float* norm, const void* gamma,
                              void* dgb, void* dn, void* dskip, float* part, int B, int HW, int nf, void* stream) {
    DASR_REQUIRE(dout && act_out && y && norm && gamma && dgb && dn && part, "null pointer");
    DASR_REQUIRE(nf == 64 || nf == 32, "nf must be 32 or 64 (got %d)", nf);
    const int ppb = sean_ppb(HW), slots = (HW + ppb - 1) / ppb;
    dim3 grid(slots, B);
    if (nf == 64)
        sean_bwd1_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)dout, (const uint4*)act_out, (const uint4*)y, norm, (const uint4*)gamma, (uint4*)dgb, (uint4*)dn, (uint4*)dskip, part, HW, ppb, slots, planes());
    else
        sean_bwd1_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)dout, (const uint4*)act_out, (const uint4*)y, norm, (const uint4*)gamma, (uint4*)dgb, (uint4*)dn, (uint4*)dskip, part, HW, ppb, slots, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_sean_bwd2(const void* dn, const void* y, const float* norm, const float* normk, const float* part,
                              void* dy, float* dbias, int B, int HW, int nf, void* stream) {
    DASR_REQUIRE(dn && y && norm && normk && part && dy, "null pointer");
    DASR_REQUIRE(nf == 64 || nf == 32, "nf must be 32 or 64 (got %d)", nf);
    const int ppb = sean_ppb(HW), slots = (HW + ppb - 1) / ppb;
    dim3 grid(slots, B);
    if (nf == 64)
        sean_bwd2_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)dn, (const uint4*)y, norm, normk, part, (uint4*)dy, dbias, HW, ppb, slots, planes());
    else
        sean_bwd2_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)dn, (const uint4*)y, norm, normk, part, (uint4*)dy, dbias, HW, ppb, slots, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_colsum(const void* x, float* out, int64_t rows, int C, void* stream) {
    DASR_REQUIRE(x && out && rows > 0, "bad arguments");
    DASR_REQUIRE(C % 8 == 0 && C >= 8 && C / 8 <= 256, "colsum: unsupported C %d", C);   // 256 % G != 0: spare threads idle
    DASR_REQUIRE(planes() == 1, "colsum has no fp32-split form (bias gradients come from the weight-gradient kernel)");
    const int G = C / 8;
    const int lanes = 256 / G;
    size_t rpb = ((size_t)rows + 4 * (size_t)num_sms() - 1) / (4 * (size_t)num_sms());
    if (rpb < (size_t)lanes * 4) rpb = (size_t)lanes * 4;
    const int grid = (int)(((size_t)rows + rpb - 1) / rpb);
    colsum_kernel<<<grid, 256, (size_t)lanes * (C + 1) * sizeof(float), (cudaStream_t)stream>>>((const uint4*)x, out, (size_t)rows, G, rpb);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_dynconv_bwd(const void* dgb, const uint8_t* labels, const float* masks, const int32_t* flag,
                                float* dT, int B, int K, int H, int W, int nf2, void* stream) {
    DASR_REQUIRE(dgb && dT && (labels || masks), "null pointer");
    DASR_REQUIRE(nf2 <= 128, "2*nf must be <= 128");
    const int rows = 8;
    const size_t smem = (size_t)K * 9 * nf2 * sizeof(float);
    static bool configured[64] = {false};
    int dev = 0;
    DASR_CUDA_OK(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        DASR_CUDA_OK(cudaFuncSetAttribute(dynconv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured[dev & 63] = true;
    }
    DASR_REQUIRE(smem <= 100 * 1024, "table too large");
    const int bands = (H + rows - 1) / rows;
    dynconv_bwd_kernel<<<B * bands, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)dgb, labels, masks, flag, dT, K, H, W, nf2, rows, planes(), (size_t)B * H * W * nf2);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

// nS independent instances in one call (strides between instances: dT BK*N, stp BK*L, Ws / dWs N*L, dstp BK*L)
// parts: 1 = the weight gradient dWs only, 2 = the data gradient dstp only, 3 = both.  dWs only reaches parameter
// gradients (a leaf of the backward: the training step runs it on a side stream), dstp feeds the A_i_j / encoder chain.
extern "C" int dasr_table_bwd_parts(const float* dT, const void* stp, const void* Ws, float* dWs, float* dstp, int nS,
                                    int BK, int N, int L, int parts, void* stream) {
    DASR_REQUIRE(dT && nS > 0 && (parts & 3) && !(parts & ~3), "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (parts & 1) {
        DASR_REQUIRE(stp && dWs, "bad arguments");
        // dWs [N][L] = dT^T [N][BK] * stp [BK][L]
        gemm_f32_bf16_kernel<true, 8><<<dim3((L + 63) / 64, (N + 127) / 128, nS), 256, 0, st>>>(
            dT, (const __nv_bfloat16*)stp, dWs, N, L, BK, N, L, L, BK, 0, 1, (long long)BK * N, (long long)BK * L,
            (long long)N * L, planes(), (size_t)nS * BK * L);
        DASR_LAUNCH_OK();
    }
    if (parts & 2) {
        DASR_REQUIRE(Ws && dstp, "bad arguments");
        // dstp [BK][L] = dT [BK][N] * Ws [N][L]; few output tiles per instance -> split the reduction over N
        const int tiles = ((L + 63) / 64) * ((BK + 63) / 64) * nS;
        // ~8 resident blocks per SM: the kernel has no software pipelining, so it needs the warps to hide its load
        // latency (2 blocks per SM: 147 us for the 26 instances of a training step at B = 16; 8: 110 - 135 us)
        int splits = (8 * num_sms() + tiles - 1) / tiles;
        if (splits > (N + 63) / 64) splits = (N + 63) / 64;
        if (splits < 1) splits = 1;
        const int kps = (((N + splits - 1) / splits) + 15) / 16 * 16;
        splits = (N + kps - 1) / kps;
        DASR_CUDA_OK(cudaMemsetAsync(dstp, 0, (size_t)nS * BK * L * sizeof(float), st));
        gemm_f32_bf16_kernel<false, 4><<<dim3((L + 63) / 64, (BK + 63) / 64, nS * splits), 256, 0, st>>>(
            dT, (const __nv_bfloat16*)Ws, dstp, BK, L, N, N, L, L, kps, 1, splits, (long long)BK * N, (long long)N * L,
            (long long)BK * L, planes(), (size_t)nS * N * L);
        DASR_LAUNCH_OK();
    }
    return DASR_OK;
}

// nS independent instances in one call (strides between instances: dT BK*N, stp BK*L, Ws / dWs N*L, dstp BK*L)
extern "C" int dasr_table_bwd_batched(const float* dT, const void* stp, const void* Ws, float* dWs, float* dstp, int nS,
                                      int BK, int N, int L, void* stream) {
    DASR_REQUIRE(dT && stp && Ws && dWs && dstp && nS > 0, "bad arguments");
    return dasr_table_bwd_parts(dT, stp, Ws, dWs, dstp, nS, BK, N, L, 3, stream);
}

extern "C" int dasr_table_bwd(const float* dT, const void* stp, const void* Ws, float* dWs, float* dstp, int BK, int N,
                              int L, void* stream) {
    return dasr_table_bwd_batched(dT, stp, Ws, dWs, dstp, 1, BK, N, L, stream);
}

extern "C" int dasr_style_mix_bwd(const float* dstp, const float* vec, const float* A, float* dA, float* da, float* dvec,
                                  int B, int K, int L, void* stream) {
    DASR_REQUIRE(dstp && vec && A && dA && da && dvec, "null pointer");
    style_mix_bwd_a_kernel<<<K * K, 256, 0, (cudaStream_t)stream>>>(dstp, vec, dA, da, B, K, L);
    DASR_LAUNCH_OK();
    style_mix_bwd_v_kernel<<<grid_for((size_t)B * K * L, 256), 256, 0, (cudaStream_t)stream>>>(dstp, A, dvec, B, K, L);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_style_mix_bwd_batched(const float* dstp, const float* vec, const void* A_ptrs, const void* dA_ptrs,
                                          const void* da_ptrs, float* dvec, int nS, int B, int K, int L, void* stream) {
    DASR_REQUIRE(dstp && vec && A_ptrs && dA_ptrs && da_ptrs && dvec && nS > 0, "bad arguments");
    style_mix_bwd_a_batched_kernel<<<dim3(K * K, nS), 256, 0, (cudaStream_t)stream>>>(
        dstp, vec, (float* const*)dA_ptrs, (float* const*)da_ptrs, B, K, L);
    DASR_LAUNCH_OK();
    style_mix_bwd_v_batched_kernel<<<grid_for((size_t)B * K * L, 256), 256, 0, (cudaStream_t)stream>>>(
        dstp, (const float* const*)A_ptrs, dvec, nS, B, K, L);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_region_pool_bwd(const float* dvec, const float* msel, const float* cnt, void* de5, int B, int P, int C,
                                    int K, void* stream) {
    DASR_REQUIRE(dvec && msel && cnt && de5, "null pointer");
    region_pool_bwd_kernel<<<grid_for((size_t)B * P * C, 256), 256, 0, (cudaStream_t)stream>>>(dvec, msel, cnt, (__nv_bfloat16*)de5, B, P, C, K, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_actv_bwd(const void* dA, const float* depth, float* dW, float* db, int B, int H, int W, int C,
                             void* stream) {
    DASR_REQUIRE(dA && depth && dW && db && C <= 128, "bad arguments");
    DASR_REQUIRE(planes() == 1, "dasr_actv_bwd has no fp32-split form (use dasr_actv_bwd_tc)");
    const int rows = 8;
    const int bands = (H + rows - 1) / rows;
    actv_bwd_kernel<<<B * bands, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dA, depth, dW, db, H, W, C, rows);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_unshuffle_actgrad(const void* dps, const void* ps_out, void* dconv, int B, int H, int W, int Cq,
                                      float slope, int r, void* stream) {
    DASR_REQUIRE(dps && ps_out && dconv && Cq % 8 == 0, "bad arguments");
    DASR_REQUIRE(r == 2 || r == 3, "PixelShuffle factor must be 2 or 3 (got %d)", r);
    const size_t total = (size_t)B * H * W * r * r * (Cq / 8);
    unshuffle_actgrad_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>((const uint4*)dps, (const uint4*)ps_out, (uint4*)dconv, B, H, W, Cq / 8, slope, r, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_pixel_shuffle(const void* in, void* out, int B, int H, int W, int Cq, int r, void* stream) {
    DASR_REQUIRE(in && out && Cq % 8 == 0 && B > 0 && H > 0 && W > 0, "bad arguments");
    DASR_REQUIRE(r == 2 || r == 3, "PixelShuffle factor must be 2 or 3 (got %d)", r);
    B *= planes();          // fp32-split planes are extra images of a copy kernel
    const size_t total = (size_t)B * H * W * r * r * (Cq / 8);
    pixel_shuffle_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, B, H, W, Cq / 8, r);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_out9_bwd_prep(const float* dout, const float* sr, void* aprime, float* dbias, int B, int H, int W,
                                  void* stream) {
    DASR_REQUIRE(dout && sr && aprime && dbias, "null pointer");
    out9_bwd_prep_kernel<<<grid_for((size_t)B * H * W, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(dout, sr, (__nv_bfloat16*)aprime, dbias, B, H, W, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_actgrad(const void* d, const void* act_out, void* out, int64_t n, float slope, void* stream) {
    DASR_REQUIRE(d && act_out && out && n % 8 == 0, "bad arguments");
    actgrad_kernel<<<grid_for((size_t)n / 8, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>((const uint4*)d, (const uint4*)act_out, (uint4*)out, (size_t)n / 8, slope, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_zero_insert2_to(const void* x, void* out, int B, int H, int W, int C, int Ho, int Wo, void* stream) {
    DASR_REQUIRE(x && out && C % 8 == 0, "bad arguments");
    B *= planes();          // fp32-split planes are extra images of a copy kernel
    zero_insert2_to_kernel<<<grid_for((size_t)B * Ho * Wo * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)out, B, H, W, C / 8, Ho, Wo);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_nchw3_to_nhwc32(const float* x, void* out, int B, int H, int W, void* stream) {
    DASR_REQUIRE(x && out, "null pointer");
    nchw3_to_nhwc32_kernel<<<grid_for((size_t)B * H * W, 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out, B, H, W, planes());
    DASR_LAUNCH_OK();
    return DASR_OK;
}

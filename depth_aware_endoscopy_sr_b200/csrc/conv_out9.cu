// K-OUT9: the 9x9, 32 -> 3 output convolution + clamp of DepthNet (reference
// codes/models/modules/sftmd_arch.py:910,948-950) as a tensor-core kernel that is NOT tap-by-tap.
//
// A plain implicit GEMM would issue 81 taps of N = 3 (padded to 16) -- 5x wasted tensor work and an A halo of
// 8 rows per tile.  Instead the 9 horizontal taps are folded into the GEMM N dimension:
//
//     Z[q][u*3 + co] = sum_{t, ci} X[q + t*64][ci] * W[co][ci][t][u]          (9 MMAs of N = 32 per 128 pixels)
//     out[q][co]     = bias[co] + sum_u Z[q + u][u*3 + co]                     (shifted sum, in the epilogue)
//
// where q = i*64 + j enumerates a zero-padded 22 x 64 pixel patch (14 x 56 outputs + the 4-pixel halo) that ONE
// 4-D TMA box brings into shared memory (64-byte swizzled rows = 32 bf16 channels).  The A operand of vertical
// tap t is the same patch viewed 64 rows further down, so every activation is read from L2 once per tile
// (1.7x halo amplification instead of 81x).  The shifted sum goes through a small fp32 staging tile in shared
// memory (stride 29 words: conflict-free both ways); stores are coalesced NCHW fp32 rows.
//
// Warp roles: w0 TMA producer (double-buffered patches + resident weights), w1 MMA issuer (one thread),
// w2 TMEM allocator, w4..7 epilogue.  TMEM holds two accumulator sets of 7 x 32 columns.
#include "dasr_internal.h"
#include <stdlib.h>
#include "sm100_ptx.cuh"

namespace dasr {

namespace out9 {
constexpr int TH = 14;                 // output rows per tile
constexpr int TW = 56;                 // output cols per tile
constexpr int PW = 64;                 // patch width  (TW + 8)
constexpr int PH = TH + 8;             // patch height (22)
constexpr int NBLK = TH / 2;           // M blocks of 128 = 2 patch rows
constexpr int CIN = 32;
constexpr int NCOL = 32;               // GEMM N (27 used)
constexpr int A_BYTES = PH * PW * 64;  // 90112
constexpr int W_TAP_BYTES = NCOL * 64; // 2048
constexpr int ZS_STRIDE = 29;
constexpr int ZS_ROWS = 136;
constexpr int kThreads = 256;
constexpr size_t SMEM_BYTES = 2 * (size_t)A_BYTES + 9 * W_TAP_BYTES + ZS_ROWS * ZS_STRIDE * 4 + 1024;
// fp32-split planes: one patch buffer, the weights of all planes resident
constexpr size_t SMEM_BYTES_PL = (size_t)A_BYTES + kMaxPlanes * 9 * W_TAP_BYTES + ZS_ROWS * ZS_STRIDE * 4 + 1024;
}  // namespace out9

struct Out9K {
    int B, H, W, Cout;
    int n_strips, n_rows, total_tiles;
    int clamp01;
    const float* bias;
    float* out;
    // frame output (dasr_conv_out9_frames): tensor2img of the clamped value fused into the store -- uint8 BGR HWC
    uint8_t* out_u8;
    float lo, hi;
    // fp32-split planes (dasr_internal.h): cross terms of the x / weight planes, a_stages patch buffers
    int npl, n_terms, a_stages;
    unsigned char ta[6], tb[6];
};

// util.tensor2img (codes/utils/util.py:566-590) of one pixel, fused into the output store: clamp to [lo, hi], scale to
// [0, 255], round half to even, RGB -> BGR.  The same fp32 operations, in the same order, as tensor2img_kernel (io.cu),
// so the fused frames are bit-identical to conv_out9 followed by dasr_tensor2img.
__device__ __forceinline__ void store_frame_pixel(uint8_t* d, float r, float g, float b, float lo, float hi) {
    const float inv = hi - lo;
    const float v[3] = {r, g, b};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float x = fminf(fmaxf(v[c], lo), hi);
        x = __fdiv_rn(__fsub_rn(x, lo), inv);
        d[2 - c] = (uint8_t)__float2int_rn(__fmul_rn(x, 255.0f));
    }
}

// PL: the fp32-split planes form (dasr_set_planes > 1; test infrastructure).  A template parameter so that the product
// kernel keeps compile-time stage counts (with run-time ones it went from 515 to 565 us at batch 64).
template <bool PL>
__global__ void __launch_bounds__(out9::kThreads, 1)
conv_out9_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW, const Out9K p) {
    using namespace out9;
    const int a_stages = PL ? 1 : 2;
    const int n_terms = PL ? p.n_terms : 1;
    const int npl = PL ? p.npl : 1;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[2], a_empty[2], w_full, acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* a_smem = smem;
    uint8_t* w_smem = smem + a_stages * (size_t)A_BYTES;
    float* zs = reinterpret_cast<float*>(w_smem + npl * 9 * W_TAP_BYTES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        mbar_init(&w_full, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
    }
    if (warp == 2) tmem_alloc<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int tiles_per_img = p.n_strips * p.n_rows;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(&w_full, npl * 9 * W_TAP_BYTES);
            for (int t = 0; t < 9 * npl; t++) tma_load_2d(w_smem + t * W_TAP_BYTES, &mapW, &w_full, 0, t * NCOL);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int img = tile / tiles_per_img;
                const int r = tile - img * tiles_per_img;
                const int row = r / p.n_strips, strip = r - row * p.n_strips;
                for (int term = 0; term < n_terms; term++, it++) {
                    const int s = it % a_stages;
                    mbar_wait(&a_empty[s], ((it / a_stages) & 1) ^ 1);
                    mbar_expect_tx(&a_full[s], A_BYTES);
                    tma_load_4d(a_smem + (size_t)s * A_BYTES, &mapA, &a_full[s], 0, strip * TW - 4, row * TH - 4,
                                img + (int)p.ta[term] * p.B);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, NCOL);
            const uint64_t desc_hi = make_smem_desc<64>(0, 0) & ~0x3FFFull;
            const uint32_t w_base = smem_u32(w_smem);
            mbar_wait(&w_full, 0);
            tc_fence_after();
            uint32_t it = 0, ia = 0;
            const bool pl = PL;      // fp32-split planes: one tile in flight; accumulator set 1 = low-order terms
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it++) {
                const int s = pl ? 0 : (it & 1);
                const uint32_t ph = pl ? (it & 1) : ((it >> 1) & 1);
                mbar_wait(&acc_empty[s], ph ^ 1);
                const uint32_t d_base0 = tmem_base + s * 256;
                for (int term = 0; term < n_terms; term++, ia++) {
                    const int sa = ia % a_stages;
                    mbar_wait(&a_full[sa], (ia / a_stages) & 1);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(a_smem + (size_t)sa * A_BYTES);
                    const uint32_t wt_base = w_base + (uint32_t)p.tb[term] * 9 * W_TAP_BYTES;
                    const uint32_t d_base = d_base0 + (term > 0 ? 256 : 0);
                    const int term_rel = term > 1 ? 1 : 0;          // first term of its accumulator set: overwrite
#pragma unroll 1
                    for (int blk = 0; blk < NBLK; blk++) {
#pragma unroll
                        for (int t = 0; t < 9; t++) {
#pragma unroll
                            for (int k = 0; k < 2; k++) {
                                const uint32_t aa = a_base + (uint32_t)(blk * 128 + t * PW) * 64 + k * 32;
                                const uint32_t bb = wt_base + t * W_TAP_BYTES + k * 32;
                                umma_bf16(d_base + blk * NCOL, desc_hi | ((aa & 0x3FFFFu) >> 4),
                                          desc_hi | ((bb & 0x3FFFFu) >> 4), idesc, (term_rel | t | k) != 0);
                            }
                        }
                    }
                    umma_commit(&a_empty[sa]);
                }
                umma_commit(&acc_full[s]);
            }
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;
        const int m = ew * 32 + lane;
        const float b0 = __ldg(p.bias), b1 = p.Cout > 1 ? __ldg(p.bias + 1) : 0.f, b2 = p.Cout > 2 ? __ldg(p.bias + 2) : 0.f;
        const size_t plane = (size_t)p.H * p.W;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it++) {
            const int img = tile / tiles_per_img;
            const int r = tile - img * tiles_per_img;
            const int row = r / p.n_strips, strip = r - row * p.n_strips;
            const bool pl = PL;
            const int s = pl ? 0 : (it & 1);
            mbar_wait(&acc_full[s], pl ? (it & 1) : ((it >> 1) & 1));
            tc_fence_after();
            const uint32_t t_acc = tmem_base + s * 256 + (uint32_t(ew * 32) << 16);
            const int jj = m & 63;
            const int w = strip * TW + jj;
#pragma unroll 1
            for (int blk = 0; blk < NBLK; blk++) {
                uint32_t v[32];
                tmem_ld32(t_acc + blk * NCOL, v);
                tmem_ld_wait();
                if (pl) {          // add the low-order accumulator set (fp32 round-to-nearest)
                    uint32_t w[32];
                    tmem_ld32(t_acc + 256 + blk * NCOL, w);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; c++) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(w[c]));
                }
                asm volatile("bar.sync 1, 128;\n" ::: "memory");   // previous block's readers are done
#pragma unroll
                for (int c = 0; c < 27; c++) zs[m * ZS_STRIDE + c] = __uint_as_float(v[c]);
                asm volatile("bar.sync 1, 128;\n" ::: "memory");
                const int h = row * TH + blk * 2 + (m >> 6);
                if (jj < TW && w < p.W && h < p.H) {
                    float o0 = b0, o1 = b1, o2 = b2;
#pragma unroll
                    for (int u = 0; u < 9; u++) {
                        const float* zp = zs + (m + u) * ZS_STRIDE + u * 3;
                        o0 += zp[0];
                        o1 += zp[1];
                        o2 += zp[2];
                    }
                    if (p.out_u8) {
                        store_frame_pixel(p.out_u8 + ((size_t)img * plane + (size_t)h * p.W + w) * 3, o0, o1, o2, p.lo, p.hi);
                    } else {
                    if (p.clamp01) {
                        o0 = fminf(fmaxf(o0, 0.f), 1.f);
                        o1 = fminf(fmaxf(o1, 0.f), 1.f);
                        o2 = fminf(fmaxf(o2, 0.f), 1.f);
                    }
                    float* op = p.out + (size_t)img * p.Cout * plane + (size_t)h * p.W + w;
                    op[0] = o0;
                    if (p.Cout > 1) op[plane] = o1;
                    if (p.Cout > 2) op[2 * plane] = o2;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[s]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem_base);
}


// ------------------------------------------------------------------------------------------------ CTA pairs
// The same convolution on CTA pairs (tcgen05.mma.cta_group::2): N = 32 makes a single-CTA M128 MMA cost 40 tensor
// cycles for 16 cycles of math (the A operand's shared-memory reads bound it), and the kernel is MMA-bound: 126
// dispatches per 14 x 56 tile, 160 tiles per SM -> ~420 us against a 196 us HBM floor at batch 64.  A pair issues
// M256 x N32 in the same time.  The two CTAs of a pair work on the same tile position of two consecutive images; the
// 9 weight tiles stay resident, each CTA holding rows rank*16..+15 of every tile.
// Eight epilogue warps in two groups (even / odd M blocks, each with its own staging tile and named barrier): with the
// MMA time halved the shifted-sum epilogue of four warps became the critical path.
constexpr int kOut9PairThreads = 384;
constexpr size_t kOut9PairSmem = 2 * (size_t)out9::A_BYTES + 9 * 1024 + 2 * out9::ZS_ROWS * out9::ZS_STRIDE * 4 + 1024;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kOut9PairThreads, 1)
conv_out9_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW, const Out9K p) {
    using namespace out9;
    constexpr uint32_t W_HALF_BYTES = (NCOL / 2) * 64;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[2], a_empty[2], w_full, acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* a_smem = smem;
    uint8_t* w_smem = smem + 2 * (size_t)A_BYTES;
    float* zs = reinterpret_cast<float*>(w_smem + 9 * 1024);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 16);         // 8 epilogue warps x 2 CTAs (leader only)
        }
        mbar_init(&w_full, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
    }
    if (warp == 2) tmem_alloc_pair<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int tiles_per_img = p.n_strips * p.n_rows;
    const int n_pairs = gridDim.x >> 1;
    const int pair_id = blockIdx.x >> 1;

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t w_full_l = mapa_u32(smem_u32(&w_full), 0);
            if (leader) mbar_expect_tx(&w_full, 2 * 9 * W_HALF_BYTES);
            for (int t = 0; t < 9; t++)
                tma_load_2d_pair(w_smem + t * 1024, &mapW, w_full_l, 0, t * NCOL + (int)rank * (NCOL / 2));
            uint32_t it = 0;
            for (int tile = pair_id; tile < p.total_tiles; tile += n_pairs, it++) {
                const int ip = tile / tiles_per_img;
                const int r = tile - ip * tiles_per_img;
                const int row = r / p.n_strips, strip = r - row * p.n_strips;
                const int s = it & 1;
                mbar_wait(&a_empty[s], ((it >> 1) & 1) ^ 1);
                const uint32_t full_l = mapa_u32(smem_u32(&a_full[s]), 0);
                if (leader) mbar_expect_tx(&a_full[s], 2 * A_BYTES);
                tma_load_4d_pair(a_smem + (size_t)s * A_BYTES, &mapA, full_l, 0, strip * TW - 4, row * TH - 4,
                                 2 * ip + (int)rank);          // image >= B (odd batch): zero-filled
            }
        }
    } else if (warp == 1 && leader) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(256, NCOL);
            const uint64_t desc0 = make_smem_desc<64>(0, 0);
            const uint32_t desc_hi = (uint32_t)(desc0 >> 32);
            const uint32_t lo_flags = (uint32_t)desc0 & ~0x3FFFu;
            const uint32_t w_lo0 = lo_flags | ((smem_u32(w_smem) & 0x3FFFFu) >> 4);
            mbar_wait(&w_full, 0);
            tc_fence_after();
            uint32_t it = 0;
            for (int tile = pair_id; tile < p.total_tiles; tile += n_pairs, it++) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&acc_empty[s], ph ^ 1);
                mbar_wait(&a_full[s], ph);
                tc_fence_after();
                const uint32_t a_lo0 = lo_flags | ((smem_u32(a_smem + (size_t)s * A_BYTES) & 0x3FFFFu) >> 4);
                const uint32_t d_base = tmem_base + s * 256;
#pragma unroll 1
                for (int blk = 0; blk < NBLK; blk++) {
#pragma unroll
                    for (int t = 0; t < 9; t++) {
#pragma unroll
                        for (int k = 0; k < 2; k++)
                            umma_bf16_lohi_pair(d_base + blk * NCOL, a_lo0 + (((uint32_t)(blk * 128 + t * PW) * 64 + k * 32) >> 4),
                                                w_lo0 + ((t * 1024 + k * 32) >> 4), desc_hi, idesc, (t | k) != 0);
                    }
                }
                umma_commit_pair(&a_empty[s]);
                umma_commit_pair(&acc_full[s]);
            }
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;
        const int grp = (warp - 4) >> 2;              // 0: even M blocks, 1: odd M blocks
        float* zsg = zs + grp * (ZS_ROWS * ZS_STRIDE);
        const int m = ew * 32 + lane;
        const float b0 = __ldg(p.bias), b1 = p.Cout > 1 ? __ldg(p.bias + 1) : 0.f, b2 = p.Cout > 2 ? __ldg(p.bias + 2) : 0.f;
        const size_t plane = (size_t)p.H * p.W;
        uint32_t it = 0;
        for (int tile = pair_id; tile < p.total_tiles; tile += n_pairs, it++) {
            const int ip = tile / tiles_per_img;
            const int r = tile - ip * tiles_per_img;
            const int row = r / p.n_strips, strip = r - row * p.n_strips;
            const int img = 2 * ip + (int)rank;
            const bool have = img < p.B;
            const int s = it & 1;
            mbar_wait(&acc_full[s], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + s * 256 + (uint32_t(ew * 32) << 16);
            const int jj = m & 63;
            const int w = strip * TW + jj;
#pragma unroll 1
            for (int blk = grp; blk < (have ? NBLK : 0); blk += 2) {
                uint32_t v[32];
                tmem_ld32(t_acc + blk * NCOL, v);
                tmem_ld_wait();
                asm volatile("bar.sync %0, 128;\n" ::"r"(1 + grp) : "memory");   // previous block's readers are done
#pragma unroll
                for (int c = 0; c < 27; c++) zsg[m * ZS_STRIDE + c] = __uint_as_float(v[c]);
                asm volatile("bar.sync %0, 128;\n" ::"r"(1 + grp) : "memory");
                const int h = row * TH + blk * 2 + (m >> 6);
                if (jj < TW && w < p.W && h < p.H) {
                    float o0 = b0, o1 = b1, o2 = b2;
#pragma unroll
                    for (int u = 0; u < 9; u++) {
                        const float* zp = zsg + (m + u) * ZS_STRIDE + u * 3;
                        o0 += zp[0];
                        o1 += zp[1];
                        o2 += zp[2];
                    }
                    if (p.out_u8) {
                        store_frame_pixel(p.out_u8 + ((size_t)img * plane + (size_t)h * p.W + w) * 3, o0, o1, o2, p.lo, p.hi);
                    } else {
                    if (p.clamp01) {
                        o0 = fminf(fmaxf(o0, 0.f), 1.f);
                        o1 = fminf(fmaxf(o1, 0.f), 1.f);
                        o2 = fminf(fmaxf(o2, 0.f), 1.f);
                    }
                    float* op = p.out + (size_t)img * p.Cout * plane + (size_t)h * p.W + w;
                    op[0] = o0;
                    if (p.Cout > 1) op[plane] = o1;
                    if (p.Cout > 2) op[2 * plane] = o2;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[s]), 0));
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace dasr

using namespace dasr;

static int conv_out9_launch(const void* x, const void* wq, const float* bias, float* out, uint8_t* out_u8, float lo,
                            float hi, int B, int H, int W, int Cout, int clamp01, void* stream_) {
    using namespace out9;
    cudaStream_t stream = (cudaStream_t)stream_;
    DASR_REQUIRE(x && wq && bias && (out || out_u8), "null tensor pointer");
    DASR_REQUIRE(B > 0 && H > 0 && W > 0 && Cout == 3, "bad shape (the output conv has Cout == 3)");
    Out9K k;
    k.B = B; k.H = H; k.W = W; k.Cout = Cout;
    k.n_strips = (W + TW - 1) / TW;
    k.n_rows = (H + TH - 1) / TH;
    k.total_tiles = B * k.n_strips * k.n_rows;
    k.clamp01 = clamp01;
    k.bias = bias;
    k.out = out;
    k.out_u8 = out_u8;
    k.lo = lo;
    k.hi = hi;
    k.npl = planes();
    k.a_stages = k.npl > 1 ? 1 : 2;
    {
        const PlaneTerms t = plane_terms(k.npl, k.npl);
        k.n_terms = t.n;
        for (int i = 0; i < t.n; i++) { k.ta[i] = t.a[i]; k.tb[i] = t.b[i]; }
    }
    CUtensorMap mA, mW;
    {
        uint64_t dims[4] = {(uint64_t)CIN, (uint64_t)W, (uint64_t)H, (uint64_t)B * k.npl};
        uint64_t str[3] = {(uint64_t)CIN * 2, (uint64_t)W * CIN * 2, (uint64_t)H * W * CIN * 2};
        uint32_t box[4] = {(uint32_t)CIN, (uint32_t)PW, (uint32_t)PH, 1};
        int rc = encode_tmap_bf16(&mA, x, 4, dims, str, box, 64);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)CIN, (uint64_t)9 * NCOL * k.npl};
        uint64_t str[1] = {(uint64_t)CIN * 2};
        uint32_t box[2] = {(uint32_t)CIN, (uint32_t)NCOL};
        int rc = encode_tmap_bf16(&mW, wq, 2, dims, str, box, 64);
        if (rc) return rc;
    }
    static bool configured[64] = {false};
    int dev = 0;
    DASR_CUDA_OK(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        DASR_CUDA_OK(cudaFuncSetAttribute(conv_out9_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        DASR_CUDA_OK(cudaFuncSetAttribute(conv_out9_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        configured[dev & 63] = true;
    }
    // CTA pairs (two images per pair) when the batch has >= 2 images and every SM pair gets work
    static int pair_env = -1;
    if (pair_env < 0) {
        const char* e = getenv("DASR_OUT9_PAIR");
        pair_env = (e && e[0] == '0') ? 0 : 1;
    }
    const bool pairs_ok = pair_env && pair_kernels_enabled(DASR_PAIR_OUT9);
    const int pair_tiles = ((B + 1) / 2) * k.n_strips * k.n_rows;
    if (pairs_ok && k.npl == 1 && B >= 2 && pair_tiles >= num_sms() / 2) {
        CUtensorMap mWh;
        uint64_t dims[2] = {(uint64_t)CIN, (uint64_t)9 * NCOL};
        uint64_t str[1] = {(uint64_t)CIN * 2};
        uint32_t box[2] = {(uint32_t)CIN, (uint32_t)(NCOL / 2)};
        int rc = encode_tmap_bf16(&mWh, wq, 2, dims, str, box, 64);
        if (rc) return rc;
        static bool configured2[64] = {false};
        if (!configured2[dev & 63]) {
            DASR_CUDA_OK(cudaFuncSetAttribute(conv_out9_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOut9PairSmem));
            configured2[dev & 63] = true;
        }
        Out9K kp = k;
        kp.total_tiles = pair_tiles;
        int n_pairs = num_sms() / 2;
        if (n_pairs > pair_tiles) n_pairs = pair_tiles;
        conv_out9_pair_kernel<<<2 * n_pairs, kOut9PairThreads, kOut9PairSmem, stream>>>(mA, mWh, kp);
        DASR_LAUNCH_OK();
        return DASR_OK;
    }
    const int grid = k.total_tiles < num_sms() ? k.total_tiles : num_sms();
    if (k.npl > 1) conv_out9_kernel<true><<<grid, kThreads, SMEM_BYTES_PL, stream>>>(mA, mW, k);
    else conv_out9_kernel<false><<<grid, kThreads, SMEM_BYTES, stream>>>(mA, mW, k);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_conv_out9(const void* x, const void* wq, const float* bias, float* out, int B, int H, int W,
                              int Cout, int clamp01, void* stream) {
    DASR_REQUIRE(out, "null tensor pointer");
    return conv_out9_launch(x, wq, bias, out, nullptr, 0.f, 1.f, B, H, W, Cout, clamp01, stream);
}

extern "C" int dasr_conv_out9_frames(const void* x, const void* wq, const float* bias, uint8_t* img, int B, int H, int W,
                                     float lo, float hi, void* stream) {
    DASR_REQUIRE(img && hi > lo, "bad arguments");
    return conv_out9_launch(x, wq, bias, nullptr, img, lo, hi, B, H, W, 3, 0, stream);
}

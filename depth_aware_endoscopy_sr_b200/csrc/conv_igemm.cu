// K-CONV: stride-1 "same" convolution as an implicit GEMM on the 5th-gen tensor cores.
//
//   D[pixel][cout] = sum_{tap,cin} A[pixel + tap][cin] * W[cout][tap][cin]      (bf16 x bf16 -> fp32)
//
// Design (B200-first, not a cuDNN translation):
//  * HALO REUSE. One 4-D TMA box load brings a zero-padded (RB x Wp) pixel patch of KC channels into
//    shared memory ONCE (128B/64B-swizzled rows, one row per pixel).  Output pixels are enumerated in
//    "padded-flat" order q = h*Wp + w, so the A operand of tap (t,u) is the SAME smem tile viewed through
//    a descriptor whose start address is shifted by (t*Wp+u) rows -- measured to be legal on sm_100a with
//    base_offset = 0 (profiles/r01_umma_probe.md).  A 3x3 conv therefore reads each activation from L2
//    ~1.3x instead of 9x (81x for the 9x9 output conv); the tensor pipe is no longer L2-bound at N=64.
//    The price is (Wp-Wt)/Wp garbage rows (3% at W=64) that the epilogue discards.
//  * Weights stay resident in shared memory for the whole persistent CTA when they fit (trunk 64->64:
//    72 KB), otherwise they stream through a TMA ring (gamma/beta conv: 288 KB).
//  * Warp roles: w0 = A producer (TMA), w3 = B producer (TMA), w1 = MMA issuer (one elected thread,
//    tcgen05.mma cta_group::1, M=128), w2 = TMEM allocator, w4..7 = epilogue (tcgen05.ld -> fused
//    epilogue -> vectorised NHWC bf16 stores).  Accumulators are double-buffered in TMEM so the epilogue
//    of tile i overlaps the MMAs of tile i+1.
//  * Fused epilogues: bias/activation/residual, InstanceNorm statistics (warp butterfly + fp32 atomics),
//    the whole SEAN modulation (normalise, (1+gamma), beta, ReLU, residual), PixelShuffle(2) as store
//    addressing, and the final clamp to NCHW fp32.
//
// Replaces: nn.Conv2d call sites codes/models/modules/sftmd_arch.py:743-749,812,819,862-864,891-910 and
// codes/models/modules/normalization.py:41-42,87-89 of the reference.
#include "dasr_internal.h"
#include <stdlib.h>
#include <string.h>
#include "sm100_ptx.cuh"

namespace dasr {

// Optional in-kernel stall accounting (build with -DDASR_PROFILE; tools/prof_stalls.py reads it back).
// slots: 0 mma:wait acc_empty  1 mma:wait a_full  2 mma:wait b_full  3 mma:issue  4 epi:wait acc_full
//        5 epi:work            6 a-producer:wait a_empty  7 b-producer:wait b_empty  8 kernel total (CTA 0..)
#ifdef DASR_PROFILE
// ablation knob of the profile build (tools/prof_stalls.py, DASR_DBG): bit 0 = the epilogues skip their global
// loads, bit 1 = they skip their global stores (ConvK::dbg, set by dasr_prof_set)
#define DBG(p, bit) ((p).dbg & (bit))
static int g_host_dbg = 0;
__device__ unsigned long long g_prof[16];
#define PROF_DECL long long prof_t = clock64(); long long prof_acc[4] = {0, 0, 0, 0}
#define PROF_LAP(i) do { long long n_ = clock64(); prof_acc[i] += n_ - prof_t; prof_t = n_; } while (0)
#define PROF_FLUSH(base, n) do { for (int i_ = 0; i_ < (n); i_++) atomicAdd(&g_prof[(base) + i_], (unsigned long long)prof_acc[i_]); } while (0)
// whole-kernel SM cycles (slot 8) and nanoseconds (slot 9) of every CTA: their ratio is the SM clock DURING the kernel
__device__ __forceinline__ unsigned long long prof_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define PROF_KERNEL_BEGIN long long pk_c0 = clock64(); unsigned long long pk_n0 = prof_ns()
#define PROF_KERNEL_END do { if (threadIdx.x == 0) { atomicAdd(&g_prof[8], (unsigned long long)(clock64() - pk_c0)); atomicAdd(&g_prof[9], prof_ns() - pk_n0); } } while (0)
#else
#define PROF_KERNEL_BEGIN
#define PROF_KERNEL_END
#define DBG(p, bit) 0
#define PROF_DECL
#define PROF_LAP(i)
#define PROF_FLUSH(base, n)
#endif

#ifdef DASR_CONV_PRECISE_TU
constexpr bool kPrecTU = true;      // this translation unit instantiates the fp32-split (PREC) kernels
#else
constexpr bool kPrecTU = false;
#endif
constexpr int kMaxBStages = 96;
constexpr int kMaxAStages = 4;
constexpr int kThreads = 384;   // 4 role warps + 8 epilogue warps
constexpr int kMaxBias = 1152;

struct ConvK {
    // geometry
    int B, H, W, Cin, Cout;
    int kh, kw, pad_h, pad_w, taps;
    int Wt, Wp, RB;
    unsigned wp_magic;             // floor(2^32 / Wp) + 1: q / Wp == __umulhi(q, wp_magic)
    int n_strips, tiles_per_strip, ntn, total_tiles;
    int nch;                       // Cin / KC
    int SA, SB, b_resident;
    uint32_t a_stage_bytes, b_stage_bytes, a_tx_bytes, b_tx_bytes;
    // epilogue
    int epi, act, subsample, clamp01, inner_relu;
    int Ho, Wo;                    // output spatial size (after subsample / shuffle)
    const float* bias;
    __nv_bfloat16* out;
    float* out_f32;
    const __nv_bfloat16* resid;
    float* stats;
    const __nv_bfloat16* y;
    const float* norm;
    const __nv_bfloat16* gb_s;
    const __nv_bfloat16* actmask;  // STORE: multiply by (actmask > 0 ? 1 : mask_slope) -- activation backward
    float mask_slope;
    __nv_bfloat16* gamma_out;      // SEAN: optional copy of gamma (bf16 [B,H,W,nf]) for the backward pass
    const float* resid_f32;        // SEAN: fp32 residual stream (takes precedence over resid)
    float* out_aux_f32;            // SEAN: fp32 copy of the output (the residual stream of the next block)
    int nslots;                    // STATS: partial-sum slots per image
    int n_bias;                    // padded Cout (bias entries staged in shared memory)
    int dbg;                       // profile build only (see DBG)
    float* norm_out;               // SEAN with fused finalize: optional copies of (mean, scale) [B][nf][2] ...
    float* normk_out;              // ... and of k [B][nf] for the backward pass
    // K-DYN folded into the GEMM (SEAN epilogue only): after the main K loop, 9 more taps of K = 16 over the
    // mask image `aux16` (NHWC bf16 [B,H,W,16], 32-byte swizzled rows) with the PER-IMAGE filters wdyn
    // ([B*Cout][taps*16]) accumulate the dynamic convolution gb_s straight into the same TMEM accumulators.
    int dyn;
    uint32_t a2_tx_bytes, b2_tx_bytes, a2_off;
    // GEN kernels (SEAN epilogue, inference): the A operand actv = ReLU(conv3x3(depth, 1 -> Cin) + b) is GENERATED
    // inside the kernel by a fourth warpgroup instead of being read from memory (normalization.py:37-40,61)
    const float* gen_depth;        // [B,1,H,W] fp32
    const float* gen_w;            // mlp_mask weight [Cin][9] fp32
    const float* gen_b;            // mlp_mask bias [Cin] fp32
    uint32_t gen_off;              // depth-halo scratch inside the dynamic shared memory
    int w_img_rows;                // > 0: per-image weights, image b uses rows [b*w_img_rows, +Cout) of the B matrix
    int unshuffle;                 // EPI_STORE: space-to-depth store addressing (PixelShuffle(2) backward)
    // fp32-split planes (PREC kernels, dasr_set_planes > 1; dasr_internal.h): the K loop runs the cross terms
    // (ta[i], tb[i]) of the operand planes, the epilogues read plane sums and write plane splits
    int npl, n_terms;
    unsigned char ta[6], tb[6];
    int w_plane_rows;              // rows of ONE plane of the packed weight matrix
    int dynw_plane_rows;           // rows of one plane of the dynamic filters (B * N_TILE)
    size_t ps_out;                 // plane stride (elements) of out / resid / y / gamma_out
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == DASR_ACT_RELU) return fmaxf(v, 0.f);
    if (act == DASR_ACT_LRELU) return v > 0.f ? v : 0.2f * v;
    return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256).  The epilogues touch 32-byte pieces of 64..256-byte pixel
// rows, one row per lane, so every warp-level access spans 32 cache lines: halving the instruction count halves the
// LSU wavefronts, which is what bounds the fused epilogues (ncu: lg_throttle).  Addresses are 32-byte aligned.
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
    // L1::no_allocate on the epilogue's global accesses: every byte is touched once, and the L1 shares its data path
    // with the shared-memory operands of the MMAs (B = 64 step 5.63 -> 5.48 ms, SEAN conv 89.8 -> 85.5 us, trunk conv
    // 33.2 -> 30.6 us, training step 6.37 -> 6.26 ms; -DDASR_EPI_L1_ALLOC restores the default policy)
#ifndef DASR_EPI_L1_ALLOC
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
#ifndef DASR_EPI_L1_ALLOC
    asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
#else
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
#endif
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
__device__ __forceinline__ void stg256f(float* p, const float* f) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(__float_as_uint(f[0])),
                 "r"(__float_as_uint(f[1])), "r"(__float_as_uint(f[2])), "r"(__float_as_uint(f[3])),
                 "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])), "r"(__float_as_uint(f[6])),
                 "r"(__float_as_uint(f[7]))
                 : "memory");
}
__device__ __forceinline__ void load16(const __nv_bfloat16* p, float* f) {
    uint4 a, b;
    ldg256(p, a, b);
    unpack8(a, f);
    unpack8(b, f + 8);
}
__device__ __forceinline__ void store16(__nv_bfloat16* p, const float* f) {
    stg256(p, pack8(f), pack8(f + 8));
}

// plane-aware forms of load16 / store16 (PREC kernels): sum of the planes / split into planes; ps = plane stride
template <bool PREC>
__device__ __forceinline__ void add_planes16(const __nv_bfloat16* p, size_t ps, int npl, float* f) {
    if (PREC) {
        for (int k = 1; k < npl; k++) {
            float t[16];
            load16(p + (size_t)k * ps, t);
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] += t[j];
        }
    }
}
template <bool PREC>
__device__ __forceinline__ void store16p(__nv_bfloat16* p, size_t ps, int npl, const float* f) {
    const uint4 a = pack8(f), b = pack8(f + 8);
    stg256(p, a, b);
    if (PREC) {
        float r[16], t[16];
        uint4 ua = a, ub = b;
#pragma unroll
        for (int j = 0; j < 16; j++) r[j] = f[j];
        for (int k = 1; k < npl; k++) {
            unpack8(ua, t);
            unpack8(ub, t + 8);
#pragma unroll
            for (int j = 0; j < 16; j++) r[j] -= t[j];
            ua = pack8(r);
            ub = pack8(r + 8);
            stg256(p + (size_t)k * ps, ua, ub);
        }
    }
}

// TMEM read-out of 16 accumulator columns.  PREC kernels keep TWO accumulator sets per tile, ACC_COLS columns apart:
// set 0 = the hi x hi plane term, set 1 = every lower-order cross term (2^-8 .. 2^-16 of the magnitude).  tcgen05
// accumulates with truncation, ~1 ulp of the ACCUMULATOR per MMA step: keeping the small terms out of the large sum
// leaves K/16 truncations of the main term instead of 6 K/16 (measured: 1e-5 -> 2e-6 relative on a 3x3 conv); the two
// sets are added here in fp32 round-to-nearest.  (issues the loads AND waits for them)
template <bool PREC, int ACC_COLS>
__device__ __forceinline__ void tmem_ld16_acc(uint32_t taddr, uint32_t (&v)[16]) {
    tmem_ld16(taddr, v);
    if (PREC) {
        uint32_t w[16];
        tmem_ld16(taddr + ACC_COLS, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
    } else {
        tmem_ld_wait();
    }
}

// Reduce 16 per-thread values over the 32 lanes of a warp (recursive halving).  On return lane L holds in
// v[0] the warp-wide sum of column  8*b4 + 4*b3 + 2*b2 + b1  (bN = bit N of L); lanes L and L^1 agree.
__device__ __forceinline__ void warp_colsum16(float* v, int lane) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        bool hi = lane & 16;
        float send = hi ? v[i] : v[i + 8];
        float keep = hi ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        bool hi = lane & 8;
        float send = hi ? v[i] : v[i + 4];
        float keep = hi ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        bool hi = lane & 4;
        float send = hi ? v[i] : v[i + 2];
        float keep = hi ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        bool hi = lane & 2;
        float send = hi ? v[0] : v[1];
        float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Fused SEAN epilogue of one tile for one thread (accumulator row m of every M block, the 16-column chunks of
// its warp half):  out = act( relu?( IN(IN(y)) * (1 + gamma) + beta ) + resid ),  gamma/beta = acc + bias + gb_s.
// The HBM operands of step i+1 are requested before step i is computed (register double buffering), so
// their latency hides behind the TMEM load and the arithmetic of the previous step.
struct SeanOps {
    uint4 y0, y1, g0, g1, b0, b1, r0, r1;
    float4 rf[4];
    bool valid;
    size_t pix;
};

template <int N_TILE, int NB>
__device__ __forceinline__ void sean_load(const ConvK& p, SeanOps& o, int img, int q0, int w0, int m, int blk, int c0) {
    constexpr int NF = N_TILE / 2;
    const int q = q0 + blk * 128 + m;
    const int h = (int)__umulhi((unsigned)q, p.wp_magic);
    const int wl = q - h * p.Wp;
    const int w = w0 + wl;
    o.valid = (h < p.H) && (wl < p.Wt) && (w < p.W);
    o.pix = ((size_t)img * p.H + h) * p.W + w;
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    o.y0 = o.y1 = o.g0 = o.g1 = o.b0 = o.b1 = o.r0 = o.r1 = z4;
    if (!o.valid || DBG(p, 1)) return;
    ldg256(p.y + o.pix * NF + c0, o.y0, o.y1);
    if (p.gb_s) {
        const __nv_bfloat16* sp = p.gb_s + o.pix * N_TILE + c0;
        ldg256(sp, o.g0, o.g1);
        ldg256(sp + NF, o.b0, o.b1);
    }
    if (p.resid_f32) {
        const float* rp = p.resid_f32 + o.pix * NF + c0;
        uint4 a, b, c, d;
        ldg256(rp, a, b);
        ldg256(rp + 8, c, d);
        o.rf[0] = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
        o.rf[1] = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w));
        o.rf[2] = make_float4(__uint_as_float(c.x), __uint_as_float(c.y), __uint_as_float(c.z), __uint_as_float(c.w));
        o.rf[3] = make_float4(__uint_as_float(d.x), __uint_as_float(d.y), __uint_as_float(d.z), __uint_as_float(d.w));
    } else if (p.resid) {
        ldg256(p.resid + o.pix * NF + c0, o.r0, o.r1);
    }
}

template <int N_TILE, int NB, bool PREC>
__device__ __forceinline__ void sean_epilogue(const ConvK& p, uint32_t t_acc, const float* bias_t, const float* norm_s,
                                              int img, int q0, int w0, int m, int half) {
    constexpr int NF = N_TILE / 2;
    constexpr int CH = (NF + 31) / 32;     // 16-column chunks per warp half and M block
    constexpr int STEPS = NB * CH;
    if (half * 16 >= NF) return;
    SeanOps ops[2];
    sean_load<N_TILE, NB>(p, ops[0], img, q0, w0, m, 0, half * 16);
#pragma unroll
    for (int it = 0; it < STEPS; it++) {
        const int blk = it / CH, c0 = half * 16 + (it % CH) * 32;
        if (it + 1 < STEPS)
            sean_load<N_TILE, NB>(p, ops[(it + 1) & 1], img, q0, w0, m, (it + 1) / CH, half * 16 + ((it + 1) % CH) * 32);
        const SeanOps& o = ops[it & 1];
        uint32_t vg[16], vb[16];
        if (PREC) {
            tmem_ld16_acc<true, NB * N_TILE>(t_acc + blk * N_TILE + c0, vg);
            tmem_ld16_acc<true, NB * N_TILE>(t_acc + blk * N_TILE + NF + c0, vb);
        } else {
            tmem_ld16(t_acc + blk * N_TILE + c0, vg);
            tmem_ld16(t_acc + blk * N_TILE + NF + c0, vb);
            tmem_ld_wait();
        }
        if (!o.valid) continue;
        float yv[16], gs[16], f[16];
        unpack8(o.y0, yv);
        unpack8(o.y1, yv + 8);
        add_planes16<PREC>(p.y + o.pix * NF + c0, p.ps_out, p.npl, yv);
        // bias / (mean, scale) of the 16 columns with 128-bit shared-memory loads (a scalar LDS per value made this
        // loop ~530 instructions per step); run-time flags are tested once per step, not once per element
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
            const float4 bg = *reinterpret_cast<const float4*>(bias_t + c0 + j4);
            const float4 bb = *reinterpret_cast<const float4*>(bias_t + NF + c0 + j4);
            gs[j4] = __uint_as_float(vg[j4]) + bg.x;
            gs[j4 + 1] = __uint_as_float(vg[j4 + 1]) + bg.y;
            gs[j4 + 2] = __uint_as_float(vg[j4 + 2]) + bg.z;
            gs[j4 + 3] = __uint_as_float(vg[j4 + 3]) + bg.w;
            f[j4] = __uint_as_float(vb[j4]) + bb.x;
            f[j4 + 1] = __uint_as_float(vb[j4 + 1]) + bb.y;
            f[j4 + 2] = __uint_as_float(vb[j4 + 2]) + bb.z;
            f[j4 + 3] = __uint_as_float(vb[j4 + 3]) + bb.w;
        }
        if (p.gb_s) {      // dynamic-conv term from memory (when it is not folded into the GEMM)
            float t0[16];
            unpack8(o.g0, t0);
            unpack8(o.g1, t0 + 8);
#pragma unroll
            for (int j = 0; j < 16; j++) gs[j] += t0[j];
            unpack8(o.b0, t0);
            unpack8(o.b1, t0 + 8);
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] += t0[j];
        }
#pragma unroll
        for (int j2 = 0; j2 < 16; j2 += 2) {
            const float4 nm = *reinterpret_cast<const float4*>(norm_s + 2 * (c0 + j2));     // (mean, scale) x 2
            f[j2] = fmaf((yv[j2] - nm.x) * nm.y, 1.f + gs[j2], f[j2]);
            f[j2 + 1] = fmaf((yv[j2 + 1] - nm.z) * nm.w, 1.f + gs[j2 + 1], f[j2 + 1]);
        }
        if (p.inner_relu) {
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = fmaxf(f[j], 0.f);
        }
        if (DBG(p, 2)) continue;
        if (p.gamma_out) store16p<PREC>(p.gamma_out + o.pix * NF + c0, p.ps_out, p.npl, gs);
        if (p.resid_f32) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                f[4 * j] += o.rf[j].x;
                f[4 * j + 1] += o.rf[j].y;
                f[4 * j + 2] += o.rf[j].z;
                f[4 * j + 3] += o.rf[j].w;
            }
        } else if (p.resid) {
            float rr[16];
            unpack8(o.r0, rr);
            unpack8(o.r1, rr + 8);
            add_planes16<PREC>(p.resid + o.pix * NF + c0, p.ps_out, p.npl, rr);
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] += rr[j];
        }
        if (p.act == DASR_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = fmaxf(f[j], 0.f);
        } else if (p.act == DASR_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = f[j] > 0.f ? f[j] : 0.2f * f[j];
        }
        store16p<PREC>(p.out + o.pix * NF + c0, p.ps_out, p.npl, f);
        if (p.out_aux_f32) {
            stg256f(p.out_aux_f32 + o.pix * NF + c0, f);
            stg256f(p.out_aux_f32 + o.pix * NF + c0 + 8, f + 8);
        }
    }
}

// The same epilogue with ALL global operands of the tile (y and the residual: 4 steps x 2 x 32 bytes per thread)
// requested up front -- before the statistics finalize and before the accumulator is ready.  One step of look-ahead
// (sean_epilogue) keeps ~16 KB of loads in flight per SM; at ~1.5 us of loaded-memory latency that is ~6 B/clk/SM, and
// the epilogue, not the MMA, bounded the CTA-pair kernel (in-kernel accounting: 17.7 k cycles per 256-pixel tile
// against 13.9 k of MMA).  No gb_s / fp32 residual stream here (the pair kernel's K-DYN extension replaces gb_s).
template <int N_TILE, int NB>
struct SeanTileOps {
    static constexpr int CH = (N_TILE / 2 + 31) / 32;
    static constexpr int STEPS = NB * CH;
    uint4 y[STEPS][2], r[STEPS][2];
    size_t pix[NB];
    bool valid[NB];
};

template <int N_TILE, int NB>
__device__ __forceinline__ void sean_prefetch_tile(const ConvK& p, SeanTileOps<N_TILE, NB>& o, int img, int q0, int w0, int m,
                                                   int half) {
    constexpr int NF = N_TILE / 2;
    constexpr int CH = SeanTileOps<N_TILE, NB>::CH;
    const uint4 z4 = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int blk = 0; blk < NB; blk++) {
        const int q = q0 + blk * 128 + m;
        const int h = (int)__umulhi((unsigned)q, p.wp_magic);
        const int wl = q - h * p.Wp;
        const int w = w0 + wl;
        o.valid[blk] = (img < p.B) && (h < p.H) && (wl < p.Wt) && (w < p.W);
        o.pix[blk] = ((size_t)img * p.H + h) * p.W + w;
#pragma unroll
        for (int ci = 0; ci < CH; ci++) {
            const int it = blk * CH + ci, c0 = half * 16 + ci * 32;
            o.y[it][0] = o.y[it][1] = o.r[it][0] = o.r[it][1] = z4;
            if (o.valid[blk] && !DBG(p, 1)) {
                ldg256(p.y + o.pix[blk] * NF + c0, o.y[it][0], o.y[it][1]);
                if (p.resid) ldg256(p.resid + o.pix[blk] * NF + c0, o.r[it][0], o.r[it][1]);
            }
        }
    }
}

template <int N_TILE, int NB>
__device__ __forceinline__ void sean_epilogue_tile(const ConvK& p, const SeanTileOps<N_TILE, NB>& o, uint32_t t_acc,
                                                   const float* bias_t, const float* norm_s, int half) {
    constexpr int NF = N_TILE / 2;
    constexpr int CH = SeanTileOps<N_TILE, NB>::CH;
    constexpr int STEPS = NB * CH;
#pragma unroll
    for (int it = 0; it < STEPS; it++) {
        const int blk = it / CH, c0 = half * 16 + (it % CH) * 32;
        uint32_t vg[16], vb[16];
        tmem_ld16(t_acc + blk * N_TILE + c0, vg);
        tmem_ld16(t_acc + blk * N_TILE + NF + c0, vb);
        tmem_ld_wait();
        if (!o.valid[blk]) continue;
        float yv[16], gs[16], f[16];
        unpack8(o.y[it][0], yv);
        unpack8(o.y[it][1], yv + 8);
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
            const float4 bg = *reinterpret_cast<const float4*>(bias_t + c0 + j4);
            const float4 bb = *reinterpret_cast<const float4*>(bias_t + NF + c0 + j4);
            gs[j4] = __uint_as_float(vg[j4]) + bg.x;
            gs[j4 + 1] = __uint_as_float(vg[j4 + 1]) + bg.y;
            gs[j4 + 2] = __uint_as_float(vg[j4 + 2]) + bg.z;
            gs[j4 + 3] = __uint_as_float(vg[j4 + 3]) + bg.w;
            f[j4] = __uint_as_float(vb[j4]) + bb.x;
            f[j4 + 1] = __uint_as_float(vb[j4 + 1]) + bb.y;
            f[j4 + 2] = __uint_as_float(vb[j4 + 2]) + bb.z;
            f[j4 + 3] = __uint_as_float(vb[j4 + 3]) + bb.w;
        }
#pragma unroll
        for (int j2 = 0; j2 < 16; j2 += 2) {
            const float4 nm = *reinterpret_cast<const float4*>(norm_s + 2 * (c0 + j2));     // (mean, scale) x 2
            f[j2] = fmaf((yv[j2] - nm.x) * nm.y, 1.f + gs[j2], f[j2]);
            f[j2 + 1] = fmaf((yv[j2 + 1] - nm.z) * nm.w, 1.f + gs[j2 + 1], f[j2 + 1]);
        }
        if (p.inner_relu) {
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = fmaxf(f[j], 0.f);
        }
        if (DBG(p, 2)) continue;
        if (p.gamma_out) store16(p.gamma_out + o.pix[blk] * NF + c0, gs);
        if (p.resid) {
            float rr[16];
            unpack8(o.r[it][0], rr);
            unpack8(o.r[it][1], rr + 8);
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] += rr[j];
        }
        if (p.act == DASR_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = fmaxf(f[j], 0.f);
        } else if (p.act == DASR_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = f[j] > 0.f ? f[j] : 0.2f * f[j];
        }
        store16(p.out + o.pix[blk] * NF + c0, f);
    }
}

// STATS epilogue of one tile for one thread: store y = acc + bias (bf16) and accumulate, per column, the sum and the
// sum of squares of the STORED values over this warp's 32 rows of BOTH M blocks in registers; one butterfly reduction
// per 16-column chunk and tile (v1 reduced every M block separately: twice the shuffles on the critical epilogue).
// (EW = epilogue warps per TMEM lane quadrant: warp `half` of a quadrant takes the 16-column chunks half, half + EW, ...)
// (UNROLL: the pair kernel unrolls the chunk loop, so that the butterflies of one chunk overlap the TMEM load, the store
// and the accumulation of the next -- each is a latency chain of its own)
template <int N_TILE, int NB, bool PREC, int EW = 2, bool UNROLL = false>
__device__ __forceinline__ void stats_epilogue(const ConvK& p, uint32_t t_acc, const float* bias_t, float* red_s, int img,
                                               int q0, int w0, int nt, int m, int half, int ew, int lane) {
    if (half * 16 >= N_TILE) return;
    bool valid[NB];
    size_t pix[NB];
#pragma unroll
    for (int blk = 0; blk < NB; blk++) {
        const int q = q0 + blk * 128 + m;
        const int h = (int)__umulhi((unsigned)q, p.wp_magic);
        const int wl = q - h * p.Wp;
        const int w = w0 + wl;
        valid[blk] = (h < p.H) && (wl < p.Wt) && (w < p.W);
        pix[blk] = ((size_t)img * p.H + h) * p.W + w;
    }
    constexpr int NCHUNK = (N_TILE + 16 * EW - 1) / (16 * EW);
#pragma unroll(UNROLL ? NCHUNK : 1)
    for (int ck = 0; ck < NCHUNK; ck++) {
        const int c0 = half * 16 + ck * 16 * EW;
        if (c0 >= N_TILE) break;
        float s1[16], s2[16];
#pragma unroll
        for (int j = 0; j < 16; j++) s1[j] = s2[j] = 0.f;
#pragma unroll
        for (int blk = 0; blk < NB; blk++) {
            uint32_t v[16];
            tmem_ld16_acc<PREC, NB * N_TILE>(t_acc + blk * N_TILE + c0, v);
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = __uint_as_float(v[j]) + bias_t[c0 + j];
            // statistics of the values as stored (bf16-rounded), so IN(y) is self-consistent
            const uint4 o0 = pack8(f), o1 = pack8(f + 8);
            if (valid[blk]) {
                if (PREC) {      // the planes hold the fp32 value: statistics of f itself
                    store16p<true>(p.out + pix[blk] * p.Cout + nt * N_TILE + c0, p.ps_out, p.npl, f);
                } else {
                if (!DBG(p, 2)) stg256(p.out + pix[blk] * p.Cout + nt * N_TILE + c0, o0, o1);
                unpack8(o0, f);
                unpack8(o1, f + 8);
                }
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    s1[j] += f[j];
                    s2[j] = fmaf(f[j], f[j], s2[j]);
                }
            }
        }
        if (!DBG(p, 4)) {
        warp_colsum16(s1, lane);
        warp_colsum16(s2, lane);
        }
        if (!(lane & 1)) {
            const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            // per-tile partials of this warp's row quadrant: (ew, column) has exactly one owner lane -> deterministic
            float* rp2 = red_s + (ew * N_TILE + c0 + col) * 2;
            rp2[0] = s1[0];
            rp2[1] = s2[0];
        }
    }
}

// EPI (the epilogue variant) is a template parameter: with a run-time switch the epilogue warps executed ~1500
// instructions per 128x16 accumulator piece (ncu source view) and were ISSUE-bound, which also starved the MMA
// issuer that shares a scheduler with them.
// GEN (SEAN epilogue only): 16 warps -- a fourth warpgroup computes the A operand (actv) of every tile straight into
// the 128-byte-swizzled shared-memory stages the MMA reads (no actv tensor in HBM, no actv launch); registers are
// redistributed between the warpgroups with setmaxnreg (roles 56, epilogue 168, generators 120 per thread).
// LEAN (SEAN epilogue without the generator): the kernel is compiled for 128 registers per thread (launch bound of
// 512 threads, launched with 384) and setmaxnreg moves them to where they are needed (roles 40, epilogue 168), so
// that 16 K registers of the SM stay free: the actv kernel of the NEXT SEAN instance (side stream, 256 threads x 64
// registers) can then be resident next to this kernel instead of waiting for it (Engine._ActvPrefetch).
#ifndef DASR_LEAN_ROLE_REGS
#define DASR_LEAN_ROLE_REGS 40
#define DASR_LEAN_EPI_REGS 168
#endif
template <int SWZ, int N_TILE, int NB, int EPI, bool GEN, bool PREC>
__global__ void __launch_bounds__(GEN || EPI == DASR_EPI_SEAN ? 512 : kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2, const ConvK p) {
    constexpr int KC = SWZ / 2;          // channels per K chunk (one swizzle span per pixel row)
    constexpr int KSTEPS = SWZ / 32;     // UMMA K = 16 bf16 = 32 bytes
    constexpr int ACC_COLS = NB * N_TILE;
    // (PREC: one tile in flight, its two accumulator sets -- main / low-order terms -- take the place of the double buffer)
    constexpr int NACC = PREC ? 1 : ((2 * ACC_COLS <= 512) ? 2 : 1);
    static_assert(!PREC || 2 * ACC_COLS <= 512, "PREC needs two accumulator sets in TMEM");
    constexpr int TMEM_COLS = 512;

    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
    __shared__ uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];
    __shared__ uint64_t acc_full[2], acc_empty[2];
    __shared__ uint64_t a2_full, a2_empty;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float norm_s[512];         // SEAN: (mean, scale) of the image; STATS: scratch of the fused finalize
    __shared__ __align__(16) float bias_s[kMaxBias];
    __shared__ uint32_t tap_lo_s[81];     // descriptor-low-word offset of tap (t,u): ((t*Wp + u) * SWZ) >> 4

    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* a_smem = smem;
    uint8_t* b_smem = smem + (size_t)p.SA * p.a_stage_bytes;
    uint8_t* a2_smem = smem + p.a2_off;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    PROF_KERNEL_BEGIN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.SA; i++) {
            mbar_init(&a_full[i], GEN ? 64 : 1);       // GEN: the 64 generator threads of a channel chunk arrive
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < p.SB; i++) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 8);
        }
        mbar_init(&a2_full, 1);
        mbar_init(&a2_empty, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        if (p.dyn) {
            tma_prefetch_desc(&mapA2);
            tma_prefetch_desc(&mapB2);
        }
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(&tmem_base_s);
    // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlaps the tail of the
    // previous kernel in the stream; nothing below may run before that kernel has completed and flushed.  Our own
    // dependents may be scheduled as soon as every CTA of this grid got here (they block at the same point).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int i = threadIdx.x; i < p.n_bias; i += (GEN ? 512 : kThreads)) bias_s[i] = __ldg(p.bias + i);
    if (threadIdx.x < p.taps) {
        const int t = threadIdx.x / p.kw, u = threadIdx.x - t * p.kw;
        tap_lo_s[threadIdx.x] = (uint32_t)((t * p.Wp + u) * SWZ) >> 4;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int tiles_per_img = p.n_strips * p.tiles_per_strip * p.ntn;
    const int nch_t = PREC ? p.nch * p.n_terms : p.nch;      // K chunks x plane cross terms

    // GEN: the register file is redistributed between the warpgroups at the top of each warpgroup-level branch
    // (setmaxnreg applies to the code it dominates): roles 56, epilogue 168, generators 120 registers per thread
    if (warp < 4) {
    if (GEN) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
    else if (EPI == DASR_EPI_SEAN) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(DASR_LEAN_ROLE_REGS));
    if (warp == 0) {
        // ===================================================== A producer: one halo box per K chunk
        // (elect.sync under a warp-uniform branch: ptxas then issues TMA/MMA straight from uniform registers;
        //  a `lane == 0` test costs a 72-cycle waterfall loop per tcgen05.mma -- tools/rate_probe.cu)
        if (elect_one()) {
            uint32_t a_it = 0, t_it = 0;
            PROF_DECL;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int img = tile / tiles_per_img;
                int r = tile - img * tiles_per_img;
                const int strip = r / (p.tiles_per_strip * p.ntn);
                r -= strip * (p.tiles_per_strip * p.ntn);
                const int tps = r / p.ntn;
                const int q0 = tps * (NB * 128);
                const int r0 = q0 / p.Wp;
                const int w0 = strip * p.Wt;
                for (int cc = 0; cc < (GEN ? 0 : nch_t); cc++) {
                    const int term = PREC ? cc / p.nch : 0;
                    const int c = cc - term * p.nch;
                    const int sa = a_it % p.SA;
                    const uint32_t ph = (a_it / p.SA) & 1;
                    PROF_LAP(1);
                    mbar_wait(&a_empty[sa], ph ^ 1);
                    PROF_LAP(0);
                    mbar_expect_tx(&a_full[sa], p.a_tx_bytes);
                    tma_load_4d(a_smem + (size_t)sa * p.a_stage_bytes, &mapA, &a_full[sa], c * KC,
                                w0 - p.pad_w, r0 - p.pad_h, img + (PREC ? (int)p.ta[term] * p.B : 0));
                    a_it++;
                }
                if (p.dyn) {       // the mask patch of this tile (single buffer, released by the MMA issuer)
                    mbar_wait(&a2_empty, (t_it & 1) ^ 1);
                    mbar_expect_tx(&a2_full, p.a2_tx_bytes);
                    tma_load_4d(a2_smem, &mapA2, &a2_full, 0, w0 - p.pad_w, r0 - p.pad_h, img);
                }
                t_it++;
            }
            PROF_FLUSH(6, 1);
        }
    } else if (warp == 3) {
        // ===================================================== B producer: weight tiles (tap, chunk)
        if (elect_one()) {
            uint32_t b_it = 0;
            bool first = true;
            PROF_DECL;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nt = tile % p.ntn;
                const int wrow0 = p.w_img_rows ? (tile / tiles_per_img) * p.w_img_rows : 0;
                if (p.b_resident && !first) continue;
                for (int cc = 0; cc < nch_t; cc++) {
                    const int term = PREC ? cc / p.nch : 0;
                    const int c = cc - term * p.nch;
                    const int wrow = wrow0 + nt * N_TILE + (PREC ? (int)p.tb[term] * p.w_plane_rows : 0);
                    for (int tap = 0; tap < p.taps; tap++) {
                        int sb;
                        if (p.b_resident) {
                            sb = c * p.taps + tap;
                        } else {
                            sb = b_it % p.SB;
                            const uint32_t ph = (b_it / p.SB) & 1;
                            PROF_LAP(1);
                            mbar_wait(&b_empty[sb], ph ^ 1);
                            PROF_LAP(0);
                            b_it++;
                        }
                        mbar_expect_tx(&b_full[sb], p.b_tx_bytes);
                        tma_load_2d(b_smem + (size_t)sb * p.b_stage_bytes, &mapB, &b_full[sb],
                                    tap * p.Cin + c * KC, wrow);
                    }
                }
                if (p.dyn) {       // this image's dynamic filters, one [Cout x 16] tile per tap through the same ring
                    const int img = tile / tiles_per_img;
                    for (int pl = 0; pl < (PREC ? p.npl : 1); pl++)     // the mask image is exact: one plane
                    for (int tap = 0; tap < p.taps; tap++) {
                        const int sb = b_it % p.SB;
                        const uint32_t ph = (b_it / p.SB) & 1;
                        mbar_wait(&b_empty[sb], ph ^ 1);
                        b_it++;
                        mbar_expect_tx(&b_full[sb], p.b2_tx_bytes);
                        tma_load_2d(b_smem + (size_t)sb * p.b_stage_bytes, &mapB2, &b_full[sb], tap * 16,
                                    img * N_TILE + pl * p.dynw_plane_rows);
                    }
                }
                first = false;
            }
            PROF_FLUSH(7, 1);
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (single elected thread)
        if (elect_one()) {
            // Descriptor low words are additive in the smem address: lo(addr + x) = lo(addr) + (x >> 4).  All
            // per-tap / per-block / per-k offsets are precomputed (tap_lo_s) or immediates, so the loop body is a
            // handful of 32-bit adds per tcgen05.mma (the tensor pipe, not this thread, must be the limiter).
            const uint32_t idesc = make_idesc_bf16(128, N_TILE);
            const uint64_t desc0 = make_smem_desc<SWZ>(0, 0);
            const uint32_t desc_hi = (uint32_t)(desc0 >> 32);
            const uint32_t lo_flags = (uint32_t)desc0 & ~0x3FFFu;
            const uint32_t a_lo0 = lo_flags | ((smem_u32(a_smem) & 0x3FFFFu) >> 4);
            const uint32_t b_lo0 = lo_flags | ((smem_u32(b_smem) & 0x3FFFFu) >> 4);
            const uint32_t a_stage_lo = p.a_stage_bytes >> 4, b_stage_lo = p.b_stage_bytes >> 4;
            constexpr uint32_t BLK_LO = (128u * SWZ) >> 4;
            uint32_t a_it = 0, b_it = 0, acc_it = 0, t_it = 0;
            const uint32_t desc_hi32 = (uint32_t)(make_smem_desc<32>(0, 0) >> 32);
            const uint32_t a2_lo0 = lo_flags | ((smem_u32(a2_smem) & 0x3FFFFu) >> 4);
            bool first = true;
            PROF_DECL;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int r = tile % tiles_per_img;
                r %= (p.tiles_per_strip * p.ntn);
                const int tps = r / p.ntn;
                const int q0 = tps * (NB * 128);
                const int soff = q0 - (q0 / p.Wp) * p.Wp;  // first output row inside the smem patch
                const int buf = acc_it % NACC;
                const uint32_t aph = (acc_it / NACC) & 1;
                PROF_LAP(3);
                mbar_wait(&acc_empty[buf], aph ^ 1);
                PROF_LAP(0);
                tc_fence_after();
                const uint32_t d_tmem0 = tmem_base + buf * ACC_COLS;
                for (int c = 0; c < nch_t; c++) {
                    // PREC: chunks of the lower-order plane terms (c >= nch) accumulate into the second set
                    const bool low = PREC && c >= p.nch;
                    const uint32_t d_tmem = d_tmem0 + (low ? ACC_COLS : 0);
                    const int c_rel = low ? c - p.nch : c;
                    const int sa = a_it % p.SA;
                    PROF_LAP(3);
                    mbar_wait(&a_full[sa], (a_it / p.SA) & 1);
                    PROF_LAP(1);
                    tc_fence_after();
                    const uint32_t a_lo_tile = a_lo0 + sa * a_stage_lo + (uint32_t)soff * (SWZ >> 4);
                    if (p.b_resident) {
                        if (first) {
                            for (int tap = 0; tap < p.taps; tap++) mbar_wait(&b_full[c * p.taps + tap], 0);
                            tc_fence_after();
                        }
                        uint32_t b_lo = b_lo0 + (uint32_t)(c * p.taps) * b_stage_lo;
#pragma unroll 1
                        for (int tap = 0; tap < p.taps; tap++, b_lo += b_stage_lo) {
                            const uint32_t a_lo = a_lo_tile + tap_lo_s[tap];
#pragma unroll
                            for (int blk = 0; blk < NB; blk++) {
#pragma unroll
                                for (int k = 0; k < KSTEPS; k++)
                                    umma_bf16_lohi(d_tmem + blk * N_TILE, a_lo + blk * BLK_LO + k * 2, b_lo + k * 2,
                                                   desc_hi, idesc, (c_rel | tap | k) != 0);
                            }
                        }
                    } else {
#pragma unroll 1
                        for (int tap = 0; tap < p.taps; tap++) {
                            const int sb = b_it % p.SB;
                            PROF_LAP(3);
                            mbar_wait(&b_full[sb], (b_it / p.SB) & 1);
                            PROF_LAP(2);
                            tc_fence_after();
                            const uint32_t a_lo = a_lo_tile + tap_lo_s[tap];
                            const uint32_t b_lo = b_lo0 + sb * b_stage_lo;
#pragma unroll
                            for (int blk = 0; blk < NB; blk++) {
#pragma unroll
                                for (int k = 0; k < KSTEPS; k++)
                                    umma_bf16_lohi(d_tmem + blk * N_TILE, a_lo + blk * BLK_LO + k * 2, b_lo + k * 2,
                                                   desc_hi, idesc, (c_rel | tap | k) != 0);
                            }
                            umma_commit(&b_empty[sb]);
                            b_it++;
                        }
                    }
                    umma_commit(&a_empty[sa]);
                    a_it++;
                }
                if (p.dyn) {
                    // K extension: gb_s = sum_{tap,k} mask[p+tap][k] * T[img][k][tap][:]  (K = 16 per tap, 32-byte rows)
                    PROF_LAP(3);
                    mbar_wait(&a2_full, t_it & 1);
                    PROF_LAP(1);
                    tc_fence_after();
                    const uint32_t a2_lo_tile = a2_lo0 + (uint32_t)soff * (32u >> 4);
#pragma unroll 1
                    for (int ti = 0; ti < (PREC ? p.taps * p.npl : p.taps); ti++) {
                        const int tap = PREC ? ti % p.taps : ti;
                        const uint32_t d_tmem = d_tmem0 + ((PREC && ti >= p.taps) ? ACC_COLS : 0);   // lower filter planes
                        const int sb = b_it % p.SB;
                        PROF_LAP(3);
                        mbar_wait(&b_full[sb], (b_it / p.SB) & 1);
                        PROF_LAP(2);
                        tc_fence_after();
                        const uint32_t a_lo = a2_lo_tile + (tap_lo_s[tap] * 32u) / SWZ;     // row offset at 32 B per row
                        const uint32_t b_lo = b_lo0 + sb * b_stage_lo;
#pragma unroll
                        for (int blk = 0; blk < NB; blk++)
                            umma_bf16_lohi(d_tmem + blk * N_TILE, a_lo + blk * ((128u * 32u) >> 4), b_lo, desc_hi32, idesc, 1);
                        umma_commit(&b_empty[sb]);
                        b_it++;
                    }
                    umma_commit(&a2_empty);
                }
                t_it++;
                umma_commit(&acc_full[buf]);
                acc_it++;
                first = false;
            }
            PROF_LAP(3);
            PROF_FLUSH(0, 4);
        }
    }
    } else if (GEN && warp >= 12) {
        // ===================================================== A generator (warps 12..15): actv tiles in shared memory
        asm volatile("setmaxnreg.dec.sync.aligned.u32 120;\n");
        // thread = (channel chunk c = stage, 16-byte piece g of the 128-byte pixel row, one of 8 row runs)
        const int gt = threadIdx.x - 384;
        const int c = gt >> 6;
        const int g = gt & 7;
        const int run = (gt & 63) >> 3;
        float2 wr[9][4], br[4];
        {
            const int ch0 = c * 64 + g * 8;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                br[j] = make_float2(__ldg(p.gen_b + ch0 + 2 * j), __ldg(p.gen_b + ch0 + 2 * j + 1));
#pragma unroll
                for (int t = 0; t < 9; t++)
                    wr[t][j] = make_float2(__ldg(p.gen_w + (ch0 + 2 * j) * 9 + t), __ldg(p.gen_w + (ch0 + 2 * j + 1) * 9 + t));
            }
        }
        float* dsm = reinterpret_cast<float*>(smem + p.gen_off);      // (RB + 2) x (Wp + 2) depth halo, zero padded
        const int DW = p.Wp + 2;
        const int nrows = p.RB * p.Wp;
        const int per_run = (nrows + 7) >> 3;
        uint8_t* stage = a_smem + (size_t)c * p.a_stage_bytes;
        uint32_t t_it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, t_it++) {
            const int img = tile / tiles_per_img;
            int r = tile - img * tiles_per_img;
            const int strip = r / (p.tiles_per_strip * p.ntn);
            r -= strip * (p.tiles_per_strip * p.ntn);
            const int tps = r / p.ntn;
            const int q0 = tps * (NB * 128);
            const int y0 = q0 / p.Wp - p.pad_h;           // image row / column of patch position (0, 0)
            const int x0 = strip * p.Wt - p.pad_w;
            asm volatile("bar.sync 2, 128;\n" ::: "memory");          // the previous tile's readers are done
            const float* dp = p.gen_depth + (size_t)img * p.H * p.W;
            for (int i = gt; i < (p.RB + 2) * DW; i += 128) {
                const int rr = i / DW, cc = i - rr * DW;
                const int hh = y0 - 1 + rr, ww = x0 - 1 + cc;
                dsm[i] = (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) ? __ldg(dp + (size_t)hh * p.W + ww) : 0.f;
            }
            asm volatile("bar.sync 2, 128;\n" ::: "memory");
            mbar_wait(&a_empty[c], (t_it & 1) ^ 1);
            const int pr0 = run * per_run, pr1 = min(nrows, pr0 + per_run);
            // (two pixels per iteration -- 8 FMA chains -- measured slower: 155 vs 138 us; the generator is bound by the
            // issue slots it shares with the epilogue warps, not by latency)
            int pi = pr0 / p.Wp, pj = pr0 - pi * p.Wp;
            for (int pr = pr0; pr < pr1; pr++) {
                const int hh = y0 + pi, ww = x0 + pj;
                uint4 o = make_uint4(0, 0, 0, 0);
                if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) {         // outside the image actv is ZERO (padding)
                    const float* d0 = dsm + pi * DW + pj;               // window origin = (hh - 1, ww - 1)
                    float2 acc[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[j] = br[j];
#pragma unroll
                    for (int t = 0; t < 3; t++)
#pragma unroll
                        for (int u = 0; u < 3; u++) {
                            const float d = d0[t * DW + u];
                            const float2 d2 = make_float2(d, d);
#pragma unroll
                            for (int j = 0; j < 4; j++) acc[j] = __ffma2_rn(d2, wr[t * 3 + u][j], acc[j]);
                        }
                    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                    for (int j = 0; j < 4; j++) oh[j] = __floats2bfloat162_rn(fmaxf(acc[j].x, 0.f), fmaxf(acc[j].y, 0.f));
                }
                // 128-byte swizzle of the K-major operand: 16-byte piece index XOR (row & 7)
                *reinterpret_cast<uint4*>(stage + (size_t)pr * 128 + ((g ^ (pr & 7)) << 4)) = o;
                if (++pj == p.Wp) {
                    pj = 0;
                    pi++;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> tensor-core reads
            mbar_arrive(&a_full[c]);
        }
    } else if (warp >= 4 && warp < 12) {
        if (GEN) asm volatile("setmaxnreg.inc.sync.aligned.u32 168;\n");
        else if (EPI == DASR_EPI_SEAN) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(DASR_LEAN_EPI_REGS));
        // ===================================================== epilogue warps (8: 2 per TMEM lane quadrant)
        // warp w may only read TMEM lanes 32*(w%4)..+31; the two warps of a quadrant take the even / odd
        // 16-column chunks.  Operands that live in HBM (residual, y, gb_s) are requested BEFORE the TMEM load
        // so that their latency overlaps it; bias / norm come from shared memory.
        const int ew = warp & 3;              // TMEM lane group
        const int half = (warp - 4) >> 2;     // 0: even chunks, 1: odd chunks
        const int m = ew * 32 + lane;         // accumulator row inside an M block
        const int et = threadIdx.x - 128;     // 0..255
        uint32_t acc_it = 0;
        PROF_DECL;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int img = tile / tiles_per_img;
            int r = tile - img * tiles_per_img;
            const int strip = r / (p.tiles_per_strip * p.ntn);
            r -= strip * (p.tiles_per_strip * p.ntn);
            const int tps = r / p.ntn;
            const int nt = r - tps * p.ntn;
            const int q0 = tps * (NB * 128);
            const int w0 = strip * p.Wt;
            const int buf = acc_it % NACC;
            const uint32_t aph = (acc_it / NACC) & 1;
            const float* bias_t = bias_s + nt * N_TILE;

            if (EPI == DASR_EPI_SEAN) {
                // (mean, scale) of this image for the nf = N_TILE/2 normalised channels
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
                if (p.norm) {
                    if (et < N_TILE) norm_s[et] = __ldg(p.norm + (size_t)img * N_TILE + et);
                } else if (et < N_TILE / 2) {
                    // fused dasr_instats_finalize: the per-tile partial sums of the producing convolution
                    // ([B][nslots][nf][2], 17 slots at 64x64) are reduced here in slot order -- bit-reproducible --
                    // while the MMAs of this tile are still running; no separate finalize launch
                    constexpr int NF = N_TILE / 2;
                    const float2* sp = reinterpret_cast<const float2*>(p.stats) + (size_t)img * p.nslots * NF + et;
                    float a1 = 0.f, a2 = 0.f;
                    for (int sl = 0; sl < p.nslots; sl += 8) {
                        float2 v[8];
#pragma unroll
                        for (int u = 0; u < 8; u++)
                            v[u] = (sl + u < p.nslots) ? __ldg(sp + (size_t)(sl + u) * NF) : make_float2(0.f, 0.f);
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            a1 += v[u].x;
                            a2 += v[u].y;
                        }
                    }
                    const float inv_hw = 1.f / (float)(p.H * p.W);
                    const float mean = a1 * inv_hw;
                    const float var = fmaxf(a2 * inv_hw - mean * mean, 0.f);
                    const float eps = 1e-5f;
                    // IN(IN(y)) = (y - mean) * (var+eps)^-1/2 * (var/(var+eps) + eps)^-1/2
                    const float r1 = rsqrtf(var + eps);
                    const float r2 = rsqrtf(var * r1 * r1 + eps);
                    norm_s[2 * et] = mean;
                    norm_s[2 * et + 1] = r1 * r2;
                    if (p.norm_out && tps == 0 && strip == 0) {      // saved for the backward pass
                        p.norm_out[((size_t)img * NF + et) * 2] = mean;
                        p.norm_out[((size_t)img * NF + et) * 2 + 1] = r1 * r2;
                        if (p.normk_out) {
                            const float a = var + eps, rr = var / a + eps;
                            p.normk_out[(size_t)img * NF + et] = 1.f / a + eps / (a * a * rr);
                        }
                    }
                }
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
            }

            // STORE epilogue with few pieces per thread (N tile <= 64): the residual and the activation mask of the WHOLE
            // tile (<= 4 pieces of 32 bytes each) are requested BEFORE the accumulator is awaited -- with the loads issued
            // piece by piece after the wait their latency was exposed once per piece (the residual convolution of a
            // classic block took 261 us against 159 us without residual for 1.5x the bytes).
            constexpr int PRE_CH = (N_TILE + 31) / 32;
            constexpr bool PRE = (EPI == DASR_EPI_STORE) && !PREC && (NB * PRE_CH <= 4);
            uint4 pre_r[PRE ? NB * PRE_CH : 1][2], pre_m[PRE ? NB * PRE_CH : 1][2];
            bool pre_valid[PRE ? NB : 1];
            __nv_bfloat16* pre_op[PRE ? NB : 1];
            if (PRE && half * 16 < N_TILE) {
#pragma unroll
                for (int blk = 0; blk < NB; blk++) {
                    const int q = q0 + blk * 128 + m;
                    const int h = (int)__umulhi((unsigned)q, p.wp_magic);
                    const int wl = q - h * p.Wp;
                    const int w = w0 + wl;
                    bool valid = (h < p.H) && (wl < p.Wt) && (w < p.W);
                    int ho = h, wo = w;
                    if (p.subsample == 2) {
                        valid = valid && !(h & 1) && !(w & 1);
                        ho = h >> 1;
                        wo = w >> 1;
                    }
                    const size_t pix = ((size_t)img * p.Ho + ho) * p.Wo + wo;
                    __nv_bfloat16* op = p.out + pix * p.Cout + nt * N_TILE;
                    if (p.unshuffle)
                        op = p.out + ((((size_t)img * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1)) * 4 +
                                      ((ho & 1) * 2 + (wo & 1))) * p.Cout + nt * N_TILE;
                    pre_valid[blk] = valid;
                    pre_op[blk] = op;
#pragma unroll
                    for (int ci = 0; ci < PRE_CH; ci++) {
                        const int c0 = half * 16 + ci * 32;
                        const uint4 z4 = make_uint4(0, 0, 0, 0);
                        pre_r[blk * PRE_CH + ci][0] = pre_r[blk * PRE_CH + ci][1] = z4;
                        pre_m[blk * PRE_CH + ci][0] = pre_m[blk * PRE_CH + ci][1] = z4;
                        if (c0 < N_TILE && valid && !DBG(p, 1)) {
                            if (p.resid) ldg256(p.resid + pix * p.Cout + nt * N_TILE + c0, pre_r[blk * PRE_CH + ci][0], pre_r[blk * PRE_CH + ci][1]);
                            if (p.actmask) ldg256(p.actmask + pix * p.Cout + nt * N_TILE + c0, pre_m[blk * PRE_CH + ci][0], pre_m[blk * PRE_CH + ci][1]);
                        }
                    }
                }
            }

            PROF_LAP(1);
            mbar_wait(&acc_full[buf], aph);
            PROF_LAP(0);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + buf * ACC_COLS + (uint32_t(ew * 32) << 16);

            if (PRE) {
                if (half * 16 < N_TILE) {
#pragma unroll
                    for (int blk = 0; blk < NB; blk++) {
#pragma unroll
                        for (int ci = 0; ci < PRE_CH; ci++) {
                            const int c0 = half * 16 + ci * 32;
                            if (c0 >= N_TILE) continue;
                            uint32_t v[16];
                            tmem_ld16(t_acc + blk * N_TILE + c0, v);
                            tmem_ld_wait();
                            float f[16];
#pragma unroll
                            for (int j4 = 0; j4 < 16; j4 += 4) {
                                const float4 bb = *reinterpret_cast<const float4*>(bias_t + c0 + j4);
                                f[j4] = __uint_as_float(v[j4]) + bb.x;
                                f[j4 + 1] = __uint_as_float(v[j4 + 1]) + bb.y;
                                f[j4 + 2] = __uint_as_float(v[j4 + 2]) + bb.z;
                                f[j4 + 3] = __uint_as_float(v[j4 + 3]) + bb.w;
                            }
                            if (p.resid) {
                                float rr[16];
                                unpack8(pre_r[blk * PRE_CH + ci][0], rr);
                                unpack8(pre_r[blk * PRE_CH + ci][1], rr + 8);
#pragma unroll
                                for (int j = 0; j < 16; j++) f[j] += rr[j];
                            }
#pragma unroll
                            for (int j = 0; j < 16; j++) f[j] = apply_act(f[j], p.act);
                            if (p.actmask) {
                                float mm[16];
                                unpack8(pre_m[blk * PRE_CH + ci][0], mm);
                                unpack8(pre_m[blk * PRE_CH + ci][1], mm + 8);
#pragma unroll
                                for (int j = 0; j < 16; j++) f[j] *= (mm[j] > 0.f ? 1.f : p.mask_slope);
                            }
                            if (pre_valid[blk] && !DBG(p, 2)) store16(pre_op[blk] + c0, f);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
                acc_it++;
                continue;
            }

            if (EPI == DASR_EPI_SEAN) {
                sean_epilogue<N_TILE, NB, PREC>(p, t_acc, bias_t, norm_s, img, q0, w0, m, half);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
                acc_it++;
                continue;
            }
            if (EPI == DASR_EPI_STATS)
                stats_epilogue<N_TILE, NB, PREC>(p, t_acc, bias_t, norm_s, img, q0, w0, nt, m, half, ew, lane);
#pragma unroll 1
            for (int blk = 0; blk < (EPI == DASR_EPI_STATS ? 0 : NB); blk++) {
                const int q = q0 + blk * 128 + m;
                const int h = (int)__umulhi((unsigned)q, p.wp_magic);     // q / Wp (exact for q < 2^32 / Wp)
                const int wl = q - h * p.Wp;
                const int w = w0 + wl;
                bool valid = (h < p.H) && (wl < p.Wt) && (w < p.W);
                const uint32_t t_blk = t_acc + blk * N_TILE;

                if (EPI == DASR_EPI_STORE) {      // (EPI_STATS has its own function, stats_epilogue)
                    int ho = h, wo = w;
                    if (p.subsample == 2) {
                        valid = valid && !(h & 1) && !(w & 1);
                        ho = h >> 1;
                        wo = w >> 1;
                    }
                    const size_t pix = ((size_t)img * p.Ho + ho) * p.Wo + wo;
                    __nv_bfloat16* op = p.out + pix * p.Cout + nt * N_TILE;
                    if (p.unshuffle)        // pixel (ho, wo) -> channel block 2*(ho%2) + wo%2 of pixel (ho/2, wo/2)
                        op = p.out + ((((size_t)img * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1)) * 4 +
                                      ((ho & 1) * 2 + (wo & 1))) * p.Cout + nt * N_TILE;
                    const __nv_bfloat16* rp = p.resid ? p.resid + pix * p.Cout + nt * N_TILE : nullptr;
                    const __nv_bfloat16* mp = p.actmask ? p.actmask + pix * p.Cout + nt * N_TILE : nullptr;
#pragma unroll 1
                    for (int c0 = half * 16; c0 < N_TILE; c0 += 32) {
                        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, m0 = r0, m1 = r0;
                        if (rp && valid && !DBG(p, 1)) ldg256(rp + c0, r0, r1);
                        if (mp && valid && !DBG(p, 1)) ldg256(mp + c0, m0, m1);
                        uint32_t v[16];
                        tmem_ld16_acc<PREC, ACC_COLS>(t_blk + c0, v);
                        float f[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) f[j] = __uint_as_float(v[j]) + bias_t[c0 + j];
                        {
                            if (rp) {
                                float rr[16];
                                unpack8(r0, rr);
                                unpack8(r1, rr + 8);
                                if (valid) add_planes16<PREC>(rp + c0, p.ps_out, p.npl, rr);
#pragma unroll
                                for (int j = 0; j < 16; j++) f[j] += rr[j];
                            }
#pragma unroll
                            for (int j = 0; j < 16; j++) f[j] = apply_act(f[j], p.act);
                            if (mp) {
                                float mm[16];
                                unpack8(m0, mm);
                                unpack8(m1, mm + 8);
#pragma unroll
                                for (int j = 0; j < 16; j++) f[j] *= (mm[j] > 0.f ? 1.f : p.mask_slope);
                            }
                            if (valid && !DBG(p, 2)) store16p<PREC>(op + c0, p.ps_out, p.npl, f);
                        }
                    }
                } else if (EPI == DASR_EPI_SHUFFLE2) {
                    const int Cq = p.Cout >> 2;  // channels after the shuffle
#pragma unroll 1
                    for (int c0 = half * 16; c0 < N_TILE; c0 += 32) {
                        uint32_t v[16];
                        tmem_ld16_acc<PREC, ACC_COLS>(t_blk + c0, v);
                        if (valid) {
                            const int n0 = nt * N_TILE + c0;  // permuted row: n0 = s*Cq + c
                            const int s = n0 / Cq, c = n0 - s * Cq;
                            const size_t pix = ((size_t)img * p.Ho + (2 * h + (s >> 1))) * p.Wo + (2 * w + (s & 1));
                            float f[16];
#pragma unroll
                            for (int j = 0; j < 16; j++)
                                f[j] = apply_act(__uint_as_float(v[j]) + bias_t[c0 + j], p.act);
                            if (!DBG(p, 2)) store16p<PREC>(p.out + pix * Cq + c, p.ps_out, p.npl, f);
                        }
                    }
                } else {  // DASR_EPI_NCHW_F32
#pragma unroll 1
                    for (int c0 = half * 16; c0 < N_TILE; c0 += 32) {
                        uint32_t v[16];
                        tmem_ld16(t_blk + c0, v);
                        tmem_ld_wait();
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                const int co = nt * N_TILE + c0 + j;
                                if (co < p.Cout) {
                                    float f = apply_act(__uint_as_float(v[j]) + bias_t[c0 + j], p.act);
                                    if (p.clamp01) f = fminf(fmaxf(f, 0.f), 1.f);
                                    p.out_f32[(((size_t)img * p.Cout + co) * p.H + h) * p.W + w] = f;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            acc_it++;
            if (EPI == DASR_EPI_STATS) {
                // one statistics slot per (image, tile): the four row quadrants are added in a fixed order
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
                if (et < N_TILE) {
                    float a1 = 0.f, a2 = 0.f;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        a1 += norm_s[(q * N_TILE + et) * 2];
                        a2 += norm_s[(q * N_TILE + et) * 2 + 1];
                    }
                    const int slot = strip * p.tiles_per_strip + tps;
                    float2* sp = reinterpret_cast<float2*>(p.stats) + ((size_t)img * p.nslots + slot) * p.Cout + nt * N_TILE + et;
                    *sp = make_float2(a1, a2);
                }
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
            }
        }
        PROF_LAP(1);
#ifdef DASR_PROFILE
        if (threadIdx.x == 128) PROF_FLUSH(4, 2);
#endif
    }

    tc_fence_before();
    __syncthreads();
    PROF_KERNEL_END;
    if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

#ifndef DASR_CONV_PRECISE_TU
// ------------------------------------------------------------------------------------------------ SEAN conv, CTA pairs
// The [gamma_o; beta_o] convolution of SEAN (128 -> 128, 3x3, K-DYN extension, SEAN epilogue) on CTA PAIRS:
// tcgen05.mma.cta_group::2, M = 256.  With one CTA per tile an M128 x N128 x K16 SS-MMA reads 4 KB of A and 4 KB of B
// from shared memory per 64 tensor cycles -- exactly the 128 B/clk a B200 SM can deliver -- so the TMA fills of the
// 288 KB weight ring and the epilogue's own shared-memory reads push the MMA to ~110 cycles (ncu, round 1).  A pair
// shares ONE weight stream: each CTA stages only HALF of every weight tile (rows rank*64..+63), the tensor cores of
// both SMs read both halves, and each CTA multiplies them with its own 128 pixel rows: 6 KB per MMA and CTA, half the
// weight fill per CTA.
//   * Pairing: the two CTAs of a pair work on the SAME tile position of TWO CONSECUTIVE IMAGES (identical geometry,
//     one set of descriptors).  The main K loop is shared.  The K-DYN extension has per-image filters; it runs as two
//     passes, pass j with the filters of image j, and every CTA keeps TWO mask-patch buffers at the same offsets: its
//     own patch in buffer `rank`, zeros in the other -- so the foreign pass adds 0 to its accumulator.
//   * Only the leader (cluster rank 0) issues MMAs.  Both CTAs run TMA producers; their transaction bytes are counted
//     on the LEADER's full barriers: the leader's producer posts ONE arrive.expect_tx for the bytes of both CTAs, the
//     peer's producer only issues its TMA (a remote arrive per stage is a cluster-scope release: ~2000 cycles each,
//     measured -- it made the kernel 2.2x slower than the single-CTA one).  "Stage free" / "accumulator ready" come from a
//     multicast tcgen05.commit that arrives on the same barrier in both CTAs; "accumulator drained" is counted on the
//     leader's acc_empty (8 local + 8 remote epilogue warps).
//   * Epilogue = sean_epilogue (unchanged), each CTA on its own TMEM half and its own image.
constexpr int kPairSB = 8;
template <int N_TILE, int NB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1)
sean_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2, const ConvK p) {
    constexpr int SWZ = 128, KC = 64, KSTEPS = 4;
    constexpr int ACC_COLS = NB * N_TILE;
    constexpr int TMEM_COLS = 512;
    static_assert(2 * ACC_COLS <= 512, "two accumulator buffers must fit TMEM");

    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[2], a_empty[2];
    __shared__ uint64_t b_full[kPairSB], b_empty[kPairSB];
    __shared__ uint64_t acc_full[2], acc_empty[2];
    __shared__ uint64_t a2_full, a2_empty;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float norm_s[512];
    __shared__ __align__(16) float bias_s[N_TILE];
    __shared__ uint32_t tap_lo_s[9];

    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* a_smem = smem;
    uint8_t* b_smem = smem + 2 * (size_t)p.a_stage_bytes;
    uint8_t* a2_smem = smem + p.a2_off;                 // two buffers of a2_bytes: [0] image 0's patch, [1] image 1's
    const uint32_t a2_bytes = (p.a2_tx_bytes + 1023u) & ~1023u;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    PROF_KERNEL_BEGIN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&a_full[i], 1);           // leader only: ONE arrive.expect_tx for the bytes of BOTH CTAs
            mbar_init(&a_empty[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 16);       // 8 epilogue warps x 2 CTAs (leader only)
        }
        for (int i = 0; i < kPairSB; i++) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        mbar_init(&a2_full, 1);
        mbar_init(&a2_empty, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        tma_prefetch_desc(&mapA2);
        tma_prefetch_desc(&mapB2);
    }
    if (warp == 2) tmem_alloc_pair<TMEM_COLS>(&tmem_base_s);
    // the foreign mask-patch buffer stays zero for the whole kernel
    {
        uint4* z = reinterpret_cast<uint4*>(a2_smem + (size_t)(1 - rank) * a2_bytes);
        for (int i = threadIdx.x; i < (int)(a2_bytes / 16); i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int i = threadIdx.x; i < N_TILE; i += kThreads) bias_s[i] = __ldg(p.bias + i);
    if (threadIdx.x < 9) {
        const int t = threadIdx.x / 3, u = threadIdx.x - t * 3;
        tap_lo_s[threadIdx.x] = (uint32_t)((t * p.Wp + u) * SWZ) >> 4;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // barriers of both CTAs are initialised before any remote arrive / TMA
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int tiles_per_img = p.n_strips * p.tiles_per_strip;
    const int n_pairs = gridDim.x >> 1;
    const int pair_id = blockIdx.x >> 1;
    // the leader's copies of the barriers the producers of BOTH CTAs signal
    const uint32_t a2_full_l = mapa_u32(smem_u32(&a2_full), 0);

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(DASR_LEAN_ROLE_REGS));
    if (warp == 0) {
        // ===================================================== A producer (both CTAs): own image, own patches
        if (elect_one()) {
            uint32_t a_it = 0, t_it = 0;
            for (int tile = pair_id; tile < p.total_tiles; tile += n_pairs, t_it++) {
                const int ip = tile / tiles_per_img;
                int r = tile - ip * tiles_per_img;
                const int strip = r / p.tiles_per_strip;
                const int tps = r - strip * p.tiles_per_strip;
                const int img = 2 * ip + (int)rank;            // >= B for the odd image out: TMA zero-fills
                const int q0 = tps * (NB * 128);
                const int r0 = q0 / p.Wp;
                const int w0 = strip * p.Wt;
                for (int c = 0; c < 2; c++, a_it++) {
                    const int sa = a_it & 1;
                    mbar_wait(&a_empty[sa], ((a_it >> 1) & 1) ^ 1);
                    const uint32_t full_l = mapa_u32(smem_u32(&a_full[sa]), 0);
                    if (leader) mbar_expect_tx(&a_full[sa], 2 * p.a_tx_bytes);
                    tma_load_4d_pair(a_smem + (size_t)sa * p.a_stage_bytes, &mapA, full_l, c * KC, w0 - p.pad_w,
                                     r0 - p.pad_h, img);
                }
                mbar_wait(&a2_empty, (t_it & 1) ^ 1);
                if (leader) mbar_expect_tx(&a2_full, 2 * p.a2_tx_bytes);
                tma_load_4d_pair(a2_smem + (size_t)rank * a2_bytes, &mapA2, a2_full_l, 0, w0 - p.pad_w, r0 - p.pad_h, img);
            }
        }
    } else if (warp == 3) {
        // ===================================================== B producer (both CTAs): rows rank*64.. of every weight tile
        if (elect_one()) {
            uint32_t b_it = 0;
            for (int tile = pair_id; tile < p.total_tiles; tile += n_pairs) {
                const int ip = tile / tiles_per_img;
                for (int c = 0; c < 2; c++)
                    for (int tap = 0; tap < 9; tap++, b_it++) {
                        const int sb = b_it % kPairSB;
                        mbar_wait(&b_empty[sb], ((b_it / kPairSB) & 1) ^ 1);
                        const uint32_t full_l = mapa_u32(smem_u32(&b_full[sb]), 0);
                        if (leader) mbar_expect_tx(&b_full[sb], 2 * p.b_tx_bytes);
                        tma_load_2d_pair(b_smem + (size_t)sb * p.b_stage_bytes, &mapB, full_l, tap * p.Cin + c * KC,
                                         (int)rank * (N_TILE / 2));
                    }
                for (int j = 0; j < 2; j++)             // dynamic filters of image j of the pair, this CTA's half
                    for (int tap = 0; tap < 9; tap++, b_it++) {
                        const int sb = b_it % kPairSB;
                        mbar_wait(&b_empty[sb], ((b_it / kPairSB) & 1) ^ 1);
                        const uint32_t full_l = mapa_u32(smem_u32(&b_full[sb]), 0);
                        if (leader) mbar_expect_tx(&b_full[sb], 2 * p.b2_tx_bytes);
                        tma_load_2d_pair(b_smem + (size_t)sb * p.b_stage_bytes, &mapB2, full_l, tap * 16,
                                         (2 * ip + j) * N_TILE + (int)rank * (N_TILE / 2));
                    }
            }
        }
    } else if (warp == 1 && leader) {
        // ===================================================== MMA issuer (leader CTA, one thread)
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(256, N_TILE);
            const uint64_t desc0 = make_smem_desc<SWZ>(0, 0);
            const uint32_t desc_hi = (uint32_t)(desc0 >> 32);
            const uint32_t lo_flags = (uint32_t)desc0 & ~0x3FFFu;
            const uint32_t a_lo0 = lo_flags | ((smem_u32(a_smem) & 0x3FFFFu) >> 4);
            const uint32_t b_lo0 = lo_flags | ((smem_u32(b_smem) & 0x3FFFFu) >> 4);
            const uint32_t a_stage_lo = p.a_stage_bytes >> 4, b_stage_lo = p.b_stage_bytes >> 4;
            constexpr uint32_t BLK_LO = (128u * SWZ) >> 4;
            const uint32_t desc_hi32 = (uint32_t)(make_smem_desc<32>(0, 0) >> 32);
            const uint32_t a2_lo0 = lo_flags | ((smem_u32(a2_smem) & 0x3FFFFu) >> 4);
            uint32_t a_it = 0, b_it = 0, acc_it = 0, t_it = 0;
            PROF_DECL;
            for (int tile = pair_id; tile < p.total_tiles; tile += n_pairs, t_it++, acc_it++) {
                const int r = tile % tiles_per_img;
                const int tps = r % p.tiles_per_strip;
                const int q0 = tps * (NB * 128);
                const int soff = q0 - (q0 / p.Wp) * p.Wp;
                const int buf = acc_it & 1;
                PROF_LAP(3);
                mbar_wait(&acc_empty[buf], ((acc_it >> 1) & 1) ^ 1);
                PROF_LAP(0);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
                for (int c = 0; c < 2; c++, a_it++) {
                    const int sa = a_it & 1;
                    PROF_LAP(3);
                    mbar_wait(&a_full[sa], (a_it >> 1) & 1);
                    PROF_LAP(1);
                    tc_fence_after();
                    const uint32_t a_lo_tile = a_lo0 + sa * a_stage_lo + (uint32_t)soff * (SWZ >> 4);
#pragma unroll 1
                    for (int tap = 0; tap < 9; tap++, b_it++) {
                        const int sb = b_it % kPairSB;
                        PROF_LAP(3);
                        mbar_wait(&b_full[sb], (b_it / kPairSB) & 1);
                        PROF_LAP(2);
                        tc_fence_after();
                        const uint32_t a_lo = a_lo_tile + tap_lo_s[tap];
                        const uint32_t b_lo = b_lo0 + sb * b_stage_lo;
#pragma unroll
                        for (int blk = 0; blk < NB; blk++) {
#pragma unroll
                            for (int k = 0; k < KSTEPS; k++)
                                umma_bf16_lohi_pair(d_tmem + blk * N_TILE, a_lo + blk * BLK_LO + k * 2, b_lo + k * 2, desc_hi,
                                                    idesc, (c | tap | k) != 0);
                        }
                        umma_commit_pair(&b_empty[sb]);
                    }
                    umma_commit_pair(&a_empty[sa]);
                }
                // K-DYN extension, two passes: pass j = filters of image j x mask-patch buffer j (own patch in one
                // CTA, zeros in the other)
                mbar_wait(&a2_full, t_it & 1);
                tc_fence_after();
                for (int j = 0; j < 2; j++) {
                    const uint32_t a2_lo_tile = a2_lo0 + ((uint32_t)j * a2_bytes >> 4) + (uint32_t)soff * (32u >> 4);
#pragma unroll 1
                    for (int tap = 0; tap < 9; tap++, b_it++) {
                        const int sb = b_it % kPairSB;
                        mbar_wait(&b_full[sb], (b_it / kPairSB) & 1);
                        tc_fence_after();
                        const uint32_t a_lo = a2_lo_tile + (tap_lo_s[tap] * 32u) / SWZ;
                        const uint32_t b_lo = b_lo0 + sb * b_stage_lo;
#pragma unroll
                        for (int blk = 0; blk < NB; blk++)
                            umma_bf16_lohi_pair(d_tmem + blk * N_TILE, a_lo + blk * ((128u * 32u) >> 4), b_lo, desc_hi32, idesc, 1);
                        umma_commit_pair(&b_empty[sb]);
                    }
                }
                umma_commit_pair(&a2_empty);
                umma_commit_pair(&acc_full[buf]);
            }
            PROF_LAP(3);
            PROF_FLUSH(0, 4);
        }
    }
    } else if (warp >= 4 && warp < 12) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(DASR_LEAN_EPI_REGS));
        // ===================================================== epilogue warps (both CTAs): sean_epilogue on the own image
        const int ew = warp & 3;
        const int half = (warp - 4) >> 2;
        const int m = ew * 32 + lane;
        const int et = threadIdx.x - 128;
        uint32_t acc_it = 0;
        PROF_DECL;
        for (int tile = pair_id; tile < p.total_tiles; tile += n_pairs, acc_it++) {
            const int ip = tile / tiles_per_img;
            int r = tile - ip * tiles_per_img;
            const int strip = r / p.tiles_per_strip;
            const int tps = r - strip * p.tiles_per_strip;
            const int img = 2 * ip + (int)rank;
            const bool have = img < p.B;
            const int q0 = tps * (NB * 128);
            const int w0 = strip * p.Wt;
            const int buf = acc_it & 1;
            constexpr int NF = N_TILE / 2;
            SeanTileOps<N_TILE, NB> ops;              // every global operand of the tile is requested here
            sean_prefetch_tile<N_TILE, NB>(p, ops, img, q0, w0, m, half);
            asm volatile("bar.sync 1, 256;\n" ::: "memory");
            if (et < NF && have) {
                // fused double-InstanceNorm finalize of this image (see conv_halo_kernel)
                const float2* sp = reinterpret_cast<const float2*>(p.stats) + (size_t)img * p.nslots * NF + et;
                float a1 = 0.f, a2 = 0.f;
                for (int sl = 0; sl < p.nslots; sl += 8) {
                    float2 v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        v[u] = (sl + u < p.nslots) ? __ldg(sp + (size_t)(sl + u) * NF) : make_float2(0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        a1 += v[u].x;
                        a2 += v[u].y;
                    }
                }
                const float inv_hw = 1.f / (float)(p.H * p.W);
                const float mean = a1 * inv_hw;
                const float var = fmaxf(a2 * inv_hw - mean * mean, 0.f);
                const float eps = 1e-5f;
                const float r1 = rsqrtf(var + eps);
                const float r2 = rsqrtf(var * r1 * r1 + eps);
                norm_s[2 * et] = mean;
                norm_s[2 * et + 1] = r1 * r2;
                if (p.norm_out && tps == 0 && strip == 0) {
                    p.norm_out[((size_t)img * NF + et) * 2] = mean;
                    p.norm_out[((size_t)img * NF + et) * 2 + 1] = r1 * r2;
                    if (p.normk_out) {
                        const float a = var + eps, rr = var / a + eps;
                        p.normk_out[(size_t)img * NF + et] = 1.f / a + eps / (a * a * rr);
                    }
                }
            }
            asm volatile("bar.sync 1, 256;\n" ::: "memory");
            PROF_LAP(1);
            mbar_wait(&acc_full[buf], (acc_it >> 1) & 1);
            PROF_LAP(0);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + buf * ACC_COLS + (uint32_t(ew * 32) << 16);
            if (have && half * 16 < NF) sean_epilogue_tile<N_TILE, NB>(p, ops, t_acc, bias_s, norm_s, half);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[buf]), 0));
        }
        PROF_LAP(1);
#ifdef DASR_PROFILE
        if (threadIdx.x == 128) PROF_FLUSH(4, 2);
#endif
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // both CTAs are done with both TMEM halves and with each other's barriers
    PROF_KERNEL_END;
    if (warp == 2) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ trunk conv, CTA pairs
// The 64 -> 64 convolution in front of every InstanceNorm (DASR_EPI_STATS) on CTA pairs.  A single CTA issues
// M128 x N64 x K16 MMAs, which cost max(N/2, (128 + N)/4) = 48 tensor cycles for 32 cycles of math (the A operand's
// shared-memory reads bound it); a pair issues M256 x N64 in the same ~46 cycles (tools/pair_probe.cu): twice the
// pixels per dispatch.  Pairing as in sean_pair_kernel: the same tile position of two consecutive images; the weights
// (9 taps x 64 x 64) stay resident, each CTA holding rows rank*32..+31 of every tap (36 KB).  Statistics slots, store
// addressing and summation order are those of the single-CTA kernel (bit-identical partial sums).
// 16 epilogue warps (4 per TMEM lane quadrant, one 16-column chunk each at N = 64): with the MMA time halved the
// statistics epilogue became the critical path (in-kernel accounting: 5.3 k cycles per tile with 8 warps against 3.3 k
// of MMA); its cost is the latency chain TMEM load -> round -> store -> butterfly, so more warps shorten it.
#ifndef DASR_STATS_PAIR_EW
#define DASR_STATS_PAIR_EW 2
#endif
constexpr int kStatsPairEW = DASR_STATS_PAIR_EW;
constexpr int kStatsPairThreads = 128 + 128 * kStatsPairEW;
template <int N_TILE, int NB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kStatsPairThreads, 1)
stats_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ConvK p) {
    constexpr int SWZ = 128, KSTEPS = 4;
    constexpr int ACC_COLS = NB * N_TILE;
    constexpr int TMEM_COLS = 512;
    constexpr uint32_t B_TAP_BYTES = (N_TILE / 2) * 128;

    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[2], a_empty[2], b_full, acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float norm_s[512];         // scratch of the per-tile statistics reduction
    __shared__ __align__(16) float bias_s[N_TILE];
    __shared__ uint32_t tap_lo_s[9];

    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* a_smem = smem;
    uint8_t* b_smem = smem + 2 * (size_t)p.a_stage_bytes;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    PROF_KERNEL_BEGIN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 8 * kStatsPairEW);
        }
        mbar_init(&b_full, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
    }
    if (warp == 2) tmem_alloc_pair<TMEM_COLS>(&tmem_base_s);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int i = threadIdx.x; i < N_TILE; i += kStatsPairThreads) bias_s[i] = __ldg(p.bias + i);
    if (threadIdx.x < 9) {
        const int t = threadIdx.x / 3, u = threadIdx.x - t * 3;
        tap_lo_s[threadIdx.x] = (uint32_t)((t * p.Wp + u) * SWZ) >> 4;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int tiles_per_img = p.n_strips * p.tiles_per_strip;
    const int n_pairs = gridDim.x >> 1;
    const int pair_id = blockIdx.x >> 1;

    if (warp == 0) {
        // ===================================================== producer (both CTAs): half of the weights once, own patches
        if (elect_one()) {
            const uint32_t b_full_l = mapa_u32(smem_u32(&b_full), 0);
            if (leader) mbar_expect_tx(&b_full, 2 * 9 * B_TAP_BYTES);
            for (int tap = 0; tap < 9; tap++)
                tma_load_2d_pair(b_smem + (size_t)tap * B_TAP_BYTES, &mapB, b_full_l, tap * p.Cin, (int)rank * (N_TILE / 2));
            uint32_t a_it = 0;
            for (int tile = pair_id; tile < p.total_tiles; tile += n_pairs, a_it++) {
                const int ip = tile / tiles_per_img;
                int r = tile - ip * tiles_per_img;
                const int strip = r / p.tiles_per_strip;
                const int tps = r - strip * p.tiles_per_strip;
                const int img = 2 * ip + (int)rank;
                const int q0 = tps * (NB * 128);
                const int sa = a_it & 1;
                mbar_wait(&a_empty[sa], ((a_it >> 1) & 1) ^ 1);
                const uint32_t full_l = mapa_u32(smem_u32(&a_full[sa]), 0);
                if (leader) mbar_expect_tx(&a_full[sa], 2 * p.a_tx_bytes);
                tma_load_4d_pair(a_smem + (size_t)sa * p.a_stage_bytes, &mapA, full_l, 0, strip * p.Wt - p.pad_w,
                                 q0 / p.Wp - p.pad_h, img);
            }
        }
    } else if (warp == 1 && leader) {
        // ===================================================== MMA issuer (leader CTA, one thread)
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(256, N_TILE);
            const uint64_t desc0 = make_smem_desc<SWZ>(0, 0);
            const uint32_t desc_hi = (uint32_t)(desc0 >> 32);
            const uint32_t lo_flags = (uint32_t)desc0 & ~0x3FFFu;
            const uint32_t a_lo0 = lo_flags | ((smem_u32(a_smem) & 0x3FFFFu) >> 4);
            const uint32_t b_lo0 = lo_flags | ((smem_u32(b_smem) & 0x3FFFFu) >> 4);
            const uint32_t a_stage_lo = p.a_stage_bytes >> 4;
            constexpr uint32_t BLK_LO = (128u * SWZ) >> 4;
            uint32_t it = 0;
            PROF_DECL;
            mbar_wait(&b_full, 0);
            tc_fence_after();
            for (int tile = pair_id; tile < p.total_tiles; tile += n_pairs, it++) {
                const int tps = (tile % tiles_per_img) % p.tiles_per_strip;
                const int q0 = tps * (NB * 128);
                const int soff = q0 - (q0 / p.Wp) * p.Wp;
                const int buf = it & 1;
                PROF_LAP(3);
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                PROF_LAP(0);
                mbar_wait(&a_full[buf], (it >> 1) & 1);          // A stage index == accumulator buffer index
                PROF_LAP(1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
                const uint32_t a_lo_tile = a_lo0 + buf * a_stage_lo + (uint32_t)soff * (SWZ >> 4);
#pragma unroll 1
                for (int tap = 0; tap < 9; tap++) {
                    const uint32_t a_lo = a_lo_tile + tap_lo_s[tap];
                    const uint32_t b_lo = b_lo0 + tap * (B_TAP_BYTES >> 4);
#pragma unroll
                    for (int blk = 0; blk < NB; blk++) {
#pragma unroll
                        for (int k = 0; k < KSTEPS; k++)
                            umma_bf16_lohi_pair(d_tmem + blk * N_TILE, a_lo + blk * BLK_LO + k * 2, b_lo + k * 2, desc_hi, idesc,
                                                (tap | k) != 0);
                    }
                }
                umma_commit_pair(&a_empty[buf]);
                umma_commit_pair(&acc_full[buf]);
            }
            PROF_LAP(3);
            PROF_FLUSH(0, 4);
        }
    } else if (warp >= 4) {
        // ===================================================== epilogue warps (both CTAs): stats_epilogue on the own image
        const int ew = warp & 3;
        const int half = (warp - 4) >> 2;
        const int m = ew * 32 + lane;
        const int et = threadIdx.x - 128;
        uint32_t it = 0;
        PROF_DECL;
        for (int tile = pair_id; tile < p.total_tiles; tile += n_pairs, it++) {
            const int ip = tile / tiles_per_img;
            int r = tile - ip * tiles_per_img;
            const int strip = r / p.tiles_per_strip;
            const int tps = r - strip * p.tiles_per_strip;
            const int img = 2 * ip + (int)rank;
            const bool have = img < p.B;
            const int q0 = tps * (NB * 128);
            const int w0 = strip * p.Wt;
            const int buf = it & 1;
            PROF_LAP(1);
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            PROF_LAP(0);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + buf * ACC_COLS + (uint32_t(ew * 32) << 16);
            if (have) stats_epilogue<N_TILE, NB, false, kStatsPairEW, true>(p, t_acc, bias_s, norm_s, img, q0, w0, 0, m, half, ew, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[buf]), 0));
            // one statistics slot per (image, tile): the four row quadrants are added in a fixed order
            if (!DBG(p, 16)) asm volatile("bar.sync 1, %0;\n" ::"n"(128 * kStatsPairEW) : "memory");
            if (et < N_TILE && have && !DBG(p, 8)) {
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    a1 += norm_s[(q * N_TILE + et) * 2];
                    a2 += norm_s[(q * N_TILE + et) * 2 + 1];
                }
                const int slot = strip * p.tiles_per_strip + tps;
                float2* sp = reinterpret_cast<float2*>(p.stats) + ((size_t)img * p.nslots + slot) * p.Cout + et;
                *sp = make_float2(a1, a2);
            }
            if (!DBG(p, 16)) asm volatile("bar.sync 1, %0;\n" ::"n"(128 * kStatsPairEW) : "memory");
        }
        PROF_LAP(1);
#ifdef DASR_PROFILE
        if (threadIdx.x == 128) PROF_FLUSH(4, 2);
#endif
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    PROF_KERNEL_END;
    if (warp == 2) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
}
#endif  // DASR_CONV_PRECISE_TU

// ------------------------------------------------------------------------------------------------ host
// Widest strip of a frame one tile row covers (wider frames are cut into equal strips).  The trunk convolutions of a
// depth-guided block (STATS / SEAN epilogue -- the producer's statistics slots and the consumer's finalize must agree on
// the strips) use strips of at most 96 pixels, so that two A stages and the CTA-pair kernels fit in shared memory at
// W = 128 (x4: 4.90 -> 4.15 ms per 16-frame forward) and W = 240 (1080p frames: +1.6 %); everything else keeps 128
// (narrower strips only add halo columns there).  DASR_MAX_WT overrides both, DASR_TRUNK_WT the trunk width (measurements).
static int max_strip_width(int epi) {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DASR_MAX_WT");
        v = e ? atoi(e) : 0;
        if (v < 16 || v > 128) v = 0;
    }
    if (v) return v;
    static int trunk = -1;
    if (trunk < 0) {
        const char* e = getenv("DASR_TRUNK_WT");
        trunk = e ? atoi(e) : 96;
        if (trunk < 16 || trunk > 128) trunk = 96;
    }
    return (epi == DASR_EPI_STATS || epi == DASR_EPI_SEAN) ? trunk : 128;
}

// DASR_PDL=0 turns programmatic dependent launch off (A/B measurements)
static bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DASR_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

template <int SWZ, int N_TILE, int NB, int EPI, bool GEN = false, bool PREC = kPrecTU>
static int launch(const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mA2, const CUtensorMap& mB2,
                  const ConvK& k, size_t smem_bytes, cudaStream_t stream) {
    auto fn = conv_halo_kernel<SWZ, N_TILE, NB, EPI, GEN, PREC>;
    static bool configured[64] = {false};
    int dev = 0;
    DASR_CUDA_OK(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        DASR_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 10 * 1024));
        if (EPI == DASR_EPI_SEAN) {
            // setmaxnreg.inc blocks until the CTA's register pool (registers at launch x threads) can serve it: the
            // redistribution below must fit the pool ptxas actually gave this kernel, or the epilogue warps hang
            cudaFuncAttributes fa;
            DASR_CUDA_OK(cudaFuncGetAttributes(&fa, fn));
            const int threads = GEN ? 512 : kThreads;
            const int need = GEN ? (128 * 56 + 256 * 168 + 128 * 120)
                                 : (128 * DASR_LEAN_ROLE_REGS + 256 * DASR_LEAN_EPI_REGS);
            DASR_REQUIRE(fa.numRegs * threads >= need,
                         "SEAN convolution compiled with %d registers per thread: the setmaxnreg redistribution needs "
                         "%d registers per CTA (rebuild with the launch bounds of conv_igemm.cu)", fa.numRegs, need);
        }
        configured[dev & 63] = true;
    }
    int grid = k.total_tiles < num_sms() ? k.total_tiles : num_sms();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEN ? 512 : kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    DASR_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, mA, mB, mA2, mB2, k));
    DASR_LAUNCH_OK();
    return DASR_OK;
}

// only the (epilogue, N tile) pairs the network uses are instantiated
template <int SWZ, int NB>
static int dispatch_n(int epi, int n_tile, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mA2,
                      const CUtensorMap& mB2, const ConvK& k, size_t smem, cudaStream_t s) {
#define DASR_CASE(E, N) \
    if (epi == (E) && n_tile == (N)) return launch<SWZ, N, NB, E>(mA, mB, mA2, mB2, k, smem, s)
    DASR_CASE(DASR_EPI_STORE, 16);
    DASR_CASE(DASR_EPI_STORE, 32);
    DASR_CASE(DASR_EPI_STORE, 64);
    DASR_CASE(DASR_EPI_STORE, 128);
    DASR_CASE(DASR_EPI_STATS, 32);
    DASR_CASE(DASR_EPI_STATS, 64);
#ifndef DASR_CONV_PRECISE_TU
    if (epi == DASR_EPI_SEAN && n_tile == 128 && k.gen_depth) {
        if (SWZ == 128) return launch<128, 128, NB, DASR_EPI_SEAN, true>(mA, mB, mA2, mB2, k, smem, s);
        return fail(DASR_ERR_BAD_ARG, "the in-kernel actv generator needs Cin = 128");
    }
#endif
    DASR_CASE(DASR_EPI_SEAN, 64);
    DASR_CASE(DASR_EPI_SEAN, 128);
    DASR_CASE(DASR_EPI_SHUFFLE2, 64);
    DASR_CASE(DASR_EPI_SHUFFLE2, 128);
#ifndef DASR_CONV_PRECISE_TU
    DASR_CASE(DASR_EPI_NCHW_F32, 16);
#endif
#undef DASR_CASE
    return fail(DASR_ERR_BAD_ARG, "unsupported epilogue %d with N tile %d", epi, n_tile);
}

#ifndef DASR_CONV_PRECISE_TU
// DASR_SEAN_PAIR=0: keep the [gamma_o; beta_o] convolution on single CTAs (A/B measurements)
static bool sean_pair_enabled() { return pair_kernels_enabled(DASR_PAIR_SEAN); }

// sean_pair_kernel launch: k holds the single-CTA geometry (NB = 2, two A stages); the weight ring and the tile
// count are re-derived for pairs.  Returns 1 when the geometry does not fit (the caller uses the single-CTA kernel).
static int launch_sean_pair(const dasr_conv_desc* d, const dasr_conv_args* a, ConvK k, const CUtensorMap& mA,
                            const CUtensorMap& mA2, cudaStream_t stream, int* used) {
    constexpr int N_TILE = 128, NB = 2;
    *used = 0;
    const uint32_t a2_bytes = (k.a2_tx_bytes + 1023u) & ~1023u;
    k.b_tx_bytes = (N_TILE / 2) * 128;
    k.b_stage_bytes = k.b_tx_bytes;
    k.b2_tx_bytes = (N_TILE / 2) * 32;
    k.SA = 2;
    k.SB = kPairSB;
    k.b_resident = 0;
    k.a2_off = (uint32_t)(2 * (size_t)k.a_stage_bytes + (size_t)kPairSB * k.b_stage_bytes);
    const size_t smem_bytes = (size_t)k.a2_off + 2 * (size_t)a2_bytes + 1024;
    if (smem_bytes > 227 * 1024 - 10 * 1024) return DASR_OK;          // wide strips: single-CTA kernel
    const int pairs_of_images = (d->B + 1) / 2;
    k.total_tiles = pairs_of_images * k.n_strips * k.tiles_per_strip;
    CUtensorMap mB, mB2;
    {
        uint64_t dims[2] = {(uint64_t)k.taps * d->Cin, (uint64_t)N_TILE};
        uint64_t str[1] = {(uint64_t)k.taps * d->Cin * 2};
        uint32_t box[2] = {64, (uint32_t)(N_TILE / 2)};
        int rc = encode_tmap_bf16(&mB, a->w, 2, dims, str, box, 128);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)k.taps * 16, (uint64_t)d->B * N_TILE};
        uint64_t str[1] = {(uint64_t)k.taps * 16 * 2};
        uint32_t box[2] = {16, (uint32_t)(N_TILE / 2)};
        int rc = encode_tmap_bf16(&mB2, a->dyn_w, 2, dims, str, box, 32);
        if (rc) return rc;
    }
    auto fn = sean_pair_kernel<N_TILE, NB>;
    static bool configured[64] = {false};
    int dev = 0;
    DASR_CUDA_OK(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        DASR_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 10 * 1024));
        cudaFuncAttributes fa;
        DASR_CUDA_OK(cudaFuncGetAttributes(&fa, fn));
        DASR_REQUIRE(fa.numRegs * kThreads >= 128 * DASR_LEAN_ROLE_REGS + 256 * DASR_LEAN_EPI_REGS,
                     "SEAN pair kernel compiled with %d registers per thread: setmaxnreg cannot be served", fa.numRegs);
        configured[dev & 63] = true;
    }
    int n_pairs = num_sms() / 2;
    if (n_pairs > k.total_tiles) n_pairs = k.total_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * n_pairs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    DASR_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, mA, mB, mA2, mB2, k));
    DASR_LAUNCH_OK();
    *used = 1;
    return DASR_OK;
}

// stats_pair_kernel launch (the 64 -> 64 trunk convolution in front of an InstanceNorm on CTA pairs)
static int launch_stats_pair(const dasr_conv_desc* d, const dasr_conv_args* a, ConvK k, const CUtensorMap& mA,
                             cudaStream_t stream, int* used) {
    constexpr int N_TILE = 64, NB = 2;
    *used = 0;
    const size_t smem_bytes = 2 * (size_t)k.a_stage_bytes + 9 * (size_t)(N_TILE / 2) * 128 + 1024;
    if (smem_bytes > 227 * 1024 - 10 * 1024) return DASR_OK;
    k.SA = 2;
    k.total_tiles = ((d->B + 1) / 2) * k.n_strips * k.tiles_per_strip;
    CUtensorMap mB;
    {
        uint64_t dims[2] = {(uint64_t)k.taps * d->Cin, (uint64_t)N_TILE};
        uint64_t str[1] = {(uint64_t)k.taps * d->Cin * 2};
        uint32_t box[2] = {64, (uint32_t)(N_TILE / 2)};
        int rc = encode_tmap_bf16(&mB, a->w, 2, dims, str, box, 128);
        if (rc) return rc;
    }
    auto fn = stats_pair_kernel<N_TILE, NB>;
    static bool configured[64] = {false};
    int dev = 0;
    DASR_CUDA_OK(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        DASR_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 10 * 1024));
        configured[dev & 63] = true;
    }
    int n_pairs = num_sms() / 2;
    if (n_pairs > k.total_tiles) n_pairs = k.total_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * n_pairs);
    cfg.blockDim = dim3(kStatsPairThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    DASR_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, mA, mB, k));
    DASR_LAUNCH_OK();
    *used = 1;
    return DASR_OK;
}
#endif

// the kernel instantiations of this translation unit (plain bf16 storage, or -- conv_igemm_precise.cu -- the
// fp32-split PREC variants): called by dasr_conv_fwd with the finished kernel parameters
#ifdef DASR_CONV_PRECISE_TU
int conv_dispatch_precise(int epi, int n_tile, int NB, int SWZ, const CUtensorMap& mA, const CUtensorMap& mB,
                          const CUtensorMap& mA2, const CUtensorMap& mB2, const ConvK& k, size_t smem_bytes,
                          cudaStream_t stream) {
#else
int conv_dispatch_precise(int epi, int n_tile, int NB, int SWZ, const CUtensorMap& mA, const CUtensorMap& mB,
                          const CUtensorMap& mA2, const CUtensorMap& mB2, const ConvK& k, size_t smem_bytes,
                          cudaStream_t stream);
static int conv_dispatch(int epi, int n_tile, int NB, int SWZ, const CUtensorMap& mA, const CUtensorMap& mB,
                         const CUtensorMap& mA2, const CUtensorMap& mB2, const ConvK& k, size_t smem_bytes,
                         cudaStream_t stream) {
#endif
    if (NB == 4) {       // Cin = 32 -> 64-byte swizzle
        if (n_tile == 32) return launch<64, 32, 4, DASR_EPI_STORE>(mA, mB, mA2, mB2, k, smem_bytes, stream);
        return launch<64, 16, 4, DASR_EPI_STORE>(mA, mB, mA2, mB2, k, smem_bytes, stream);
    }
    if (SWZ == 128) return dispatch_n<128, 2>(epi, n_tile, mA, mB, mA2, mB2, k, smem_bytes, stream);
    return dispatch_n<64, 2>(epi, n_tile, mA, mB, mA2, mB2, k, smem_bytes, stream);
}

}  // namespace dasr

#ifndef DASR_CONV_PRECISE_TU
using namespace dasr;

#ifdef DASR_PROFILE
extern "C" int dasr_prof_set(int dbg) {
    g_host_dbg = dbg;
    return DASR_OK;
}
extern "C" int dasr_prof_read(unsigned long long* host_out, int reset) {
    DASR_CUDA_OK(cudaDeviceSynchronize());
    DASR_CUDA_OK(cudaMemcpyFromSymbol(host_out, g_prof, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0};
        DASR_CUDA_OK(cudaMemcpyToSymbol(g_prof, z, sizeof z));
    }
    return DASR_OK;
}
#endif

// 1 if the SEAN conv of this geometry can generate its A operand in-kernel (two A stages fit), else 0
extern "C" int dasr_conv_gen_ok(int H, int W) {
    const int NB = 2, max_wt = max_strip_width(DASR_EPI_SEAN);
    const int n_strips = (W + max_wt - 1) / max_wt;
    const int Wt = (W + n_strips - 1) / n_strips;
    const int Wp = Wt + 2;
    const int RB = (Wp - 1 + NB * 128 + 2 * Wp + 2 + Wp - 1) / Wp;
    const size_t a_stage = (((size_t)RB * Wp * 128) + 1023u) & ~(size_t)1023u;
    const size_t a2 = (((size_t)RB * Wp * 32) + 1023u) & ~(size_t)1023u;
    const size_t gen = (((size_t)(RB + 3) * (Wp + 2) * 4) + 1023u) & ~(size_t)1023u;
    (void)H;
    return (2 * a_stage + 2 * 16384 + a2 + gen <= 216 * 1024) ? 1 : 0;
}

extern "C" int dasr_conv_stats_slots(const dasr_conv_desc* d) {
    DASR_REQUIRE(d && d->W > 0 && d->H > 0 && (d->ks == 1 || d->ks == 3 || d->ks == 9), "bad descriptor");
    const int NB = 2, max_wt = max_strip_width(DASR_EPI_STATS);
    const int n_strips = (d->W + max_wt - 1) / max_wt;
    const int Wt = (d->W + n_strips - 1) / n_strips;
    const int Wp = Wt + (d->kw > 0 ? d->kw : d->ks) - 1;
    const int span = (d->H - 1) * Wp + Wt;
    const int blocks = (span + 127) / 128;
    const int tiles_per_strip = (blocks + NB - 1) / NB;
    return n_strips * tiles_per_strip;
}

extern "C" int dasr_conv_fwd(const dasr_conv_desc* d, const dasr_conv_args* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DASR_REQUIRE(d && a, "null descriptor");
    DASR_REQUIRE(d->ks == 1 || d->ks == 3 || d->ks == 9, "ks must be 1, 3 or 9 (got %d)", d->ks);
    DASR_REQUIRE(d->Cin % 32 == 0 && d->Cin >= 32, "Cin must be a multiple of 32 (got %d)", d->Cin);
    DASR_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cout > 0, "bad shape");
    const bool gen = a->gen_depth != nullptr;
    DASR_REQUIRE((a->x || gen) && a->w && a->bias && a->out, "null tensor pointer");
    if (gen)
        DASR_REQUIRE(a->gen_w && a->gen_b && d->epi == DASR_EPI_SEAN && d->Cin == 128 && d->Cout == 128 && d->ks == 3 && a->dyn_x,
                     "the in-kernel actv generator needs gen_w / gen_b, the SEAN epilogue with the K-DYN extension and "
                     "Cin = Cout = 128");

    const int npl = planes();
    ConvK k;
    memset(&k, 0, sizeof k);
    k.npl = npl;
    {
        const PlaneTerms t = plane_terms(npl, npl);
        k.n_terms = t.n;
        for (int i = 0; i < t.n; i++) { k.ta[i] = t.a[i]; k.tb[i] = t.b[i]; }
    }
    if (npl > 1)
        DASR_REQUIRE(!gen && !a->gb_s && !a->resid_f32 && !a->out_aux_f32 && !a->norm && d->epi != DASR_EPI_NCHW_F32,
                     "fp32-split planes: the in-kernel actv generator, gb_s, the fp32 residual stream, an external "
                     "norm buffer and the NCHW fp32 epilogue are not available");
    k.B = d->B; k.H = d->H; k.W = d->W; k.Cin = d->Cin; k.Cout = d->Cout;
    k.kh = d->ks; k.kw = d->kw > 0 ? d->kw : d->ks;
    k.pad_h = k.kh / 2; k.pad_w = k.kw / 2; k.taps = k.kh * k.kw;
    DASR_REQUIRE(k.kw == 1 || k.kw == 3 || k.kw == 9, "kw must be 1, 3 or 9 (got %d)", k.kw);
    k.epi = d->epi; k.act = d->act; k.subsample = d->subsample ? d->subsample : 1;
    k.clamp01 = d->clamp01; k.inner_relu = d->inner_relu;

    const int SWZ = (d->Cin % 64 == 0) ? 128 : 64;
    const int KC = SWZ / 2;
    k.nch = d->Cin / KC;

    // N tiling
    int n_tile;
    if (d->epi == DASR_EPI_NCHW_F32) {
        DASR_REQUIRE(d->Cout <= 16, "NCHW fp32 epilogue supports Cout <= 16");
        n_tile = 16;
        k.ntn = 1;
    } else {
        n_tile = d->Cout >= 128 ? 128 : d->Cout;
        // Cout = 288 (the x3 tail, 32 * 3^2 channels in front of PixelShuffle(3)): nine N tiles of 32
        if (d->Cout > 128 && d->Cout % 128 != 0) n_tile = (d->Cout % 64 == 0) ? 64 : 32;
        DASR_REQUIRE(n_tile == 16 || n_tile == 32 || n_tile == 64 || n_tile == 128,
                     "Cout must be 16, 32, 64 or a multiple of 128 (got %d)", d->Cout);
        DASR_REQUIRE(d->Cout % n_tile == 0, "Cout %d not a multiple of the N tile %d", d->Cout, n_tile);
        k.ntn = d->Cout / n_tile;
    }
    if (d->epi == DASR_EPI_SEAN) {
        DASR_REQUIRE(k.ntn == 1 && n_tile >= 32, "SEAN epilogue needs Cout = 2*nf in {64,128}");
        DASR_REQUIRE(a->y && (a->norm || a->stats), "SEAN epilogue needs y and norm (or the statistics partials)");
    }
    if (d->epi == DASR_EPI_STATS) DASR_REQUIRE(a->stats, "STATS epilogue needs a stats buffer");
    k.n_bias = k.ntn * n_tile;
    DASR_REQUIRE(k.n_bias <= kMaxBias, "Cout %d too large (max %d)", d->Cout, kMaxBias);
    if (d->epi == DASR_EPI_SHUFFLE2) DASR_REQUIRE(d->Cout % 64 == 0, "shuffle needs Cout multiple of 64");

    // strips / tiles.  Narrow STORE convolutions (N tile <= 32: the 32-channel tail) use 512-pixel tiles (NB = 4):
    // the halo amplification of the A patch drops from 2.5x to 1.8x at W = 256 and the per-tile overheads halve;
    // TMEM still double-buffers (2 x 4 x 32 columns).
    // Measured: 32->32 at 256x256 160 -> 139 us; with Cin = 64 (K = 576) it is slower, so only Cin = 32 layers use it.
    const int NB = (n_tile <= 32 && d->Cin == 32 && d->epi == DASR_EPI_STORE && d->H * d->W >= 128 * 128) ? 4 : 2;
    const int max_wt = max_strip_width(d->epi);
    k.n_strips = (d->W + max_wt - 1) / max_wt;
    k.Wt = (d->W + k.n_strips - 1) / k.n_strips;
    k.Wp = k.Wt + k.kw - 1;
    k.wp_magic = (unsigned)((1ull << 32) / (unsigned)k.Wp) + 1u;
    const int span = (d->H - 1) * k.Wp + k.Wt;  // padded-flat positions that contain valid outputs
    const int blocks = (span + 127) / 128;
    k.tiles_per_strip = (blocks + NB - 1) / NB;
    k.total_tiles = d->B * k.n_strips * k.tiles_per_strip * k.ntn;
    k.RB = (k.Wp - 1 + NB * 128 + (k.kh - 1) * k.Wp + (k.kw - 1) + k.Wp - 1) / k.Wp;
    DASR_REQUIRE(k.RB <= 256 && k.Wp <= 256, "TMA box too large");

    k.a_tx_bytes = (uint32_t)k.RB * k.Wp * SWZ;
    k.a_stage_bytes = (k.a_tx_bytes + 1023u) & ~1023u;
    k.b_tx_bytes = (uint32_t)n_tile * SWZ;
    k.b_stage_bytes = (k.b_tx_bytes + 1023u) & ~1023u;

    size_t budget = 216 * 1024;         // + 1 KB alignment slack + ~9.5 KB static <= 227 KB
    size_t a2_bytes = 0;
    if (a->dyn_x || a->dyn_w) {
        DASR_REQUIRE(a->dyn_x && a->dyn_w && d->epi == DASR_EPI_SEAN && SWZ == 128 && !a->gb_s,
                     "the K-DYN extension needs dyn_x and dyn_w, the SEAN epilogue, Cin %% 64 == 0 and no gb_s");
        k.dyn = 1;
        k.a2_tx_bytes = (uint32_t)k.RB * k.Wp * 32;
        k.b2_tx_bytes = (uint32_t)n_tile * 32;
        a2_bytes = (k.a2_tx_bytes + 1023u) & ~(size_t)1023u;
        budget -= a2_bytes;
    }
    size_t gen_bytes = 0;
    if (gen) {
        k.gen_depth = a->gen_depth;
        k.gen_w = a->gen_w;
        k.gen_b = a->gen_b;
        gen_bytes = (((size_t)(k.RB + 3) * (k.Wp + 2) * sizeof(float)) + 1023u) & ~(size_t)1023u;   // + one row of slack
        budget -= gen_bytes;
    }
    const size_t all_b = (size_t)k.nch * k.taps * k.b_stage_bytes;
    k.SA = 2;
    if ((size_t)k.SA * k.a_stage_bytes + 2 * (size_t)k.b_stage_bytes > budget) k.SA = 1;
    DASR_REQUIRE((size_t)k.SA * k.a_stage_bytes + 2 * (size_t)k.b_stage_bytes <= budget,
                 "A tile (%u bytes) does not fit in shared memory", k.a_stage_bytes);
    // per-image weights always stream; so does a GEMM with the K-DYN extension (its dynamic filters share the weight ring:
    // with Cin = 64 -- a SEAN instance of a 32-channel block -- the static weights alone would fit)
    const bool may_reside = (k.ntn == 1) && (d->w_img_rows == 0) && npl == 1 && !k.dyn;
    if (may_reside && k.nch * k.taps <= kMaxBStages &&
        (size_t)k.SA * k.a_stage_bytes + all_b <= budget) {
        k.b_resident = 1;
        k.SB = k.nch * k.taps;
    } else if (may_reside && k.nch * k.taps <= kMaxBStages && (size_t)k.a_stage_bytes + all_b <= budget) {
        k.b_resident = 1;
        k.SA = 1;
        k.SB = k.nch * k.taps;
    } else {
        k.b_resident = 0;
        size_t room = budget - (size_t)k.SA * k.a_stage_bytes;
        int sb = (int)(room / k.b_stage_bytes);
        if (sb > 8) sb = 8;
        if (sb > k.nch * k.taps) sb = k.nch * k.taps;
        k.SB = sb;
        DASR_REQUIRE(sb >= 1, "not enough shared memory for the weight ring");
    }
    if (k.dyn) DASR_REQUIRE(!k.b_resident, "the K-DYN extension streams its weights (resident mode not supported)");
    k.a2_off = (uint32_t)((size_t)k.SA * k.a_stage_bytes + (size_t)k.SB * k.b_stage_bytes);
    k.gen_off = k.a2_off + (uint32_t)a2_bytes;
    if (gen) DASR_REQUIRE(k.SA == 2 && k.nch == 2, "the in-kernel actv generator needs two A stages (one per channel chunk); "
                          "use dasr_conv_gen_ok() to test a geometry");
    const size_t smem_bytes = (size_t)k.SA * k.a_stage_bytes + (size_t)k.SB * k.b_stage_bytes + a2_bytes + gen_bytes + 1024;

    // outputs
    k.Ho = d->H; k.Wo = d->W;
    if (k.subsample == 2) { k.Ho = (d->H + 1) / 2; k.Wo = (d->W + 1) / 2; }
    if (d->epi == DASR_EPI_SHUFFLE2) { k.Ho = 2 * d->H; k.Wo = 2 * d->W; }
    k.bias = a->bias;
    k.out = (__nv_bfloat16*)a->out;
    k.out_f32 = (float*)a->out;
    k.resid = (const __nv_bfloat16*)a->resid;
    k.stats = a->stats;
    k.y = (const __nv_bfloat16*)a->y;
    k.norm = a->norm;
    k.gb_s = (const __nv_bfloat16*)a->gb_s;
    k.actmask = (const __nv_bfloat16*)a->actmask;
    k.mask_slope = d->mask_slope;
    k.gamma_out = (__nv_bfloat16*)a->gamma_out;
    k.resid_f32 = a->resid_f32;
    k.out_aux_f32 = a->out_aux_f32;
    k.nslots = k.n_strips * k.tiles_per_strip;
    k.norm_out = a->norm_out;
    k.normk_out = a->normk_out;
#ifdef DASR_PROFILE
    k.dbg = g_host_dbg;
#endif
    if (d->epi == DASR_EPI_STATS)
        DASR_REQUIRE(n_tile <= 64, "STATS epilogue supports Cout tiles up to 64 (shared-memory partials)");
    k.w_img_rows = d->w_img_rows;
    k.unshuffle = d->unshuffle;
    // plane strides: every plane of an act tensor is a full copy of its logical shape
    k.ps_out = (size_t)d->B * k.Ho * k.Wo * (d->epi == DASR_EPI_SEAN ? d->Cout / 2 : d->epi == DASR_EPI_SHUFFLE2 ? d->Cout / 4 : d->Cout);
    k.w_plane_rows = k.ntn * n_tile * (d->w_img_rows ? d->B : 1);
    k.dynw_plane_rows = d->B * n_tile;
    if (npl > 1) DASR_REQUIRE(k.ntn * n_tile == d->Cout, "fp32-split planes need Cout (%d) to be a multiple of the N tile", d->Cout);
    if (d->unshuffle)
        DASR_REQUIRE(d->unshuffle == 2 && d->epi == DASR_EPI_STORE && k.subsample == 1 && d->H % 2 == 0 && d->W % 2 == 0,
                     "unshuffle store: factor 2, EPI_STORE, stride 1 and even frame sizes only");

    // tensor maps
    CUtensorMap mA, mB;
    {
        uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B * npl};   // planes follow each other
        uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
        uint32_t box[4] = {(uint32_t)KC, (uint32_t)k.Wp, (uint32_t)k.RB, 1};
        int rc = encode_tmap_bf16(&mA, gen ? a->out : a->x, 4, dims, str, box, SWZ);     // unused when generating
        if (rc) return rc;
    }
    {
        const uint64_t ktot = (uint64_t)k.taps * d->Cin;
        const uint64_t rows = (uint64_t)k.ntn * n_tile * (d->w_img_rows ? d->B : 1) * npl;
        if (d->w_img_rows) DASR_REQUIRE(d->w_img_rows == k.ntn * n_tile, "per-image weights: w_img_rows must equal the padded Cout");
        uint64_t dims[2] = {ktot, rows};
        uint64_t str[1] = {ktot * 2};
        uint32_t box[2] = {(uint32_t)KC, (uint32_t)n_tile};
        int rc = encode_tmap_bf16(&mB, a->w, 2, dims, str, box, SWZ);
        if (rc) return rc;
    }
    CUtensorMap mA2 = mA, mB2 = mB;
    if (k.dyn) {
        {
            uint64_t dims[4] = {16, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
            uint64_t str[3] = {32, (uint64_t)d->W * 32, (uint64_t)d->H * d->W * 32};
            uint32_t box[4] = {16, (uint32_t)k.Wp, (uint32_t)k.RB, 1};
            int rc = encode_tmap_bf16(&mA2, a->dyn_x, 4, dims, str, box, 32);
            if (rc) return rc;
        }
        {
            uint64_t dims[2] = {(uint64_t)k.taps * 16, (uint64_t)d->B * n_tile * npl};
            uint64_t str[1] = {(uint64_t)k.taps * 16 * 2};
            uint32_t box[2] = {16, (uint32_t)n_tile};
            int rc = encode_tmap_bf16(&mB2, a->dyn_w, 2, dims, str, box, 32);
            if (rc) return rc;
        }
    }
    if (npl > 1) return conv_dispatch_precise(d->epi, n_tile, NB, SWZ, mA, mB, mA2, mB2, k, smem_bytes, stream);
    // the [gamma_o; beta_o] convolution of SEAN on CTA pairs (cta_group::2) when there are at least two images and
    // enough pair tiles to occupy every SM pair
    if (d->epi == DASR_EPI_SEAN && n_tile == 128 && d->Cin == 128 && d->ks == 3 && k.kw == 3 && k.dyn && !gen && !a->norm &&
        !a->gb_s && !a->resid_f32 && !a->out_aux_f32 && d->B >= 2 && sean_pair_enabled() &&
        ((d->B + 1) / 2) * k.n_strips * k.tiles_per_strip >= num_sms() / 2) {
        int used = 0;
        int rc = launch_sean_pair(d, a, k, mA, mA2, stream, &used);
        if (rc || used) return rc;
    }
    if (d->epi == DASR_EPI_STATS && n_tile == 64 && d->Cin == 64 && d->Cout == 64 && d->ks == 3 && k.kw == 3 && NB == 2 &&
        k.subsample == 1 && !d->w_img_rows && d->B >= 2 && pair_kernels_enabled(DASR_PAIR_STATS) &&
        ((d->B + 1) / 2) * k.n_strips * k.tiles_per_strip >= num_sms() / 2) {
        int used = 0;
        int rc = launch_stats_pair(d, a, k, mA, stream, &used);
        if (rc || used) return rc;
    }
    return conv_dispatch(d->epi, n_tile, NB, SWZ, mA, mB, mA2, mB2, k, smem_bytes, stream);
}
#endif  // DASR_CONV_PRECISE_TU

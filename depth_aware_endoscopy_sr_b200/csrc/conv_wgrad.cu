// K-CONV-WGRAD: weight gradient of a stride-1 "same" convolution on the 5th-gen tensor cores.
//
//   dWp[o][tap*Cin + i] += sum_{b,h,w} dY[b,h,w,o] * X[b, h+t-pad_h, w+u-pad_w, i]        (bf16 x bf16 -> fp32)
//
// (the packed GEMM-B layout of conv_igemm.cu, so dasr_unpack_grads can undo weight-norm / alpha folding /
// PixelShuffle permutation row by row).  It is the backward of every nn.Conv2d call site cited in conv_igemm.cu
// (autograd's conv2d weight gradient in the reference, codes/models/F_model_depthCond.py:191).
//
// Both GEMM operands have the reduction (pixel) dimension OUTERMOST in memory ([pixel][channel], channels
// contiguous), i.e. they are MN-major.  tcgen05 consumes MN-major operands directly (instruction-descriptor
// bits 15/16, canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- verified on B200 with
// tools/umma_probe.cu test 6, including arbitrary row shifts of the start address), so NO transposition pass
// exists anywhere: a TMA box of [rows = pixels][64 channels] IS the operand tile, the 64-channel blocks of a
// wider operand sit LBO bytes apart, and the tap (t,u) of the B operand is the same X patch viewed
// (t*Wp + u) rows further down -- the halo-reuse trick of the forward kernel applied to the K dimension.
//
// Work split: grid.x = split-K over (image, row-tile, strip) tiles; grid.y = (tap-group, Cin chunk) column
// slices of at most 512 fp32 TMEM columns; grid.z = blocks of 128 (64) output channels.  Each CTA keeps its
// [M x cols] slice in TMEM across ALL its K tiles and flushes once with fp32 red.global.add.
// Warp roles: w0 TMA producer (multi-stage ring), w1 MMA issuer (one elected thread), w2 TMEM allocator,
// w4..7 final flush.
#include "dasr_internal.h"
#include <stdlib.h>
#include <string.h>
#include "sm100_ptx.cuh"

namespace dasr {

constexpr int kWgThreads = 256;
constexpr int kWgMaxStages = 4;

struct WgK {
    int B, H, W, Cout, Cin;
    int kh, kw, pad_h, pad_w;
    int TR, Wt, Wp, PR;                 // rows per K tile, strip width, patch width / rows
    int n_strips, n_rowtiles, ktiles_total;
    int taps_per_slice, n_tapgroups, n_cchunks, Nc;
    int Mb, m_valid_last;               // MMA M (64 / 128); grid.z blocks
    int swz_a, swz_b, cb_a, cb_b;       // swizzle bytes and channels per smem block of dY / X
    int a_blocks, b_blocks;             // TMA loads per stage
    uint32_t a_blk_bytes, b_blk_bytes, stage_bytes, a_lbo, b_lbo;
    uint32_t a_tx, b_tx;                // bytes one TMA box really transfers (blocks are padded to 1 KB)
    int stages;
    // stacked taps (Cin fits ONE smem block): the column blocks of the B operand of one MMA are the SAME X block
    // viewed stack_lbo bytes further down each (the next horizontal / vertical tap), so stack_g taps are one MMA of
    // N = stack_g * Nc instead of stack_g MMAs of N = Nc (tools/umma_probe.cu test 7)
    int stack_g;
    uint32_t stack_lbo;
    float* dw;
    int ldw;
    // per-image mode (K-DYN backward): grid.x = B * ksplit_i, every CTA reduces over the pixels of ONE image and
    // flushes columns c < n_valid of tap tp to  dw[img*dw_img_stride + (c*taps + tap)*Cout + o]  ("bin-major":
    // the [k][tap][channel] layout of the style table)
    int per_image, ksplit_i, n_valid;
    long long dw_img_stride;
    const int* skip_flag;               // optional device int: the whole kernel is a no-op when *skip_flag != 0
    int col_first, col_count;           // flush window inside every tap's Nc columns (default: all of them)
    int tma_flush;                      // flush through cp.reduce.async.bulk.tensor (mapW) instead of per-lane red
    // bias gradient db[o] += sum_pixels dY[p][o], fused: the CTAs of slice 0 issue one more MMA per K step whose B
    // operand is a constant block of ones (16 columns) -- the column sum of the A operand already in shared memory
    float* db;
    uint32_t ones_off;                  // byte offset of the 2 KB block of bf16 ones behind the pipeline stages
    int dbg;                            // -DDASR_PROFILE ablation knob (env DASR_WG_DBG): 1 no flush, 2 no MMA, 4 no TMA
    // fp32-split planes (dasr_internal.h): the reduction runs over the cross terms (ta[i], tb[i]) of the dY / X planes
    // -- planes are extra "images" of the same tensor map
    int n_terms;
    unsigned char ta[6], tb[6];
};

#ifdef DASR_PROFILE
#define WG_DBG(p, bit) ((p).dbg & (bit))
#else
#define WG_DBG(p, bit) 0
#endif

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapX,
                  const __grid_constant__ CUtensorMap mapW, const WgK p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full[kWgMaxStages], empty[kWgMaxStages], acc_full;
    __shared__ uint32_t tmem_base_s;

    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(&acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapY);
        tma_prefetch_desc(&mapX);
    }
    if (warp == 2) tmem_alloc<512>(&tmem_base_s);
    if (p.db) {
        uint32_t* ones = reinterpret_cast<uint32_t*>(smem + p.ones_off);
        for (int i = threadIdx.x; i < 512; i += kWgThreads) ones[i] = 0x3F803F80u;        // bf16 1.0 x 2
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                      // read by tcgen05.mma
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // programmatic dependent launch: the set-up above overlaps the tail of the previous kernel (see conv_igemm.cu)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // this CTA's slice
    const int tapgroup = blockIdx.y / p.n_cchunks;
    const int cchunk = blockIdx.y - tapgroup * p.n_cchunks;
    const int tap0 = tapgroup * p.taps_per_slice;
    const int ntaps = min(p.taps_per_slice, p.kh * p.kw - tap0);
    const int t_first = tap0 / p.kw;
    const int mblk = blockIdx.z;
    const int ca0 = mblk * p.Mb;             // first dY channel of this M block
    const int cb0 = cchunk * p.Nc;           // first X channel of this column slice
    const int tiles_per_img = p.n_rowtiles * p.n_strips;
    const uint32_t a_bytes_stage = p.a_blocks * p.a_blk_bytes;
    int kt_begin = blockIdx.x, kt_end = p.ktiles_total, kt_step = gridDim.x, my_img = 0;
    if (p.per_image) {
        my_img = blockIdx.x / p.ksplit_i;
        kt_begin = my_img * tiles_per_img + (int)(blockIdx.x % p.ksplit_i);
        kt_end = (my_img + 1) * tiles_per_img;
        kt_step = p.ksplit_i;
    }
    if (p.skip_flag && *p.skip_flag != 0) kt_end = kt_begin;     // uniform over the grid: nothing to do
    const int n_my = kt_end > kt_begin ? (kt_end - kt_begin + kt_step - 1) / kt_step : 0;
    const bool do_bias = p.db != nullptr && blockIdx.y == 0;      // tap group 0, channel chunk 0

    if (warp == 0) {
        if (elect_one()) {
            uint32_t it = 0;
            for (int term = 0; term < p.n_terms; term++)
            for (int kt = kt_begin; kt < kt_end; kt += kt_step, it++) {
                const int img = kt / tiles_per_img;
                const int r = kt - img * tiles_per_img;
                const int rt = r / p.n_strips, strip = r - rt * p.n_strips;
                const int s = it % p.stages;
                mbar_wait(&empty[s], ((it / p.stages) & 1) ^ 1);
                if (WG_DBG(p, 4)) {
                    mbar_arrive(&full[s]);
                    continue;
                }
                mbar_expect_tx(&full[s], p.a_blocks * p.a_tx + p.b_blocks * p.b_tx);
                uint8_t* sa = smem + (size_t)s * p.stage_bytes;
                uint8_t* sb = sa + a_bytes_stage;
                for (int j = 0; j < p.a_blocks; j++)
                    tma_load_4d(sa + (size_t)j * p.a_blk_bytes, &mapY, &full[s], ca0 + j * p.cb_a, strip * p.Wt, rt * p.TR,
                                img + (int)p.ta[term] * p.B);
                for (int j = 0; j < p.b_blocks; j++)
                    tma_load_4d(sb + (size_t)j * p.b_blk_bytes, &mapX, &full[s], cb0 + j * p.cb_b, strip * p.Wt - p.pad_w,
                                rt * p.TR + t_first - p.pad_h, img + (int)p.tb[term] * p.B);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // MN-major operands: idesc bits 15 (A) and 16 (B); descriptor = addr>>4 | LBO>>4 <<16 | SBO>>4 <<32 | ...
            const uint32_t mn_major = (1u << 15) | (1u << 16);
            const int g = p.stack_g;                              // taps per MMA
            const int g_last = ntaps - ((ntaps - 1) / g) * g;     // taps of the last group of the slice
            const uint32_t idesc_full = make_idesc_bf16(p.Mb, g * p.Nc) | mn_major;
            const uint32_t idesc_last = make_idesc_bf16(p.Mb, g_last * p.Nc) | mn_major;
            const uint32_t hi_a = ((8u * p.swz_a) >> 4) | (1u << 14) | ((p.swz_a == 128 ? 2u : 4u) << 29);
            const uint32_t hi_b = ((8u * p.swz_b) >> 4) | (1u << 14) | ((p.swz_b == 128 ? 2u : 4u) << 29);
            const uint32_t lo_a_flags = ((p.a_lbo >> 4) & 0x3FFFu) << 16;
            const uint32_t lo_b_flags = (((g > 1 ? p.stack_lbo : p.b_lbo) >> 4) & 0x3FFFu) << 16;
            const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
            const uint32_t idesc_ones = make_idesc_bf16(p.Mb, 16) | mn_major;
            const uint32_t ones_lo = smem_lo + (p.ones_off >> 4);        // every element is 1: LBO irrelevant
            const int ksteps = p.Wt >> 4;
            const uint32_t a_kstep = (16u * p.swz_a) >> 4, b_kstep = (16u * p.swz_b) >> 4;
            uint32_t it = 0;
            for (int term = 0; term < p.n_terms; term++)
            for (int kt = kt_begin; kt < kt_end; kt += kt_step, it++) {
                const int s = it % p.stages;
                mbar_wait(&full[s], (it / p.stages) & 1);
                tc_fence_after();
                const uint32_t sa_lo = smem_lo + ((uint32_t)s * p.stage_bytes >> 4);
                const uint32_t sb_lo = sa_lo + (a_bytes_stage >> 4);
#pragma unroll 1
                for (int r = 0; r < (WG_DBG(p, 2) ? 0 : p.TR); r++) {
                    const uint32_t a_row = sa_lo + (((uint32_t)(r * p.Wt) * p.swz_a) >> 4);
#pragma unroll 1
                    for (int tp = 0; tp < ntaps; tp += g) {
                        const int tap = tap0 + tp;
                        const int t = tap / p.kw, u = tap - t * p.kw;
                        const uint32_t b_row = sb_lo + (((uint32_t)((r + t - t_first) * p.Wp + u) * p.swz_b) >> 4);
                        const uint32_t d = tmem_base + tp * p.Nc;
                        const uint32_t idesc = (tp + g >= ntaps) ? idesc_last : idesc_full;
                        const uint32_t acc0 = (it | r) != 0;
#pragma unroll 4
                        for (int ks = 0; ks < ksteps; ks++) {
                            // umma_bf16_lohi with distinct high words for A and B
                            const uint32_t a_lo = (a_row + ks * a_kstep) | lo_a_flags;
                            const uint32_t b_lo = (b_row + ks * b_kstep) | lo_b_flags;
                            const uint32_t accum = acc0 | (ks != 0);
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                                "setp.ne.b32 p, %6, 0;\n\t"
                                "mov.b64 da, {%1, %3};\n\t"
                                "mov.b64 db, {%2, %4};\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(d),
                                "r"(a_lo), "r"(b_lo), "r"(hi_a), "r"(hi_b), "r"(idesc), "r"(accum)
                                : "memory");
                        }
                    }
                    if (do_bias && p.tb[term] == 0) {      // every dY plane exactly once
                        const uint32_t d = tmem_base + ntaps * p.Nc;
                        const uint32_t acc0 = (it | r) != 0;
#pragma unroll 4
                        for (int ks = 0; ks < ksteps; ks++) {
                            const uint32_t a_lo = (a_row + ks * a_kstep) | lo_a_flags;
                            const uint32_t accum = acc0 | (ks != 0);
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                                "setp.ne.b32 p, %6, 0;\n\t"
                                "mov.b64 da, {%1, %3};\n\t"
                                "mov.b64 db, {%2, %4};\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(d),
                                "r"(a_lo), "r"(ones_lo), "r"(hi_a), "r"(hi_b), "r"(idesc_ones), "r"(accum)
                                : "memory");
                        }
                    }
                }
                umma_commit(&empty[s]);
            }
            if (n_my > 0) umma_commit(&acc_full);
        }
    } else if (warp >= 4) {
        // ===================================================== flush: TMEM -> red.global.add.f32
        if (n_my > 0 && !WG_DBG(p, 1)) {
            mbar_wait(&acc_full, 0);
            tc_fence_after();
            const int ew = warp & 3;
            // accumulator row of this lane: M = 128 -> lane index; M = 64 -> lanes 0..15 of each 32-lane quadrant
            int m;
            bool row_ok;
            if (p.Mb == 128) {
                m = ew * 32 + lane;
                row_ok = true;
            } else {
                m = ew * 16 + lane;
                row_ok = lane < 16;
            }
            const int o = ca0 + m;
            row_ok = row_ok && (o < p.Cout);
            const uint32_t t_row = tmem_base + (uint32_t(ew * 32) << 16);
            if (do_bias) {
                uint32_t v[16];
                tmem_ld16(t_row + ntaps * p.Nc, v);
                tmem_ld_wait();
                if (row_ok) atomicAdd(p.db + o, __uint_as_float(v[0]));
            }
            if (p.tma_flush) {
                // Bulk tensor reductions: the slice goes TMEM -> registers -> 128-byte-swizzled [128 rows][32 floats]
                // tiles in the (now idle) pipeline stages -> ONE cp.reduce.async.bulk.tensor (fp32 add, performed at
                // L2) per tile, issued by one thread.  The per-lane red flush it replaces cost ~1.3 SM cycles per
                // lane-request: 10 of the 35 us of a 128 -> 128 weight gradient at B = 16.
                constexpr int kFlushBufs = 8;
                // M = 64: the accumulator rows sit in lanes 0..15 of every quadrant (tile of 64 rows)
                const int row = (p.Mb == 128) ? ew * 32 + lane : ew * 16 + lane;
                const bool writer = (p.Mb == 128) || lane < 16;
                const int ngrp = p.Nc >> 5;
                const bool issuer = (threadIdx.x == 128);
                int ti = 0;
                for (int tp = 0; tp < ntaps; tp++)
                    for (int cg = 0; cg < ngrp; cg++, ti++) {
                        uint8_t* tile = smem + (size_t)(ti % kFlushBufs) * 16384;
                        if (ti >= kFlushBufs) {      // the reduction that last read this buffer has read it
                            if (issuer) asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
                            asm volatile("bar.sync 1, 128;" ::: "memory");
                        }
                        uint32_t v[32];
                        tmem_ld32(t_row + tp * p.Nc + cg * 32, v);
                        tmem_ld_wait();
                        const uint32_t trow = smem_u32(tile) + row * 128;
                        if (writer)
#pragma unroll
                        for (int c = 0; c < 8; c++)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(trow + ((c ^ (row & 7)) << 4)),
                                         "r"(v[4 * c]), "r"(v[4 * c + 1]), "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                                         : "memory");
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        if (issuer) {
                            const int col = (tap0 + tp) * p.Cin + cb0 + cg * 32;
                            asm volatile(
                                "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                                ::"l"(&mapW), "r"(col), "r"(ca0), "r"(smem_u32(tile))
                                : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            } else
            for (int tp = 0; tp < ntaps; tp++) {
                float* dst = p.dw + (size_t)o * p.ldw + (size_t)(tap0 + tp) * p.Cin + cb0;
                float* dst_img = p.dw + (size_t)my_img * p.dw_img_stride + (size_t)(tap0 + tp) * p.Cout + o;
                const int ncols = p.per_image ? min(p.Nc, p.n_valid - cb0) : p.Nc;
                for (int c0 = 0; c0 < ncols; c0 += 16) {
                    if (c0 + 16 <= p.col_first || c0 >= p.col_first + p.col_count) continue;
                    uint32_t v[16];
                    tmem_ld16(t_row + tp * p.Nc + c0, v);
                    tmem_ld_wait();
                    {
                        if (p.per_image) {
                            // a lane owns one channel (row) and 16 bins (columns); the destination is bin-major, so
                            // a 4x4 transpose inside every lane quad (two shfl.bfly rounds) gives each lane ONE bin
                            // of FOUR consecutive channels: one 128-bit reduction instead of four scalar ones
                            // (the flush was 90 scalar red instructions per warp, ~8 of the kernel's 17 us)
                            const int taps = p.kh * p.kw;
                            const int qi = lane & 3;
#pragma unroll
                            for (int g4 = 0; g4 < 16; g4 += 4) {
                                if (cb0 + c0 + g4 >= p.n_valid) break;       // warp-uniform
                                float x0 = __uint_as_float(v[g4]), x1 = __uint_as_float(v[g4 + 1]);
                                float x2 = __uint_as_float(v[g4 + 2]), x3 = __uint_as_float(v[g4 + 3]);
                                {
                                    const bool odd = qi & 1;
                                    const float r0 = __shfl_xor_sync(0xffffffffu, odd ? x0 : x1, 1);
                                    const float r1 = __shfl_xor_sync(0xffffffffu, odd ? x2 : x3, 1);
                                    if (odd) { x0 = r0; x2 = r1; } else { x1 = r0; x3 = r1; }
                                }
                                {
                                    const bool hi = qi & 2;
                                    const float r0 = __shfl_xor_sync(0xffffffffu, hi ? x0 : x2, 2);
                                    const float r1 = __shfl_xor_sync(0xffffffffu, hi ? x1 : x3, 2);
                                    if (hi) { x0 = r0; x1 = r1; } else { x2 = r0; x3 = r1; }
                                }
                                const int bin = cb0 + c0 + g4 + qi;
                                if (row_ok && bin < p.n_valid)
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                                                     dst_img - qi + (size_t)bin * taps * p.Cout),
                                                 "f"(x0), "f"(x1), "f"(x2), "f"(x3)
                                                 : "memory");
                            }
                        } else if (row_ok) {
                            // 128-bit vector reductions (REDG.ADD.F32x4): a lane owns one row of dW, so every
                            // scalar red was its own L2 request (7.2 M per launch, 80 % of the kernel time)
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                if (c0 + j >= p.col_first && c0 + j < p.col_first + p.col_count)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j),
                                             "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                             "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                             : "memory");
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace dasr

using namespace dasr;

struct WgOpts {
    float* db = nullptr;                    // fused bias gradient (or NULL)
    int per_image = 0, n_valid = 0;
    int x_planes = 0;                       // planes of the X operand (0 = planes(); 1 for the exact auxiliary tensor)
    int col_first = 0, col_count = 0;       // flush window (multiples of 4); 0 / 0 = every column
    long long dw_img_stride = 0;
    const int* skip_flag = nullptr;
};

static int wgrad_launch(const dasr_wgrad_desc* d, const void* dy, const void* x, float* dw, const WgOpts& o,
                        void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DASR_REQUIRE(d && dy && x && dw, "null argument");
    DASR_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0, "bad shape");
    DASR_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->kh * d->kw <= 81, "bad kernel size");
    DASR_REQUIRE(d->Cin % 32 == 0 && d->Cout % 32 == 0, "channels must be multiples of 32 (Cout %d, Cin %d)", d->Cout, d->Cin);
    DASR_REQUIRE(((uintptr_t)dw & 15) == 0, "dw must be 16-byte aligned (vector reductions)");
    DASR_REQUIRE(!o.per_image || (o.dw_img_stride % 4 == 0 && d->Cout % 4 == 0), "per-image gradient slices must be 16-byte aligned");

    WgK k;
    memset(&k, 0, sizeof k);
    k.B = d->B; k.H = d->H; k.W = d->W; k.Cout = d->Cout; k.Cin = d->Cin;
    k.kh = d->kh; k.kw = d->kw; k.pad_h = d->kh / 2; k.pad_w = d->kw / 2;
    k.dw = dw;
    k.ldw = d->kh * d->kw * d->Cin;
    const int npl_y = planes(), npl_x = o.x_planes > 0 ? o.x_planes : planes();
    {
        const PlaneTerms t = plane_terms(npl_y, npl_x);
        k.n_terms = t.n;
        for (int i = 0; i < t.n; i++) { k.ta[i] = t.a[i]; k.tb[i] = t.b[i]; }
    }

    // operand blocks
    k.swz_a = (d->Cout % 64 == 0) ? 128 : 64;
    k.swz_b = (d->Cin % 64 == 0) ? 128 : 64;
    k.cb_a = k.swz_a / 2;
    k.cb_b = k.swz_b / 2;
    // M blocks of dY channels
    k.Mb = d->Cout >= 128 ? 128 : 64;
    const int n_mblocks = (d->Cout + k.Mb - 1) / k.Mb;
    // a partial last M block (Cout = 288) reads zero rows: TMA zero-fills channel blocks beyond the tensor, and the
    // flush skips rows >= Cout
    DASR_REQUIRE(d->Cout % k.Mb == 0 || d->Cout < k.Mb || d->Cout % 32 == 0, "Cout %d not supported", d->Cout);
    const int m_rows = d->Cout < k.Mb ? d->Cout : k.Mb;   // valid dY channels per M block
    k.a_blocks = m_rows / k.cb_a;
    // column slices: Nc X-channels per MMA, taps_per_slice taps per CTA, <= 512 TMEM columns
    k.Nc = d->Cin > 128 ? 128 : d->Cin;
    DASR_REQUIRE(d->Cin % k.Nc == 0, "Cin %d not supported", d->Cin);
    k.n_cchunks = d->Cin / k.Nc;
    k.b_blocks = k.Nc / k.cb_b;
    k.taps_per_slice = d->kw;                               // one kernel row per slice
    if (k.taps_per_slice * k.Nc > 512) k.taps_per_slice = 512 / k.Nc;
    if (d->kw == 1) {                                        // vertical-only kernels: all taps in one slice if they fit
        k.taps_per_slice = d->kh;
        if (k.taps_per_slice * k.Nc > 512) k.taps_per_slice = 512 / k.Nc;
    }
    DASR_REQUIRE(k.taps_per_slice >= 1, "slice does not fit TMEM");
    if (d->kw > 1) DASR_REQUIRE(k.taps_per_slice == d->kw, "kernel row does not fit one TMEM slice (kw %d, Cin chunk %d)", d->kw, k.Nc);
    // stacked taps: one MMA covers a whole kernel row (or up to 256 / Nc vertical taps) when X is one smem block
    k.stack_g = 1;
    k.stack_lbo = 0;
    {
        const char* e = getenv("DASR_WG_STACK");
        const bool allow = !(e && e[0] == '0');
        if (allow && k.b_blocks == 1) {
            if (d->kw > 1 && d->kw * k.Nc <= 256) {
                k.stack_g = d->kw;
                k.stack_lbo = (uint32_t)k.swz_b;                     // next pixel of the patch row
                // the whole kernel in one CTA when its columns fit TMEM: dY and X are read once instead of kh times
                if (d->kh * d->kw * k.Nc <= 512) k.taps_per_slice = d->kh * d->kw;
            } else if (d->kw == 1 && k.taps_per_slice > 1) {
                k.stack_g = 256 / k.Nc < k.taps_per_slice ? 256 / k.Nc : k.taps_per_slice;
                // stack_lbo = Wp * swz_b, set below once the strip width is known
            }
        }
    }
    k.n_tapgroups = (d->kh * d->kw + k.taps_per_slice - 1) / k.taps_per_slice;
    // extra patch rows spanned by one slice
    const int dt = (d->kw > 1) ? (k.taps_per_slice / d->kw - 1) : (k.taps_per_slice - 1);

    // K tiles
    const int max_wt = 128;
    k.n_strips = (d->W + max_wt - 1) / max_wt;
    k.Wt = (((d->W + k.n_strips - 1) / k.n_strips) + 15) & ~15;
    k.n_strips = (d->W + k.Wt - 1) / k.Wt;
    k.Wp = k.Wt + d->kw - 1;
    DASR_REQUIRE(k.Wp <= 256, "strip too wide");
    if (d->kw == 1 && k.stack_g > 1) {
        k.stack_lbo = (uint32_t)k.Wp * k.swz_b;                      // next row of the patch
        if ((k.stack_lbo >> 4) > 0x3FFFu) k.stack_g = 1;
    }
    const size_t budget = 200 * 1024;
    int TR = 8;
    for (;; TR >>= 1) {
        k.TR = TR;
        k.PR = TR + dt;
        k.a_blk_bytes = (uint32_t)TR * k.Wt * k.swz_a;
        k.b_blk_bytes = (uint32_t)k.PR * k.Wp * k.swz_b;
        const uint32_t a_al = (k.a_blocks * k.a_blk_bytes + 1023u) & ~1023u;
        k.stage_bytes = a_al + ((k.b_blocks * k.b_blk_bytes + 1023u) & ~1023u);
        if (2 * (size_t)k.stage_bytes <= budget || TR == 1) break;
    }
    DASR_REQUIRE(2 * (size_t)k.stage_bytes <= budget && k.PR <= 256, "K tile does not fit in shared memory");
    // blocks inside one operand must each start 1024-byte aligned for the swizzle pattern to match TMA's
    k.a_blk_bytes = (k.a_blk_bytes + 1023u) & ~1023u;
    k.b_blk_bytes = (k.b_blk_bytes + 1023u) & ~1023u;
    {
        const uint32_t a_al = k.a_blocks * k.a_blk_bytes;
        k.stage_bytes = a_al + k.b_blocks * k.b_blk_bytes;
    }
    DASR_REQUIRE(2 * (size_t)k.stage_bytes <= budget + 16 * 1024, "K tile does not fit in shared memory");
    k.stages = (int)(budget / k.stage_bytes);
    if (k.stages > kWgMaxStages) k.stages = kWgMaxStages;
    if (k.stages < 2) k.stages = 2;
    // LBO: distance between channel blocks; a lone / duplicated block (Cout = 32 with M = 64) uses LBO = 0
    k.a_lbo = (k.a_blocks * k.cb_a >= k.Mb) ? k.a_blk_bytes : 0;
    k.b_lbo = k.b_blk_bytes;
    k.n_rowtiles = (d->H + k.TR - 1) / k.TR;
    k.ktiles_total = d->B * k.n_rowtiles * k.n_strips;

    const int slices = k.n_tapgroups * k.n_cchunks * n_mblocks;
    int ksplit = (2 * num_sms()) / slices;      // a CTA owns all 512 TMEM columns -> one CTA per SM; 2 waves
    if (ksplit > num_sms() / slices && slices <= num_sms()) ksplit = num_sms() / slices;
    // ksplit_div = d: d times fewer split-K CTAs per launch (each reduces d times more pixels and the launch flushes
    // d times fewer partial dW slices): fewer SM-microseconds per gradient at a longer latency, which is the better
    // trade when the gradient runs beside the data-gradient chain on a side stream
    const int ks_div = d->ksplit_div > 1 ? d->ksplit_div : 1;
    ksplit = (ksplit + ks_div - 1) / ks_div;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > k.ktiles_total) ksplit = k.ktiles_total;
    k.skip_flag = o.skip_flag;
    k.col_first = o.col_count ? o.col_first : 0;
    k.col_count = o.col_count ? o.col_count : k.Nc;
    DASR_REQUIRE(k.col_first % 4 == 0 && k.col_count % 4 == 0, "flush window must be a multiple of 4 columns");
#ifdef DASR_PROFILE
    if (const char* e = getenv("DASR_WG_DBG")) k.dbg = atoi(e);
#endif
    if (o.per_image) {
        const int tiles_per_img = k.n_rowtiles * k.n_strips;
        int ks_i = num_sms() / (slices * d->B * ks_div);
        if (ks_i < 1) ks_i = 1;
        if (ks_i > tiles_per_img) ks_i = tiles_per_img;
        k.per_image = 1;
        k.ksplit_i = ks_i;
        k.n_valid = o.n_valid;
        k.dw_img_stride = o.dw_img_stride;
        ksplit = d->B * ks_i;
    }

    CUtensorMap mY, mX;
    {
        uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B * npl_y};
        uint64_t str[3] = {(uint64_t)d->Cout * 2, (uint64_t)d->W * d->Cout * 2, (uint64_t)d->H * d->W * d->Cout * 2};
        uint32_t box[4] = {(uint32_t)k.cb_a, (uint32_t)k.Wt, (uint32_t)k.TR, 1};
        int rc = encode_tmap_bf16(&mY, dy, 4, dims, str, box, k.swz_a);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B * npl_x};
        uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
        uint32_t box[4] = {(uint32_t)k.cb_b, (uint32_t)k.Wp, (uint32_t)k.PR, 1};
        int rc = encode_tmap_bf16(&mX, x, 4, dims, str, box, k.swz_b);
        if (rc) return rc;
    }
    // flush through bulk tensor reductions when whole 32-column groups are flushed and eight 16 KB tiles fit the
    // pipeline stages
    CUtensorMap mW = mY;
    {
        const char* e = getenv("DASR_WG_TMA_FLUSH");
        const bool allow = !(e && e[0] == '0');
        k.tma_flush = allow && !o.per_image && k.Nc % 32 == 0 && k.col_first == 0 &&
                      k.col_count == k.Nc && (size_t)k.stages * k.stage_bytes >= 8 * 16384 && (k.ldw % 4) == 0;
        if (k.tma_flush) {
            uint64_t dims[2] = {(uint64_t)k.ldw, (uint64_t)d->Cout};
            uint64_t str[1] = {(uint64_t)k.ldw * 4};
            uint32_t box[2] = {32, (uint32_t)k.Mb};
            int rc = encode_tmap_f32(&mW, dw, 2, dims, str, box, 128);
            if (rc) return rc;
        }
    }
    k.a_tx = (uint32_t)k.TR * k.Wt * k.swz_a;
    k.b_tx = (uint32_t)k.PR * k.Wp * k.swz_b;
    static bool configured[64] = {false};
    int dev = 0;
    DASR_CUDA_OK(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        DASR_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured[dev & 63] = true;
    }
    k.db = nullptr;
    k.ones_off = 0;
    if (o.db) {
        DASR_REQUIRE(!o.per_image && k.taps_per_slice * k.Nc + 16 <= 512,
                     "fused bias gradient: no TMEM column left (taps %d x %d columns)", k.taps_per_slice, k.Nc);
        k.db = o.db;
        k.ones_off = (uint32_t)(((size_t)k.stages * k.stage_bytes + 1023) & ~(size_t)1023);
    }
    const size_t smem_bytes = (size_t)k.stages * k.stage_bytes + 1024 + (o.db ? 3072 : 0);
    DASR_REQUIRE(smem_bytes <= 220 * 1024, "shared memory budget exceeded (%zu)", smem_bytes);
    dim3 grid(ksplit, k.n_tapgroups * k.n_cchunks, n_mblocks);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(kWgThreads);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        const char* e = getenv("DASR_PDL");
        cfg.numAttrs = (e && e[0] == '0') ? 0 : 1;
        DASR_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_wgrad_kernel, mY, mX, mW, k));
    }
    DASR_LAUNCH_OK();
    return DASR_OK;
}

extern "C" int dasr_conv_wgrad(const dasr_wgrad_desc* d, const void* dy, const void* x, float* dw, float* db,
                               void* stream) {
    WgOpts o;
    o.db = db;
    return wgrad_launch(d, dy, x, dw, o, stream);
}

// K-DYN backward on the tensor cores: per image, dT[b][k][tap][c] += sum_p dgb[b,p,c] * onehot[b, p+tap-1, k]
// -- the weight gradient of the dynamic 3x3 convolution of the one-hot depth-mask image (channels 0..K-1 of the
// auxiliary tensor built by dasr_build_aux) with the per-image filters.  No-op when *flag != 0 (masks not one-hot:
// dasr_dynconv_bwd's exact path runs instead).
extern "C" int dasr_dynconv_bwd_tc(const void* dgb, const void* aux, const int32_t* flag, float* dT, int B, int K, int H,
                                   int W, int nf2, void* stream) {
    DASR_REQUIRE(dgb && aux && dT, "null pointer");
    DASR_REQUIRE(K >= 1 && K <= 16, "K-DYN backward: at most 16 depth masks (got %d)", K);
    dasr_wgrad_desc d;
    d.B = B; d.H = H; d.W = W; d.Cout = nf2; d.Cin = DASR_AUX_CH; d.kh = 3; d.kw = 3; d.ksplit_div = 0;
    WgOpts o;
    o.per_image = 1;
    o.n_valid = K;
    o.dw_img_stride = (long long)K * 9 * nf2;
    o.skip_flag = flag;
    o.x_planes = 1;
    return wgrad_launch(&d, dgb, aux, dT, o, stream);
}

// mlp_mask (1 -> 2nf, 3x3) weight / bias gradient on the tensor cores: dasr_conv_wgrad of dA against the auxiliary
// tensor, whose channels DASR_AUX_DEPTH_HI / _LO hold the depth split into two bf16 parts (hi + lo carries 16
// mantissa bits) and DASR_AUX_ONE is 1 (its centre tap is the bias gradient).  scratch: fp32 [C][9*DASR_AUX_CH],
// zeroed by the caller.  dW[c][tap] += , db[c] += .
namespace dasr {
__global__ void actv_bwd_gather_kernel(const float* __restrict__ scr, float* __restrict__ dW, float* __restrict__ db,
                                       int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * 9) return;
    const int c = i / 9, tap = i - c * 9;
    const float* row = scr + (size_t)c * 9 * DASR_AUX_CH + tap * DASR_AUX_CH;
    dW[i] += row[DASR_AUX_DEPTH_HI] + (row[DASR_AUX_DEPTH_LO] + row[DASR_AUX_DEPTH_LO2]);
    if (tap == 4) db[c] += row[DASR_AUX_ONE];
}
}  // namespace dasr

extern "C" int dasr_actv_bwd_tc(const void* dA, const void* aux, float* scratch, float* dW, float* db, int B, int H,
                                int W, int C, void* stream) {
    DASR_REQUIRE(dA && aux && scratch && dW && db, "null pointer");
    dasr_wgrad_desc d;
    d.B = B; d.H = H; d.W = W; d.Cout = C; d.Cin = DASR_AUX_CH; d.kh = 3; d.kw = 3; d.ksplit_div = 0;
    // only the depth (hi, lo, lo2) and the constant-one channel are read back: flush those four columns per tap
    static_assert(DASR_AUX_DEPTH_HI % 4 == 0 && DASR_AUX_DEPTH_LO2 == DASR_AUX_DEPTH_HI + 3 &&
                      DASR_AUX_ONE > DASR_AUX_DEPTH_HI && DASR_AUX_ONE < DASR_AUX_DEPTH_LO2,
                  "the four auxiliary channels of the mlp_mask gradient must share one 16-byte group");
    WgOpts o;
    o.col_first = DASR_AUX_DEPTH_HI;
    o.col_count = 4;
    o.x_planes = 1;
    int rc = wgrad_launch(&d, dA, aux, scratch, o, stream);
    if (rc) return rc;
    actv_bwd_gather_kernel<<<(C * 9 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(scratch, dW, db, C);
    DASR_LAUNCH_OK();
    return DASR_OK;
}

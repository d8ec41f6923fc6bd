// Internal helpers shared by the translation units of libdasr_b200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/dasr.h"

namespace dasr {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define DASR_CUDA_OK(expr)                                                                  \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return ::dasr::fail(DASR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,              \
                                cudaGetErrorString(e__), __FILE__, __LINE__);               \
    } while (0)

#define DASR_LAUNCH_OK()                                                                    \
    do {                                                                                    \
        ::dasr::count_launch();                                                             \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess)                                                             \
            return ::dasr::fail(DASR_ERR_CUDA, "kernel launch failed: %s (%s:%d)",          \
                                cudaGetErrorString(e__), __FILE__, __LINE__);               \
    } while (0)

#define DASR_REQUIRE(cond, ...)                                                             \
    do {                                                                                    \
        if (!(cond)) return ::dasr::fail(DASR_ERR_BAD_ARG, __VA_ARGS__);                    \
    } while (0)

// cuTensorMapEncodeTiled for a bf16 tensor (rank <= 4). swizzle_bytes in {32, 64, 128}.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);
// the same for an fp32 tensor (destination of the bulk tensor reductions of conv_wgrad.cu)
int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

int num_sms();
// the CTA-pair (cta_group::2) kernels that may be used (dasr_set_sean_pair / DASR_SEAN_PAIR): bit mask
#define DASR_PAIR_SEAN 1
#define DASR_PAIR_STATS 2
#define DASR_PAIR_OUT9 4
#ifndef DASR_PAIR_DEFAULT
#define DASR_PAIR_DEFAULT 7      /* all three; measured through bench.py (two alternating runs each): mask 7 5.68 ms per step, mask 5 (no trunk pair kernel) 5.70 ms, trunk conv 34 vs 37 us per launch; equal in the power-bound steady state */
#endif
bool pair_kernels_enabled(int which);
void count_launch();   // bumps the counter behind dasr_launch_count()

// ---------------------------------------------------------------------------------------------- fp32-split planes
// Precise mode (dasr_set_planes(3), tests only): every "act" tensor and every packed weight matrix is stored as
// kMaxPlanes bf16 planes [plane][elements] whose sum IS the fp32 value (hi = bf16(x), mid = bf16(x - hi),
// lo = bf16(x - hi - mid): exact for normal fp32 numbers).  The tensor-core kernels accumulate the cross terms
// (i, j) with i + j < planes of the operand planes into the same fp32 TMEM accumulators; the epilogues and the
// memory-bound kernels read the plane sum and write the split.  Same kernels, same schedule, fp32-class arithmetic.
// planes() == 1 is the product configuration (plain bf16 storage): every loop below then runs exactly once.
constexpr int kMaxPlanes = 3;
int planes();
struct PlaneTerms {
    int n;
    unsigned char a[6], b[6];
};
// cross terms of an (npl_a planes) x (npl_b planes) product that a `planes()`-plane result keeps
inline PlaneTerms plane_terms(int npl_a, int npl_b) {
    PlaneTerms t;
    t.n = 0;
    const int lim = npl_a > npl_b ? npl_a : npl_b;
    for (int s = 0; s < lim; s++)
        for (int i = 0; i <= s; i++) {
            const int j = s - i;
            if (i < npl_a && j < npl_b) {
                t.a[t.n] = (unsigned char)i;
                t.b[t.n] = (unsigned char)j;
                t.n++;
            }
        }
    return t;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint4 pl_pack8(const float* f) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}
__device__ __forceinline__ void pl_unpack8(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
// 8 consecutive values of an act tensor: sum of its planes (lowest plane first: the sum is then exact).
// p: plane 0 as uint4, i: uint4 index, ps8: plane stride in uint4 units.
__device__ __forceinline__ void pl_load8(const uint4* p, size_t i, size_t ps8, int npl, float* f) {
    pl_unpack8(__ldg(p + i + (size_t)(npl - 1) * ps8), f);
    for (int k = npl - 2; k >= 0; k--) {
        float t[8];
        pl_unpack8(__ldg(p + i + (size_t)k * ps8), t);
#pragma unroll
        for (int j = 0; j < 8; j++) f[j] += t[j];
    }
}
// the inverse: split f into npl planes (f is destroyed)
__device__ __forceinline__ void pl_store8(uint4* p, size_t i, size_t ps8, int npl, float* f) {
    uint4 u = pl_pack8(f);
    p[i] = u;
    for (int k = 1; k < npl; k++) {
        float t[8];
        pl_unpack8(u, t);
#pragma unroll
        for (int j = 0; j < 8; j++) f[j] -= t[j];
        u = pl_pack8(f);
        p[i + (size_t)k * ps8] = u;
    }
}
// scalar forms (element index / element plane stride)
__device__ __forceinline__ float pl_load1(const __nv_bfloat16* p, size_t i, size_t ps, int npl) {
    float f = __bfloat162float(p[i + (size_t)(npl - 1) * ps]);
    for (int k = npl - 2; k >= 0; k--) f += __bfloat162float(p[i + (size_t)k * ps]);
    return f;
}
__device__ __forceinline__ void pl_store1(__nv_bfloat16* p, size_t i, size_t ps, int npl, float f) {
    for (int k = 0; k < npl; k++) {
        const __nv_bfloat16 h = __float2bfloat16(f);
        p[i + (size_t)k * ps] = h;
        f -= __bfloat162float(h);
    }
}
#endif

}  // namespace dasr

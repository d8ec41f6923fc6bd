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
void count_launch();   // bumps the counter behind dasr_launch_count()

}  // namespace dasr

// Error plumbing, device check and TMA tensor-map encoding for libdasr_b200.so.
#include "dasr_internal.h"
#include <atomic>
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace dasr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

static int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

static int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(DASR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t d[5], s[5];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; i++) {
        d[i] = dims[i];
        b[i] = box[i];
        e[i] = 1;
    }
    for (int i = 0; i < rank - 1; i++) s[i] = strides_bytes[i];
    CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                  : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = enc(out, dtype, (cuuint32_t)rank, const_cast<void*>(base),
                     d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(DASR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, box0 %u)",
                    (int)r, rank, box[0]);
    return DASR_OK;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_planes{1};
int planes() { return g_planes.load(std::memory_order_relaxed); }

// CTA-pair (tcgen05.mma.cta_group::2) kernels: -1 = environment DASR_SEAN_PAIR (default on), 0 / 1 = dasr_set_sean_pair
// a bit mask: 1 = SEAN convolution, 2 = trunk (STATS) convolution, 4 = conv_out9
static std::atomic<int> g_pair_mode{-1};
bool pair_kernels_enabled(int which) {
    int m = g_pair_mode.load(std::memory_order_relaxed);
    if (m < 0) {
        static int v = -1;
        if (v < 0) {
            const char* e = getenv("DASR_SEAN_PAIR");
            v = e ? atoi(e) : DASR_PAIR_DEFAULT;
            if (e && e[0] == '1' && e[1] == 0) v = 7;
        }
        m = v;
    }
    return (m & which) != 0;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace dasr

extern "C" const char* dasr_last_error(void) { return dasr::g_err; }
extern "C" int dasr_version(void) { return 100; }
extern "C" int64_t dasr_launch_count(void) { return (int64_t)dasr::g_launches.load(std::memory_order_relaxed); }

extern "C" int dasr_set_planes(int n) {
    if (n != 1 && n != 2 && n != dasr::kMaxPlanes)
        return dasr::fail(DASR_ERR_BAD_ARG, "planes must be 1 (bf16 storage), 2 or %d (fp32 split); got %d", dasr::kMaxPlanes, n);
    dasr::g_planes.store(n);
    return DASR_OK;
}
extern "C" int dasr_get_planes(void) { return dasr::planes(); }
extern "C" int dasr_set_sean_pair(int on) {
    dasr::g_pair_mode.store(on < 0 ? -1 : (on == 1 ? 7 : on));
    return DASR_OK;
}

extern "C" int dasr_check_device(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess)
        return dasr::fail(DASR_ERR_CUDA, "no CUDA device: libdasr_b200 has no CPU fallback");
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10)
        return dasr::fail(DASR_ERR_ARCH, "device is sm_%d%d; libdasr_b200 is built for sm_100a only", major,
                          minor);
    return DASR_OK;
}

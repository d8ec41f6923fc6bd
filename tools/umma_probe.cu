// Hardware probe for the tcgen05/TMA encodings the conv kernels rely on (run on a B200 via gpurun).
// Each test prints PASS/FAIL lines; raw dumps go to gpurun_out/ for offline decoding.
//   probe 1 : plain 128xNx64 bf16 GEMM tile, SW128, N in {16,64,128,256}
//   probe 2 : A start address shifted by r rows inside the swizzle atom, base_offset 0 vs (addr>>7)&7
//   probe 3 : 4D TMA (NHWC) with negative / OOB coordinates -> zero fill + row order
//   probe 4 : 4D TMA with elementStrides = 2 on W and H
//   probe 5 : SW64 (32-channel rows) GEMM + row shifts
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../depth_aware_endoscopy_sr_b200/csrc/sm100_ptx.cuh"

using namespace dasr;

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    for (int i = 0; i < 4000000; i++)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                  \
        }                                                                             \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static void init_driver() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) {
        printf("no cuTensorMapEncodeTiled\n");
        exit(2);
    }
    g_encode = (EncodeTiledFn)fn;
}

static CUtensorMap make_map(void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                            const uint32_t* box, const uint32_t* estr, CUtensorMapSwizzle swz) {
    CUtensorMap m;
    cuuint64_t d[5], s[5];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; i++) {
        d[i] = dims[i];
        b[i] = box[i];
        e[i] = estr ? estr[i] : 1;
    }
    for (int i = 0; i < rank - 1; i++) s[i] = strides_b[i];
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, d, s, b, e,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
        exit(2);
    }
    return m;
}

// ------------------------------------------------------------------ GEMM probe kernel
// A: [rowsA][KB] bf16 (KB = SWZ/2 elements per row), B: [N][KB]. D[m][n] = sum_k A[r+m][k]*B[n][k]
template <int SWZ>
__global__ void __launch_bounds__(128, 1)
probe_gemm(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
           int rowsA, int N, int r, int bo_mode, float* D) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                 // rowsA * SWZ bytes (<= 256 rows)
    uint8_t* sB = smem + 256 * SWZ;     // N * SWZ bytes
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar_full, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<256>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar_full, (uint32_t)((rowsA + N) * SWZ));
        tma_load_2d(sA, &mapA, &bar_full, 0, 0);
        tma_load_2d(sB, &mapB, &bar_full, 0, 0);
        if (!mbar_wait_bounded(&bar_full, 0)) { D[0] = -12345.f; }
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t a_addr = smem_u32(sA) + r * SWZ;
        const uint32_t b_addr = smem_u32(sB);
        uint32_t bo = 0;
        if (bo_mode == 1) bo = (a_addr >> 7) & 7;
        constexpr int KSTEPS = SWZ / 32;  // 16 bf16 = 32 bytes per MMA K step
        for (int k = 0; k < KSTEPS; k++) {
            uint64_t da = make_smem_desc<SWZ>(a_addr + k * 32, bo);
            uint64_t db = make_smem_desc<SWZ>(b_addr + k * 32, 0);
            umma_bf16(tmem, da, db, idesc, k > 0);
        }
        umma_commit(&bar_mma);
    }
    if (!mbar_wait_bounded(&bar_mma, 0)) { if (threadIdx.x == 0) printf("TIMEOUT waiting for MMA commit\n"); }
    tc_fence_after();
    const int row = threadIdx.x;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[row * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------ TMA 4D probe kernel
__global__ void probe_tma4d(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3,
                            int bytes, uint8_t* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) smem[i] = 0xEE;
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        fence_proxy_async();
        mbar_expect_tx(&bar, bytes);
        tma_load_4d(smem, &map, &bar, c0, c1, c2, c3);
    }
    __syncthreads();
    if (!mbar_wait_bounded(&bar, 0)) { if (threadIdx.x == 0) printf("TIMEOUT waiting for TMA bytes\n"); }
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

// ------------------------------------------------------------------ MN-major probe (wgrad operands)
// dY: [R][CM] bf16, X: [R + 80][CN] bf16 (pixel rows, channels contiguous).  D[m][n] = sum_r dY[r][m] * X[r+shift][n]
// Both operands are MN-major: smem block j = channels [j*CB, (j+1)*CB) of all rows, CB = SWZ/2; blocks LBO apart.
template <int SWZ>
__global__ void __launch_bounds__(128, 1)
probe_mnmajor(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int R, int M, int N,
              int shift, float* D, int stack = 0) {
    // stack > 0 ("stacked taps"): X has only CB channels (ONE smem block); the N/CB column blocks of the B operand
    // are the same block viewed `stack` rows further down each: LBO = stack * SWZ bytes.
    //   D[m][g*CB + c] = sum_r dY[r][m] * X[r + shift + g*stack][c]
    constexpr int CB = SWZ / 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int RB = R + 80;
    const uint32_t blkA = (uint32_t)R * SWZ, blkB = (uint32_t)RB * SWZ;   // bytes per 64(32)-channel block
    uint8_t* sA = smem;
    uint8_t* sB = smem + (((M / CB) * blkA + 1023) & ~1023u);
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar_full, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<256>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 1) {
        if (elect_one()) {
            const int nbB = stack ? 1 : N / CB;
            mbar_expect_tx(&bar_full, (M / CB) * blkA + nbB * blkB);
            for (int j = 0; j < M / CB; j++) tma_load_2d(sA + j * blkA, &mapA, &bar_full, j * CB, 0);
            for (int j = 0; j < nbB; j++) tma_load_2d(sB + j * blkB, &mapB, &bar_full, j * CB, 0);
            if (!mbar_wait_bounded(&bar_full, 0)) { D[0] = -12345.f; }
            tc_fence_after();
            // instruction descriptor: bf16 x bf16 -> f32, A and B MN-major (bits 15, 16)
            const uint32_t idesc = make_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
            constexpr uint64_t layout = (SWZ == 128) ? 2ull : 4ull;
            constexpr uint64_t sbo = (8ull * SWZ) >> 4;     // next 8-row (K) group
            for (int k0 = 0; k0 < R; k0 += 16) {
                const uint32_t a_addr = smem_u32(sA) + k0 * SWZ;
                const uint32_t b_addr = smem_u32(sB) + (k0 + shift) * SWZ;
                const uint64_t da = uint64_t((a_addr & 0x3FFFFu) >> 4) | (uint64_t(blkA >> 4) << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
                const uint32_t lbo_b = stack ? (uint32_t)stack * SWZ : blkB;
                const uint64_t db = uint64_t((b_addr & 0x3FFFFu) >> 4) | (uint64_t(lbo_b >> 4) << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
                umma_bf16(tmem, da, db, idesc, k0 > 0);
            }
            umma_commit(&bar_mma);
        }
    }
    __syncwarp();
    if (!mbar_wait_bounded(&bar_mma, 0)) { if (threadIdx.x == 0) printf("TIMEOUT waiting for MMA commit\n"); }
    tc_fence_after();
    const int row = threadIdx.x;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[row * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

static float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

static std::vector<__nv_bfloat16> rand_bf16(size_t n, unsigned seed) {
    std::vector<__nv_bfloat16> v(n);
    srand(seed);
    for (size_t i = 0; i < n; i++) v[i] = __float2bfloat16((float)((rand() % 9) - 4));
    return v;
}

template <int SWZ>
static int run_gemm(int N, int r, int bo_mode, const char* tag, bool dump) {
    const int KB = SWZ / 2;
    const int rowsA = (r == 0) ? 128 : 144;
    auto hA = rand_bf16((size_t)rowsA * KB, 1 + r);
    auto hB = rand_bf16((size_t)N * KB, 77 + N);
    __nv_bfloat16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, 128 * N * 4));
    uint64_t dimsA[2] = {(uint64_t)KB, (uint64_t)rowsA}, strA[1] = {(uint64_t)KB * 2};
    uint32_t boxA[2] = {(uint32_t)KB, (uint32_t)rowsA};
    uint64_t dimsB[2] = {(uint64_t)KB, (uint64_t)N}, strB[1] = {(uint64_t)KB * 2};
    uint32_t boxB[2] = {(uint32_t)KB, (uint32_t)N};
    CUtensorMapSwizzle sw = SWZ == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : (SWZ == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUtensorMap mA = make_map(dA, 2, dimsA, strA, boxA, nullptr, sw);
    CUtensorMap mB = make_map(dB, 2, dimsB, strB, boxB, nullptr, sw);
    size_t smem = 512 * SWZ + 1024;
    CK(cudaFuncSetAttribute(probe_gemm<SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_gemm<SWZ><<<1, 128, smem>>>(mA, mB, rowsA, N, r, bo_mode, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%s SWZ=%d N=%d r=%d bo=%d : CUDA ERROR %s\n", tag, SWZ, N, r, bo_mode, cudaGetErrorString(e));
        exit(3);
    }
    std::vector<float> hD(128 * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    double maxerr = 0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < N; n++) {
            float ref = 0;
            for (int k = 0; k < KB; k++) ref += bf2f(hA[(size_t)(r + m) * KB + k]) * bf2f(hB[(size_t)n * KB + k]);
            float d = fabsf(ref - hD[m * N + n]);
            if (!(d <= 1e-3f)) bad++;
            if (d > maxerr) maxerr = d;
        }
    printf("%s SWZ=%d N=%d r=%d bo_mode=%d : %s (bad=%d/%d maxerr=%g)\n", tag, SWZ, N, r, bo_mode,
           bad ? "FAIL" : "PASS", bad, 128 * N, maxerr);
    if (bad && dump) {
        // identity-B diagnostic: which A element did the MMA see at (m, k)?
        std::vector<__nv_bfloat16> iA((size_t)rowsA * KB), iB((size_t)N * KB);
        for (int pass = 0; pass < 2; pass++) {
            for (int row = 0; row < rowsA; row++)
                for (int k = 0; k < KB; k++) iA[(size_t)row * KB + k] = __float2bfloat16(pass == 0 ? (float)row : (float)k);
            for (int n = 0; n < N; n++)
                for (int k = 0; k < KB; k++) iB[(size_t)n * KB + k] = __float2bfloat16(n == k ? 1.f : 0.f);
            CK(cudaMemcpy(dA, iA.data(), iA.size() * 2, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dB, iB.data(), iB.size() * 2, cudaMemcpyHostToDevice));
            probe_gemm<SWZ><<<1, 128, smem>>>(mA, mB, rowsA, N, r, bo_mode, dD);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
            char fn[256];
            snprintf(fn, sizeof fn, "gpurun_out/probe_%s_swz%d_r%d_bo%d_pass%d.txt", tag, SWZ, r, bo_mode, pass);
            FILE* f = fopen(fn, "w");
            if (f) {
                for (int m = 0; m < 128; m++) {
                    for (int n = 0; n < (N < KB ? N : KB); n++) fprintf(f, "%g ", hD[m * N + n]);
                    fprintf(f, "\n");
                }
                fclose(f);
            }
        }
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
    return bad;
}

template <int SWZ>
static void run_mnmajor(int R, int M, int N, int shift, int stack = 0) {
    const int CB = SWZ / 2;
    const int Nx = stack ? CB : N;      // channels X really has
    const int RB = R + 80;
    auto hA = rand_bf16((size_t)R * M, 5 + shift);
    auto hB = rand_bf16((size_t)RB * Nx, 9 + N);
    __nv_bfloat16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, 128 * N * 4));
    uint64_t dimsA[2] = {(uint64_t)M, (uint64_t)R}, strA[1] = {(uint64_t)M * 2};
    uint32_t boxA[2] = {(uint32_t)CB, (uint32_t)R};
    uint64_t dimsB[2] = {(uint64_t)Nx, (uint64_t)RB}, strB[1] = {(uint64_t)Nx * 2};
    uint32_t boxB[2] = {(uint32_t)CB, (uint32_t)RB};
    CUtensorMapSwizzle sw = SWZ == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap mA = make_map(dA, 2, dimsA, strA, boxA, nullptr, sw);
    CUtensorMap mB = make_map(dB, 2, dimsB, strB, boxB, nullptr, sw);
    size_t smem = (size_t)(M / CB) * R * SWZ + (size_t)(N / CB) * RB * SWZ + 4096;
    CK(cudaFuncSetAttribute(probe_mnmajor<SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_mnmajor<SWZ><<<1, 128, smem>>>(mA, mB, R, M, N, shift, dD, stack);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("mnmajor SWZ=%d M=%d N=%d shift=%d : CUDA ERROR %s\n", SWZ, M, N, shift, cudaGetErrorString(e));
        exit(3);
    }
    std::vector<float> hD(128 * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    // hypotheses for the TMEM lane of accumulator row m: 0: lane = m ; 1 (M=64): lane = (m/16)*32 + m%16 ; 2: lane = m*2
    for (int hyp = 0; hyp < (M == 64 ? 3 : 1); hyp++) {
        int bad = 0;
        double maxerr = 0;
        for (int m = 0; m < M; m++) {
            const int lane = hyp == 0 ? m : hyp == 1 ? (m / 16) * 32 + (m % 16) : m * 2;
            for (int n = 0; n < N; n++) {
                float ref = 0;
                if (stack) {
                    const int g = n / CB, c = n % CB;
                    for (int r = 0; r < R; r++) ref += bf2f(hA[(size_t)r * M + m]) * bf2f(hB[(size_t)(r + shift + g * stack) * Nx + c]);
                } else
                    for (int r = 0; r < R; r++) ref += bf2f(hA[(size_t)r * M + m]) * bf2f(hB[(size_t)(r + shift) * N + n]);
                float d = fabsf(ref - hD[lane * N + n]);
                if (!(d <= 1e-3f)) bad++;
                if (d > maxerr) maxerr = d;
            }
        }
        printf("mnmajor SWZ=%d R=%d M=%d N=%d shift=%d stack=%d lane-hyp=%d : %s (bad=%d/%d maxerr=%g)\n", SWZ, R, M, N, shift, stack, hyp,
               bad ? "FAIL" : "PASS", bad, M * N, maxerr);
        if (!bad) break;
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
}

static void run_tma4d(bool strided) {
    // NHWC tensor [N=2][H=12][W=20][C=64] bf16; value encodes (n,h,w) in channel 0..2, c index in rest
    const int N = 2, H = 12, W = 20, C = 64;
    std::vector<__nv_bfloat16> h((size_t)N * H * W * C);
    for (int n = 0; n < N; n++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                for (int c = 0; c < C; c++) {
                    float v = (c == 0) ? n + 1 : (c == 1) ? y + 1 : (c == 2) ? x + 1 : c;
                    h[(((size_t)n * H + y) * W + x) * C + c] = __float2bfloat16(v);
                }
    __nv_bfloat16* d;
    CK(cudaMalloc(&d, h.size() * 2));
    CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    for (int variant = 0; variant < (strided ? 2 : 1); variant++) {
        // box: 8 wide x 4 high output pixels
        uint32_t box[4] = {64, 8, 4, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        if (strided) {
            es[1] = 2;
            es[2] = 2;
            if (variant == 0) {  // hypothesis A: box counts traversed elements (span), smem gets ceil(box/stride)
                box[1] = 16;
                box[2] = 8;
            }  // variant 1, hypothesis B: box counts loaded elements
        }
        CUtensorMap m = make_map(d, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
        const int bytes = 8 * 4 * 128;
        uint8_t* dout;
        CK(cudaMalloc(&dout, bytes));
        CK(cudaMemset(dout, 0xDD, bytes));
        const int c1 = -1, c2 = -1, c3 = 1;  // start at w=-1, h=-1 of image 1
        probe_tma4d<<<1, 128, bytes + 1024>>>(m, 0, c1, c2, c3, bytes, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("tma4d strided=%d variant=%d: CUDA ERROR %s\n", strided, variant, cudaGetErrorString(e));
            exit(3);
        }
        std::vector<uint8_t> ho(bytes);
        CK(cudaMemcpy(ho.data(), dout, bytes, cudaMemcpyDeviceToHost));
        // decode: smem row i (128B), un-swizzle chunk j -> logical chunk j ^ (i & 7)
        int bad = 0;
        printf("tma4d strided=%d variant=%d rows (n,h,w) per smem row:\n", strided, variant);
        for (int i = 0; i < 32; i++) {
            const __nv_bfloat16* rowp = (const __nv_bfloat16*)(ho.data() + i * 128);
            int phys_chunk0 = (0 ^ (i & 7));
            float vn = bf2f(rowp[phys_chunk0 * 8 + 0]), vh = bf2f(rowp[phys_chunk0 * 8 + 1]), vw = bf2f(rowp[phys_chunk0 * 8 + 2]);
            int phys_chunk5 = (5 ^ (i & 7));
            float vc = bf2f(rowp[phys_chunk5 * 8 + 3]);  // logical channel 43
            int st = strided ? 2 : 1;
            int eh = c2 + (i / 8) * st, ew = c1 + (i % 8) * st;
            bool inb = eh >= 0 && eh < H && ew >= 0 && ew < W;
            float en = inb ? 2 : 0, ehv = inb ? eh + 1 : 0, ewv = inb ? ew + 1 : 0, ec = inb ? 43 : 0;
            bool ok = (vn == en && vh == ehv && vw == ewv && vc == ec);
            if (!ok) bad++;
            printf("  row %2d: n=%g h=%g w=%g c43=%g  expect n=%g h=%g w=%g %s\n", i, vn, vh - 1, vw - 1, vc, en, ehv - 1, ewv - 1, ok ? "" : "<-- MISMATCH");
        }
        printf("tma4d strided=%d variant=%d : %s\n", strided, variant, bad ? "FAIL" : "PASS");
        cudaFree(dout);
    }
    cudaFree(d);
}

int main(int argc, char** argv) {
    int test = argc > 1 ? atoi(argv[1]) : 1;
    init_driver();
    if (test == 1) {
        int Ns[4] = {64, 16, 128, 256};
        for (int i = 0; i < 4; i++) run_gemm<128>(Ns[i], 0, 0, "basic", i == 0);
    } else if (test == 2) {
        for (int bo = 0; bo < 2; bo++)
            for (int r = 1; r <= 9; r++) run_gemm<128>(64, r, bo, "rowshift", r == 1 || r == 3);
    } else if (test == 3) {
        run_tma4d(false);
    } else if (test == 4) {
        run_tma4d(true);
    } else if (test == 5) {
        run_gemm<64>(64, 0, 0, "sw64", true);
        run_gemm<64>(32, 0, 0, "sw64", false);
        for (int bo = 0; bo < 2; bo++)
            for (int r = 1; r <= 5; r++) run_gemm<64>(64, r, bo, "sw64shift", r == 1);
        run_gemm<32>(64, 0, 0, "sw32", true);
        for (int bo = 0; bo < 2; bo++)
            for (int r = 1; r <= 3; r++) run_gemm<32>(64, r, bo, "sw32shift", false);
    }
    else if (test == 6) {
        for (int shift : {0, 1, 3, 8, 67}) run_mnmajor<128>(64, 128, 128, shift);
        run_mnmajor<128>(128, 128, 64, 5);
        run_mnmajor<128>(64, 64, 64, 0);
        run_mnmajor<128>(64, 64, 128, 2);
        run_mnmajor<128>(32, 128, 256, 66);
        for (int shift : {0, 1, 5, 66}) run_mnmajor<64>(64, 128, 32, shift);
        run_mnmajor<64>(64, 32, 32, 3);
        run_mnmajor<64>(64, 64, 128, 1);
    } else if (test == 7) {
        // stacked taps: the N blocks of B are ONE X block at row offsets g*stack (LBO = stack rows)
        for (int shift : {0, 1, 66}) run_mnmajor<64>(64, 128, 96, shift, 1);      // 3 horizontal taps x 32 ch
        run_mnmajor<64>(64, 64, 96, 2, 1);
        run_mnmajor<64>(32, 64, 160, 0, 10);                                      // 5 vertical taps, Wp = 10
        run_mnmajor<64>(32, 64, 128, 3, 13);
        for (int shift : {0, 1, 67}) run_mnmajor<128>(64, 64, 192, shift, 1);     // 3 horizontal taps x 64 ch
        run_mnmajor<128>(64, 128, 192, 5, 1);
        run_mnmajor<128>(32, 128, 128, 5, 7);
    }
    return 0;
}

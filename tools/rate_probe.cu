// Throughput probe (B200, run via gpurun): what bounds the halo-reuse conv kernels?
//   mma   : cycles per tcgen05.mma (M=128, K=16, SS) for N in {16,32,64,128,256}, aligned and row-shifted A
//   tma   : cycles per 4-D TMA box for 64 B / 128 B inner rows, 2 stages in flight, all SMs streaming
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../depth_aware_endoscopy_sr_b200/csrc/sm100_ptx.cuh"
using namespace dasr;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static CUtensorMap make_map4(void* base, int C, int W, int H, int B, int bw, int bh, CUtensorMapSwizzle swz) {
    CUtensorMap m;
    cuuint64_t d[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t s[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t b[4] = {(cuuint32_t)C, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t e[4] = {1, 1, 1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(2); }
    return m;
}

// ---------------------------------------------------------------- MMA issue rate
template <int SWZ>
__global__ void __launch_bounds__(128, 1) mma_rate(int N, int shift_rows, int n_mma, int mode, long long* out) {
    const int distinct = 9;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t hi = make_smem_desc<SWZ>(0, 0) & ~0x3FFFull;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 40 * 1024;
    if (mode == 0) {            // divergent single thread (what conv_igemm.cu v1 does)
        if (threadIdx.x == 0) {
            long long t0 = clock64();
            for (int i = 0; i < n_mma; i++) {
                const uint32_t aa = a0 + (uint32_t)((i % distinct) * shift_rows) * SWZ + (i & 1) * 32;
                const uint32_t bb = b0 + (i & 1) * 32;
                umma_bf16(tmem + (i & 1) * 256, hi | ((aa & 0x3FFFFu) >> 4), hi | ((bb & 0x3FFFFu) >> 4), idesc, 1);
            }
            umma_commit(&bar);
            long long t1 = clock64();
            while (!mbar_try_wait(&bar, 0)) {}
            long long t2 = clock64();
            out[blockIdx.x * 2] = t1 - t0;
            out[blockIdx.x * 2 + 1] = t2 - t0;
        }
    } else if (mode == 1) {     // warp-uniform branch + elect.sync around the whole loop
        if (warp == 0) {
            if (elect_one()) {
                long long t0 = clock64();
                for (int i = 0; i < n_mma; i++) {
                    const uint32_t aa = a0 + (uint32_t)((i % distinct) * shift_rows) * SWZ + (i & 1) * 32;
                    const uint32_t bb = b0 + (i & 1) * 32;
                    umma_bf16(tmem + (i & 1) * 256, hi | ((aa & 0x3FFFFu) >> 4), hi | ((bb & 0x3FFFFu) >> 4), idesc, 1);
                }
                umma_commit(&bar);
                long long t1 = clock64();
                while (!mbar_try_wait(&bar, 0)) {}
                long long t2 = clock64();
                out[blockIdx.x * 2] = t1 - t0;
                out[blockIdx.x * 2 + 1] = t2 - t0;
            }
        }
    } else {                    // converged warp runs the loop, elect.sync around each group of 4 MMAs
        if (warp == 0) {
            long long t0 = clock64();
            for (int i = 0; i < n_mma; i += 4) {
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t aa = a0 + (uint32_t)(((i + j) % distinct) * shift_rows) * SWZ + (j & 1) * 32;
                        const uint32_t bb = b0 + (j & 1) * 32;
                        umma_bf16(tmem + (j & 1) * 256, hi | ((aa & 0x3FFFFu) >> 4), hi | ((bb & 0x3FFFFu) >> 4), idesc, 1);
                    }
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&bar);
            __syncwarp();
            long long t1 = clock64();
            while (!mbar_try_wait(&bar, 0)) {}
            long long t2 = clock64();
            if (threadIdx.x == 0) {
                out[blockIdx.x * 2] = t1 - t0;
                out[blockIdx.x * 2 + 1] = t2 - t0;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------- TMA box rate
__global__ void __launch_bounds__(128, 1) tma_rate(const __grid_constant__ CUtensorMap map, int box_bytes, int stages, int n_boxes, int tiles_w, int tiles_h, int bw_step, int bh_step, int nimg, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[4];
    const int stage_bytes = (box_bytes + 1023) & ~1023;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) mbar_init(&full[i], 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        int issued = 0, done = 0;
        auto issue = [&](int idx) {
            const int tile = blockIdx.x + idx * gridDim.x;
            const int tpi = tiles_w * tiles_h;
            const int img = (tile / tpi) % nimg;
            const int r = tile % tpi;
            const int s = idx % stages;
            mbar_expect_tx(&full[s], box_bytes);
            tma_load_4d(smem + (size_t)s * stage_bytes, &map, &full[s], 0, (r % tiles_w) * bw_step - 1, (r / tiles_w) * bh_step - 1, img);
        };
        for (; issued < stages && issued < n_boxes; issued++) issue(issued);
        while (done < n_boxes) {
            const int s = done % stages;
            while (!mbar_try_wait(&full[s], (done / stages) & 1)) {}
            done++;
            if (issued < n_boxes) { issue(issued); issued++; }
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

int main(int argc, char** argv) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    g_encode = (EncodeTiledFn)fn;
    long long* d_out; CK(cudaMalloc(&d_out, 4096 * sizeof(long long)));
    std::vector<long long> h(4096);
    CK(cudaFuncSetAttribute(mma_rate<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(mma_rate<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(tma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    printf("== MMA issue / completion rate (cycles per tcgen05.mma, M=128 K=16 SS), 1 CTA and 148 CTAs\n");
    for (int mode : {0, 1, 2}) for (int swz : {128, 64}) for (int N : {16, 32, 64, 128, 256}) for (int shift : {0, 67}) for (int grid : {148}) {
        const int n = 2048;
        if (swz == 128) mma_rate<128><<<grid, 128, 100 * 1024>>>(N, shift, n, mode, d_out); else mma_rate<64><<<grid, 128, 100 * 1024>>>(N, shift, n, mode, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), d_out, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mi = 0, mt = 0; for (int i = 0; i < grid; i++) { mi = std::max(mi, h[2 * i]); mt = std::max(mt, h[2 * i + 1]); }
        printf("mode=%d swz=%3d N=%3d shift_rows=%2d grid=%3d : issue %.1f cyc/mma, complete %.1f cyc/mma (floor %d)\n", mode, swz, N, shift, grid, (double)mi / n, (double)mt / n, 128 * N / 256);
    }
    printf("== TMA box rate (2-4 stages in flight per SM, 148 SMs)\n");
    // tensors: C=32 (64 B rows) 64 x 256 x 256 ; C=64 (128 B rows) 64 x 128 x 128... sized 268 MB each (> L2)
    struct Case { int C, W, H, B, bw, bh, stepw, steph; const char* name; };
    Case cases[] = {
#if 0
        {32, 256, 256, 64, 130, 5, 128, 2, "C32 box 130x5 (3x3 32ch strip)"},
        {32, 512, 512, 16, 64, 22, 56, 14, "C32 box 64x22 (out9 patch)"},
        {64, 128, 256, 64, 66, 7, 64, 4, "C64 box 66x7 (trunk chunk)"},
        {64, 128, 256, 64, 128, 5, 128, 3, "C64 box 128x5"},
        {64, 128, 256, 64, 64, 16, 64, 14, "C64 box 64x16"},
#endif
        {64, 128, 256, 64, 66, 7, 64, 4, "C64 box 66x7 (trunk chunk)"},
    };
    for (auto& c : cases) {
        size_t bytes = (size_t)c.C * c.W * c.H * c.B * 2;
        void* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 0, bytes));
        CUtensorMap m = make_map4(d, c.C, c.W, c.H, c.B, c.bw, c.bh, c.C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
        const int box_bytes = c.C * 2 * c.bw * c.bh;
        const int tiles_w = (c.W + c.stepw - 1) / c.stepw, tiles_h = (c.H + c.steph - 1) / c.steph;
        for (int nimg : {c.B, 1}) for (int stages : {1, 2, 4}) {
            if ((size_t)stages * ((box_bytes + 1023) & ~1023) > 215 * 1024) continue;
            const int n_boxes = 200;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            tma_rate<<<148, 128, 216 * 1024>>>(m, box_bytes, stages, n_boxes, tiles_w, tiles_h, c.stepw, c.steph, nimg, d_out);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            tma_rate<<<148, 128, 216 * 1024>>>(m, box_bytes, stages, n_boxes, tiles_w, tiles_h, c.stepw, c.steph, nimg, d_out);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            CK(cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < 148; i++) mx = std::max(mx, h[i]);
            printf("%-32s %s stages=%d box=%6d B rows=%4d : %.0f cyc/box  %.2f cyc/row  %.1f B/cyc/SM  aggregate %.2f TB/s\n", c.name, nimg == 1 ? "L2-resident" : "HBM-stream ", stages, box_bytes, c.bw * c.bh, (double)mx / n_boxes, (double)mx / n_boxes / (c.bw * c.bh), (double)box_bytes * n_boxes / mx, 148.0 * n_boxes * box_bytes / (ms * 1e-3) / 1e12);
        }
        CK(cudaFree(d));
    }
    return 0;
}

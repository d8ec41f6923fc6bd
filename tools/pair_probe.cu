// CTA-pair (tcgen05.mma.cta_group::2) dispatch-rate probe (B200, run via gpurun):
// cycles per M=256 MMA issued by the leader of a 2-CTA cluster, operands in shared memory of both CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/pair_probe tools/pair_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../depth_aware_endoscopy_sr_b200/csrc/sm100_ptx.cuh"
using namespace dasr;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

template <int SWZ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_rate(int N, int shift_rows, int n_mma, int b_distinct, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 150 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc_pair<512>(&tmem_base_s);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_bf16(256, N);
    const uint64_t hi = make_smem_desc<SWZ>(0, 0) & ~0x3FFFull;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 72 * 1024;
    if (warp == 0 && rank == 0) {
        if (elect_one()) {
            long long t0 = clock64();
            for (int i = 0; i < n_mma; i++) {
                const uint32_t aa = a0 + (uint32_t)((i % 9) * shift_rows) * SWZ + (i & 3) * 32;
                const uint32_t bb = b0 + (uint32_t)((i >> 2) % b_distinct) * 8192 + (i & 3) * 32;
                const uint32_t alo = (uint32_t)(hi) | ((aa & 0x3FFFFu) >> 4), blo = (uint32_t)(hi) | ((bb & 0x3FFFFu) >> 4);
                umma_bf16_lohi_pair(tmem + (i & 1) * 256, alo, blo, (uint32_t)(hi >> 32), idesc, 1);
            }
            umma_commit_pair(&bar);
            long long t1 = clock64();
            while (!mbar_try_wait(&bar, 0)) {}
            long long t2 = clock64();
            out[(blockIdx.x >> 1) * 2] = t1 - t0;
            out[(blockIdx.x >> 1) * 2 + 1] = t2 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair<512>(tmem);
}

int main() {
    long long* d_out; CK(cudaMalloc(&d_out, 4096 * sizeof(long long)));
    std::vector<long long> h(4096);
    CK(cudaFuncSetAttribute(pair_rate<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(pair_rate<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    printf("== cta_group::2 MMA rate (cycles per tcgen05.mma, M=256 K=16 SS), 74 pairs\n");
    for (int swz : {128, 32}) for (int N : {64, 128, 256}) for (int shift : {0, 67}) for (int bd : {1, 8}) {
        const int n = 2048, grid = 148;
        if (swz == 128) pair_rate<128><<<grid, 128, 200 * 1024>>>(N, shift, n, bd, d_out);
        else pair_rate<32><<<grid, 128, 200 * 1024>>>(N, shift, n, bd, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mi = 0, mt = 0;
        for (int i = 0; i < grid / 2; i++) { mi = std::max(mi, h[2 * i]); mt = std::max(mt, h[2 * i + 1]); }
        printf("swz=%3d N=%3d shift_rows=%2d b_tiles=%d : issue %.1f cyc/mma, complete %.1f cyc/mma (floor %d)\n", swz, N, shift, bd, (double)mi / n, (double)mt / n, 256 * N / 512);
    }
    return 0;
}

// K-WN / weight preparation: weight-norm (w = g*v/||v||), SEAN alpha folding and repacking of fp32
// conv weights into the bf16 K-major GEMM-B layout [row][tap*I + c] consumed by conv_igemm.cu.
// Follows torch.nn.utils.weight_norm(dim=0) as used at codes/models/modules/sftmd_arch.py:740,851
// (for ConvTranspose2d dim 0 is Cin) and the blend of codes/models/modules/normalization.py:87-88.
#include "dasr_internal.h"

namespace dasr {

constexpr int kPackBatch = 28;

struct PackBatch {
    dasr_pack_desc d[kPackBatch];
    int scale_off[kPackBatch];
    int n;
};

// one block per (descriptor, dim0 index): scale[o] = (g ? g[o]/||v[o]|| : 1) * alpha_factor
__global__ void pack_scale_kernel(const PackBatch pb, float* __restrict__ scratch) {
    int o = blockIdx.x, di = 0;
    while (di < pb.n && o >= pb.d[di].dim0) {
        o -= pb.d[di].dim0;
        di++;
    }
    if (di >= pb.n) return;
    const dasr_pack_desc& d = pb.d[di];
    const int inner = d.dim1 * d.ks * d.ks;
    float scale = 1.f;
    if (d.g) {
        float ss = 0.f;
        const float* vp = d.v + (size_t)o * inner;
        for (int i = threadIdx.x; i < inner; i += blockDim.x) {
            float x = vp[i];
            ss += x * x;
        }
        __shared__ float red[32];
        for (int off = 16; off; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
        __syncthreads();
        if (threadIdx.x < 32) {
            float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
            for (int off = 16; off; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (threadIdx.x == 0) red[0] = t;
        }
        __syncthreads();
        scale = d.g[o] / sqrtf(red[0]);
    }
    if (threadIdx.x == 0) {
        if (d.alpha_mode == 1) scale *= *d.alpha;
        if (d.alpha_mode == 2) scale *= (1.f - *d.alpha);
        scratch[pb.scale_off[di] + o] = scale;
    }
}

__global__ void pack_write_kernel(const PackBatch pb, const float* __restrict__ scratch) {
    const dasr_pack_desc& d = pb.d[blockIdx.y];
    const int ks = d.ks, taps = ks * ks;
    const int O = d.mode != DASR_PACK_CONVT ? d.dim0 : d.dim1;
    const int I = d.mode != DASR_PACK_CONVT ? d.dim1 : d.dim0;
    const int K = taps * I;
    const float* sc = scratch + pb.scale_off[blockIdx.y];
    const int total = O * K;
    __nv_bfloat16* dst = (__nv_bfloat16*)d.dst;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int o = idx / K;
        const int kk = idx - o * K;
        const int tap = kk / I, c = kk - tap * I;
        const int t = tap / ks, u = tap - t * ks;
        float w;
        if (d.mode == DASR_PACK_STYLE) {
            w = d.v[(((size_t)o * I + c) * ks + t) * ks + u] * sc[o];
            const int row = tap * d.rows_per_tap + d.row_offset + o;
            dst[(size_t)row * I + c] = __float2bfloat16(w);
            continue;
        }
        if (d.mode == DASR_PACK_ROWTAPS) {
            // K-OUT9 operand: one [32][I] matrix per vertical tap t, row = u*O + o (horizontal taps folded into N)
            w = d.v[(((size_t)o * I + c) * ks + t) * ks + u] * sc[o];
            dst[((size_t)t * 32 + u * O + o) * I + c] = __float2bfloat16(w);
            if (kk == 0 && d.dst_bias) d.dst_bias[o] = d.bias ? d.bias[o] : 0.f;
            continue;
        }
        if (d.mode == DASR_PACK_CONV) {
            w = d.v[(((size_t)o * I + c) * ks + t) * ks + u] * sc[o];
        } else {  // ConvTranspose2d weight [I][O][ks][ks] as an ordinary conv: flip taps, swap in/out
            w = d.v[(((size_t)c * O + o) * ks + (ks - 1 - t)) * ks + (ks - 1 - u)] * sc[c];
        }
        int row = o;
        if (d.shuffle_r > 1) {
            const int r2 = d.shuffle_r * d.shuffle_r;
            row = (o % r2) * (O / r2) + o / r2;
        }
        row += d.row_offset;
        dst[(size_t)row * K + kk] = __float2bfloat16(w);
        if (kk == 0 && d.dst_bias) {
            float f = 1.f;
            if (d.alpha_mode == 1) f = *d.alpha;
            if (d.alpha_mode == 2) f = 1.f - *d.alpha;
            float b = d.bias ? f * d.bias[o] : 0.f;
            if (d.bias2) b += (1.f - f) * d.bias2[o];
            d.dst_bias[row] = b;
        }
    }
}

}  // namespace dasr

using namespace dasr;

extern "C" int dasr_pack_weights(const dasr_pack_desc* descs, int n, float* scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DASR_REQUIRE(descs && n > 0 && scratch, "bad arguments");
    int scale_base = 0;
    for (int start = 0; start < n; start += kPackBatch) {
        PackBatch pb;
        pb.n = (n - start) < kPackBatch ? (n - start) : kPackBatch;
        int rows = 0, max_elems = 0;
        for (int i = 0; i < pb.n; i++) {
            pb.d[i] = descs[start + i];
            const dasr_pack_desc& d = pb.d[i];
            DASR_REQUIRE(d.v && d.dst && d.dim0 > 0 && d.dim1 > 0 && d.ks > 0, "bad pack descriptor %d", start + i);
            DASR_REQUIRE(d.alpha_mode == 0 || d.alpha, "descriptor %d: alpha_mode without alpha", start + i);
            pb.scale_off[i] = scale_base + rows;
            rows += d.dim0;
            int e = d.dim0 * d.dim1 * d.ks * d.ks;
            if (e > max_elems) max_elems = e;
        }
        pack_scale_kernel<<<rows, 128, 0, stream>>>(pb, scratch);
        DASR_LAUNCH_OK();
        int bx = (max_elems + 255) / 256;
        if (bx > 256) bx = 256;
        pack_write_kernel<<<dim3(bx, pb.n), 256, 0, stream>>>(pb, scratch);
        DASR_LAUNCH_OK();
        scale_base += rows;
    }
    return DASR_OK;
}

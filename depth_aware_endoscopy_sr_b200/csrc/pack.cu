// K-WN / weight preparation: weight-norm (w = g*v/||v||), SEAN alpha folding and repacking of fp32
// conv weights into the bf16 K-major GEMM-B layout [row][tap*I + c] consumed by conv_igemm.cu.
// Follows torch.nn.utils.weight_norm(dim=0) as used at codes/models/modules/sftmd_arch.py:740,851
// (for ConvTranspose2d dim 0 is Cin) and the blend of codes/models/modules/normalization.py:87-88.
#include "dasr_internal.h"

namespace dasr {

constexpr int kPackBatch = 320;      // descriptors per launch: the 32 KB kernel-parameter space of CUDA 12.1+ (sm_70+)

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

// (npl > 1: every packed weight is split into npl bf16 planes, d.dst_plane_stride elements apart -- dasr_internal.h)
__global__ void pack_write_kernel(const PackBatch pb, const float* __restrict__ scratch, int npl) {
    const dasr_pack_desc& d = pb.d[blockIdx.y];
    const size_t ps = (size_t)d.dst_plane_stride;
    const int ks = d.ks, taps = ks * ks;
    if (d.mode >= DASR_PACK_DGRAD) {
        // data-gradient layouts: walk the SOURCE linearly (coalesced reads), scatter to dst
        const float* scd = scratch + pb.scale_off[blockIdx.y];
        __nv_bfloat16* dstd = (__nv_bfloat16*)d.dst;
        const int n = d.dim0 * d.dim1 * taps;
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
            const int a = idx / (d.dim1 * taps);
            const int rem = idx - a * (d.dim1 * taps);
            const int b = rem / taps, tap = rem - b * taps;
            const float w = d.v[idx] * scd[a];
            if (d.mode == DASR_PACK_DGRAD) {            // a = o, b = i
                const int Ot = d.rows_per_tap > 0 ? d.rows_per_tap : d.dim0;
                int no = a;
                if (d.shuffle_r > 1) {
                    const int r2 = d.shuffle_r * d.shuffle_r;
                    no = (a % r2) * (d.dim0 / r2) + a / r2;
                }
                pl_store1(dstd, (size_t)b * (taps * Ot) + (size_t)(taps - 1 - tap) * Ot + d.row_offset + no, ps, npl, w);
            } else if (d.mode == DASR_PACK_DGRAD_CONVT) {   // a = i, b = o
                pl_store1(dstd, (size_t)a * (taps * d.dim1) + (size_t)tap * d.dim1 + b, ps, npl, w);
            } else {                                     // OUT9_DGRAD: a = o (3), b = i (32), tap = t*9 + u
                const int t = tap / ks, u = tap - t * ks;
                pl_store1(dstd, (size_t)b * (ks * 32) + (size_t)(ks - 1 - t) * 32 + u * d.dim0 + a, ps, npl, w);
            }
        }
        return;
    }
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
            pl_store1(dst, (size_t)row * I + c, ps, npl, w);
            continue;
        }
        if (d.mode == DASR_PACK_ROWTAPS) {
            // K-OUT9 operand: one [32][I] matrix per vertical tap t, row = u*O + o (horizontal taps folded into N)
            w = d.v[(((size_t)o * I + c) * ks + t) * ks + u] * sc[o];
            pl_store1(dst, ((size_t)t * 32 + u * O + o) * I + c, ps, npl, w);
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
        pl_store1(dst, (size_t)row * K + kk, ps, npl, w);
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

// ------------------------------------------------------------------------------------ gradient unpacking
// Inverse of the packing above for GRADIENTS: the weight-gradient kernels produce d(packed weight) in fp32; this
// kernel gathers it back into the parameter's own layout and applies the chain rule of
//   w_eff[a] = f * (g[a] / ||v[a]||) * v[a]      (f = alpha | 1 - alpha | 1; g absent for plain convs)
// i.e. weight-norm backward (sftmd_arch.py:740,851 -> torch._weight_norm backward) and the SEAN blend
// (normalization.py:87-88).  One block per (descriptor, dim0 index a).
constexpr int kUnpackBatch = 224;
struct UnpackBatch {
    dasr_unpack_desc d[kUnpackBatch];
    int n;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[i];
    return t;
}

__device__ __forceinline__ size_t packed_index(const dasr_unpack_desc& d, int a, int b, int tap) {
    const int ks = d.ks, taps = ks * ks;
    const int ipk = d.ipack > 0 ? d.ipack : d.dim1;
    switch (d.mode) {
        case DASR_PACK_CONV: {
            int row = a;
            if (d.shuffle_r > 1) {
                const int r2 = d.shuffle_r * d.shuffle_r;
                row = (a % r2) * (d.dim0 / r2) + a / r2;
            }
            return (size_t)(row + d.row_offset) * (taps * ipk) + (size_t)tap * ipk + b;
        }
        case DASR_PACK_CONVT:   // a = i, b = o ; forward packed w'[o][(flip tap, i)]
            return (size_t)b * (taps * d.dim0) + (size_t)(taps - 1 - tap) * d.dim0 + a;
        case DASR_PACK_STYLE:
            return (size_t)(tap * d.rows_per_tap + d.row_offset + a) * d.dim1 + b;
        default: {              // DASR_PACK_ROWTAPS: gradient laid out [u*O + o][t*32 + c]
            const int t = tap / ks, u = tap - t * ks;
            return (size_t)(u * d.dim0 + a) * (ks * 32) + (size_t)t * 32 + b;
        }
    }
}

__global__ void __launch_bounds__(128) unpack_grads_kernel(const UnpackBatch ub) {
    int a = blockIdx.x, di = 0;
    while (di < ub.n && a >= ub.d[di].dim0) {
        a -= ub.d[di].dim0;
        di++;
    }
    if (di >= ub.n) return;
    const dasr_unpack_desc& d = ub.d[di];
    __shared__ float red[8];
    const int taps = d.ks * d.ks;
    const int inner = d.dim1 * taps;
    const float* vp = d.v + (size_t)a * inner;
    // pass 1: <DW, v> and ||v||^2
    float dot = 0.f, vv = 0.f;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) {
        const int b = i / taps, tap = i - b * taps;
        const float w = vp[i];
        dot = fmaf(d.dwp[packed_index(d, a, b, tap)], w, dot);
        vv = fmaf(w, w, vv);
    }
    dot = block_sum(dot, red);
    vv = block_sum(vv, red);
    float f = 1.f;
    if (d.alpha_mode == 1) f = *d.alpha;
    if (d.alpha_mode == 2) f = 1.f - *d.alpha;
    const float nrm = sqrtf(vv);
    const float s = d.g ? d.g[a] / nrm : 1.f;           // w_raw = s * v
    // pass 2: dv
    const float k = d.g ? dot / vv : 0.f;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) {
        const int b = i / taps, tap = i - b * taps;
        const float dw = d.dwp[packed_index(d, a, b, tap)];
        d.dv[(size_t)a * inner + i] = f * s * (dw - k * vp[i]);
    }
    if (threadIdx.x == 0) {
        if (d.g) d.dg[a] = f * dot / nrm;
        float dal = 0.f;
        if (d.alpha_mode) dal = (d.alpha_mode == 1 ? 1.f : -1.f) * s * dot;   // <DW, w_raw>
        if (d.dbias_p && d.mode == DASR_PACK_CONVT) {
            // ConvTranspose2d: dim0 is Cin, the bias has dim1 = Cout entries and no weight-norm / alpha factor
            if (a == 0 && d.dbias)
                for (int o = 0; o < d.dim1; o++) d.dbias[o] = d.dbias_p[o];
        } else if (d.dbias_p) {
            int row = a;
            if (d.mode == DASR_PACK_CONV && d.shuffle_r > 1) {
                const int r2 = d.shuffle_r * d.shuffle_r;
                row = (a % r2) * (d.dim0 / r2) + a / r2;
            }
            const float db = d.dbias_p[row + d.row_offset];
            if (d.dbias) d.dbias[a] = (d.alpha_mode ? f : 1.f) * db;
            if (d.dbias2) d.dbias2[a] = (1.f - f) * db;
            if (d.alpha_mode == 2 && d.bias && d.bias2) dal += db * (d.bias2[a] - d.bias[a]);
        }
        if (d.alpha_mode && d.dalpha) atomicAdd(d.dalpha, dal);
    }
}

}  // namespace dasr

using namespace dasr;

extern "C" int dasr_unpack_grads(const dasr_unpack_desc* descs, int n, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DASR_REQUIRE(descs && n > 0, "bad arguments");
    for (int start = 0; start < n; start += kUnpackBatch) {
        UnpackBatch ub;
        ub.n = (n - start) < kUnpackBatch ? (n - start) : kUnpackBatch;
        int rows = 0;
        for (int i = 0; i < ub.n; i++) {
            ub.d[i] = descs[start + i];
            const dasr_unpack_desc& d = ub.d[i];
            DASR_REQUIRE(d.dwp && d.v && d.dv && d.dim0 > 0 && d.dim1 > 0 && d.ks > 0, "bad unpack descriptor %d", start + i);
            DASR_REQUIRE(!d.g || d.dg, "descriptor %d: weight_g without a gradient buffer", start + i);
            DASR_REQUIRE(d.alpha_mode == 0 || d.alpha, "descriptor %d: alpha_mode without alpha", start + i);
            rows += d.dim0;
        }
        unpack_grads_kernel<<<rows, 128, 0, stream>>>(ub);
        DASR_LAUNCH_OK();
    }
    return DASR_OK;
}

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
            DASR_REQUIRE(planes() == 1 || d.dst_plane_stride > 0, "descriptor %d: dst_plane_stride missing (fp32-split planes)", start + i);
            pb.scale_off[i] = scale_base + rows;
            rows += d.dim0;
            int e = d.dim0 * d.dim1 * d.ks * d.ks;
            if (e > max_elems) max_elems = e;
        }
        pack_scale_kernel<<<rows, 128, 0, stream>>>(pb, scratch);
        DASR_LAUNCH_OK();
        int bx = (max_elems + 255) / 256;
        if (bx > 256) bx = 256;
        pack_write_kernel<<<dim3(bx, pb.n), 256, 0, stream>>>(pb, scratch, planes());
        DASR_LAUNCH_OK();
        scale_base += rows;
    }
    return DASR_OK;
}

/*
 * dasr.h -- C ABI of libdasr_b200.so: the B200 (sm_100a) kernels behind the DepthNet hot path of
 * CUHK-AIM-Group/Depth-Aware-Endoscopy-SR.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; the boundary it offers
 * is `models.networks.define_G(opt)` (codes/models/networks.py:15,41-49) returning an nn.Module whose
 * arithmetic lives in torch ops.  Each entry point below replaces the torch op(s) cited next to it; the
 * Python host (depth_aware_endoscopy_sr_b200/arch.py) binds them with ctypes exactly as INTEGRATION.md
 * shows.
 *
 * Conventions
 *  - every function returns 0 on success, a negative dasr_status otherwise; dasr_last_error() returns
 *    a thread-local message.  No exceptions cross the ABI, nothing falls back to the CPU.
 *  - all pointers are DEVICE pointers borrowed for the call (owner: the caller / torch caching
 *    allocator); the library allocates nothing persistent on the device.
 *  - `stream` is a cudaStream_t passed as void*; launches are asynchronous, no implicit sync.
 *  - activations are NHWC bf16 ("act" tensors) unless stated; network inputs/outputs are NCHW fp32 like
 *    the reference.
 */
#ifndef DASR_H_
#define DASR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DASR_OK = 0,
    DASR_ERR_BAD_ARG = -1,      /* shape / alignment / unsupported configuration */
    DASR_ERR_ARCH = -2,         /* device is not sm_100 */
    DASR_ERR_CUDA = -3,         /* launch or driver failure */
    DASR_ERR_WORKSPACE = -4     /* workspace too small */
} dasr_status;

const char* dasr_last_error(void);
int dasr_version(void);
/* 0 if the current device is a B200-class (sm_100) part, DASR_ERR_ARCH otherwise */
int dasr_check_device(void);
/* number of kernels this library has launched in the calling process so far (bench.py's gpu_launches) */
int64_t dasr_launch_count(void);

/* Storage of "act" tensors and packed weights (process-wide switch; TEST INFRASTRUCTURE, default 1):
 *   1  plain bf16 -- the product configuration (north_star: bf16 operands, fp32 accumulate).
 *   3  fp32 split: every act tensor / packed weight matrix is 3 consecutive bf16 planes [plane][elements] whose
 *      sum is the fp32 value (hi, mid, lo); a pointer passed to this library addresses plane 0 and the planes
 *      follow at a stride of the tensor's own element count.  The SAME kernels then run the 6 cross terms of the
 *      operand planes through tcgen05 into the same fp32 accumulators, the epilogues and memory-bound kernels read
 *      plane sums and write plane splits: fp32-class arithmetic (north_star's <= 1e-4 mode) used by the parity
 *      tests to compare forward and gradients with the fp64 goldens.  ~6x the tensor work; not a product mode.
 * Exact operands (the mask image of dasr_build_mask16, the auxiliary tensor of dasr_build_aux) stay one plane.   */
int dasr_set_planes(int n);
int dasr_get_planes(void);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution (tcgen05 / TMEM / TMA), stride 1, square kernel ks in {1,3,9}, zero padding
 * ks/2.  Replaces nn.Conv2d / weight-normed Conv2d / ConvTranspose2d-as-conv on the hot path
 * (codes/models/modules/sftmd_arch.py:743-749,812,819,862-864,891-910; normalization.py:41-42).
 * A: NHWC bf16 [B,H,W,Cin] (Cin multiple of 32).  Wp: packed weights bf16 [Npad][ks*ks*Cin]
 * (dasr_pack_weights).  bias: fp32 [Npad].
 * ------------------------------------------------------------------------------------------------ */
enum {
    DASR_EPI_STORE = 0,      /* out_bf16 NHWC = act(acc + bias [+ resid])                          */
    DASR_EPI_STATS = 1,      /* out_bf16 NHWC = acc + bias ; stats[b][slot][c][0..1] = partial sum, sum of
                                squares of the stored values (one writer per entry: deterministic)    */
    DASR_EPI_SEAN = 2,       /* acc = [gamma_o | beta_o]; fused SEAN modulate (normalization.py:87-89) */
    DASR_EPI_SHUFFLE2 = 3,   /* PixelShuffle(2) + act folded into the store (sftmd_arch.py:893-908)  */
    DASR_EPI_NCHW_F32 = 4    /* out_f32 NCHW [B,Cout,H,W] = clamp?(acc + bias)  (sftmd_arch.py:948-950) */
};
enum { DASR_ACT_NONE = 0, DASR_ACT_RELU = 1, DASR_ACT_LRELU = 2 };

typedef struct {
    int32_t B, H, W, Cin, Cout;   /* Cout = real output channels                                   */
    int32_t ks;                   /* kernel height (and width when kw == 0): 1, 3 or 9             */
    int32_t epi;                  /* DASR_EPI_*                                                    */
    int32_t act;                  /* DASR_ACT_* applied last                                       */
    int32_t subsample;            /* 1, or 2: keep even (h,w) only -> stride-2 conv output          */
    int32_t clamp01;              /* DASR_EPI_NCHW_F32: clamp to [0,1]                              */
    int32_t inner_relu;           /* DASR_EPI_SEAN: relu before the residual add (norm1 path)       */
    int32_t kw;                   /* kernel width; 0 = ks (square)                                  */
    float mask_slope;             /* DASR_EPI_STORE with actmask: factor where actmask <= 0         */
    int32_t w_img_rows;           /* 0, or Cout: per-image weights -- image b uses rows [b*Cout, (b+1)*Cout) of w (a
                                     batch of independent GEMMs in one launch: the 26 style-table GEMMs)  */
    int32_t unshuffle;            /* DASR_EPI_STORE only, 0 | 2: store pixel (h, w) of the [B,H,W,Cout] result at
                                     out[b, h/2, w/2, (2*(h%2) + w%2)*Cout + c] (out NHWC [B,H/2,W/2,4*Cout]) -- the
                                     backward of PixelShuffle(2) as store addressing; resid / actmask keep the
                                     [B,H,W,Cout] indexing (the LeakyReLU mask of the shuffled tensor)        */
} dasr_conv_desc;

typedef struct {
    const void* x;        /* A operand, NHWC bf16                                                   */
    const void* w;        /* packed weights                                                         */
    const float* bias;    /* [Npad]                                                                 */
    void* out;            /* bf16 NHWC (or fp32 NCHW for DASR_EPI_NCHW_F32)                          */
    const void* resid;    /* optional NHWC bf16 residual, same shape as out                         */
    float* stats;         /* DASR_EPI_STATS: [B][dasr_conv_stats_slots()][Cout][2] partial sums (every
                             entry is written; no zeroing needed)                                   */
    const void* y;        /* DASR_EPI_SEAN: conv output to normalise, NHWC bf16 [B,H,W,Cout/2]      */
    const float* norm;    /* DASR_EPI_SEAN: [B][Cout/2][2] = (mean, scale) from dasr_instats_finalize; or NULL with
                             `stats` = the partial sums [B][slots][Cout/2][2] written by the DASR_EPI_STATS conv of
                             the same geometry: the finalize step then runs inside this kernel (per tile, in slot
                             order, hidden behind the MMAs)                                              */
    const void* gb_s;     /* DASR_EPI_SEAN: dynamic-conv term NHWC bf16 [B,H,W,Cout] (or NULL)       */
    const void* actmask;  /* DASR_EPI_STORE: optional NHWC bf16 tensor shaped like out; the result is multiplied
                             by (actmask > 0 ? 1 : mask_slope) last -- ReLU / LeakyReLU backward fused into a
                             data-gradient convolution                                               */
    void* gamma_out;      /* DASR_EPI_SEAN: optional NHWC bf16 [B,H,W,Cout/2] copy of gamma (saved for backward) */
    const float* resid_f32; /* DASR_EPI_SEAN: fp32 NHWC residual (the trunk's fp32 residual stream); used
                               instead of `resid` when not NULL                                     */
    float* out_aux_f32;   /* DASR_EPI_SEAN: optional fp32 NHWC copy of the output                    */
    float* norm_out;      /* DASR_EPI_SEAN with fused finalize: optional [B][Cout/2][2] copy of (mean, scale) ...   */
    float* normk_out;     /* ... and [B][Cout/2] of k (see dasr_instats_finalize) for the backward pass            */
    const float* gen_depth; /* DASR_EPI_SEAN + dyn_x, Cin = Cout = 128: generate the A operand inside the kernel instead of
                               reading x:  x = ReLU(conv3x3(gen_depth [B,1,H,W] fp32, gen_w [Cin][9]) + gen_b [Cin])
                               -- SEAN's mlp_mask (normalization.py:37-40,61); x may then be NULL (inference)        */
    const float* gen_w;
    const float* gen_b;
    const void* dyn_x;    /* DASR_EPI_SEAN: K-DYN folded into the GEMM (instead of gb_s): the depth-mask image NHWC
                             bf16 [B,H,W,16] (dasr_build_mask16) ...                                            */
    const void* dyn_w;    /* ... and the per-image dynamic filters bf16 [B*Cout][ks*ks*16] (dasr_table_to_dynweights):
                             gamma/beta += sum_{tap,k} mask[p+tap][k] * T[b][k][tap][:]                          */
} dasr_conv_args;

int dasr_conv_fwd(const dasr_conv_desc* d, const dasr_conv_args* a, void* stream);
/* CTA-pair kernels (tcgen05.mma.cta_group::2, M = 256: the two CTAs of a pair work on the same tile position of two
 * consecutive images and share one weight stream): the 3x3 128 -> 128 DASR_EPI_SEAN convolution with the K-DYN
 * extension, the 64 -> 64 DASR_EPI_STATS convolution and dasr_conv_out9 use them when the batch has >= 2 images and
 * enough tiles to occupy every SM pair.  on = 0 keeps everything on single CTAs, 1 allows pairs, -1 = environment
 * DASR_SEAN_PAIR (default: pairs).  Both forms compute the same sums in the same order (A/B measurements, tests). */
int dasr_set_sean_pair(int on);
/* number of partial-statistics slots per image the DASR_EPI_STATS epilogue writes for this shape (> 0),
 * or a negative dasr_status                                                                         */
int dasr_conv_stats_slots(const dasr_conv_desc* d);
/* 1 if a 3x3, 128 -> 128 DASR_EPI_SEAN convolution over HxW frames can generate its A operand in-kernel (gen_depth):
 * it needs two A stages in shared memory, which wide strips do not leave room for                              */
int dasr_conv_gen_ok(int H, int W);

/* Weight gradient of a stride-1 "same" convolution with a kh x kw kernel (autograd of the nn.Conv2d call sites
 * above; reference codes/models/F_model_depthCond.py:191):
 *   dw[o][(t*kw+u)*Cin + i] += sum_{b,h,w} dy[b,h,w,o] * x[b,h+t-kh/2,w+u-kw/2,i]
 * dy NHWC bf16 [B,H,W,Cout], x NHWC bf16 [B,H,W,Cin], dw fp32 [Cout][kh*kw*Cin] in the packed GEMM-B layout
 * of dasr_pack_weights (accumulated with fp32 atomics: the caller zeroes it).  Channels multiples of 32.
 * db (optional, fp32 [Cout]): the bias gradient db[o] += sum_{b,h,w} dy[b,h,w,o], computed by the same kernel
 * (one more MMA per K step against a block of ones) instead of a separate pass over dy.
 * ksplit_div: 0 / 1 = the latency-optimal split-K (one CTA per SM); d > 1 = d times fewer CTAs, each reducing d
 * times more pixels (fewer partial-dW flushes, fewer SM-microseconds, longer latency): for gradients issued on a
 * side stream beside other kernels.                                                                            */
typedef struct {
    int32_t B, H, W, Cout, Cin, kh, kw, ksplit_div;
} dasr_wgrad_desc;
int dasr_conv_wgrad(const dasr_wgrad_desc* d, const void* dy, const void* x, float* dw, float* db, void* stream);

/* conv_output + clamp (sftmd_arch.py:910,948-950): 9x9, Cin = 32 -> Cout = 3, zero padding 4.
 * x NHWC bf16 [B,H,W,32]; wq packed by DASR_PACK_ROWTAPS ([9][32][32] bf16); bias fp32 [3];
 * out NCHW fp32 [B,3,H,W] = clamp01 ? clamp(conv + bias, 0, 1) : conv + bias                          */
int dasr_conv_out9(const void* x, const void* wq, const float* bias, float* out, int B, int H, int W,
                   int Cout, int clamp01, void* stream);
/* The same convolution with util.tensor2img (codes/utils/util.py:566-590, as codes/test.py:87 calls it) fused into
 * the store: img u8 [B,H,W,3] BGR = round_half_even((clamp(conv + bias, lo, hi) - lo) / (hi - lo) * 255) -- bit-identical
 * to dasr_conv_out9 followed by dasr_tensor2img, without the fp32 frames in HBM (201 MB written and read back per
 * 64-frame batch).                                                                                              */
int dasr_conv_out9_frames(const void* x, const void* wq, const float* bias, uint8_t* img, int B, int H, int W,
                          float lo, float hi, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight preparation: weight-norm (w = g*v/||v||, sftmd_arch.py:740,851), alpha folding and repacking
 * of fp32 [O][I][ks][ks] conv weights into the bf16 K-major GEMM-B layout [rows][ks*ks*I'].
 * ------------------------------------------------------------------------------------------------ */
enum {
    DASR_PACK_CONV = 0,        /* rows = O, K = (tap, I)                                             */
    DASR_PACK_CONVT = 1,       /* source is ConvTranspose2d [I][O][ks][ks]; rows = O, taps flipped   */
    DASR_PACK_STYLE = 2,       /* source [O][I][ks][ks]; dst [taps*rows_per_tap][I], row = tap*rows_per_tap
                                  + row_offset + o, K = I  (style-table GEMM B operand)             */
    DASR_PACK_ROWTAPS = 3,     /* source [O][I][ks][ks], ks*O <= 32; dst [ks][32][I], row = t*32 + u*O + o
                                  (dasr_conv_out9 operand; caller zero-fills dst once)              */
    DASR_PACK_DGRAD = 4,       /* data-gradient operand of a Conv2d [O][I][ks][ks]: dst [I][ks*ks*Ot],
                                  dst[i][tap'*Ot + row_offset + n(o)] = w[o][i][flipped tap'], Ot = rows_per_tap
                                  (0 = O), n(o) = the PixelShuffle permutation when shuffle_r > 1      */
    DASR_PACK_DGRAD_CONVT = 5, /* data-gradient operand of the ConvTranspose2d-as-conv [I][O][ks][ks]:
                                  dst[i][tap*O + o] = w[i][o][tap] (weight-norm scale of row i)        */
    DASR_PACK_OUT9_DGRAD = 6   /* conv_output [3][32][9][9] -> dst [32][9*32], dst[i][t'*32 + u*3 + o] =
                                  w[o][i][8-t'][u]  (vertical 9x1 conv over the im2row gradient)       */
};
typedef struct {
    const float* v;       /* weight or weight_v                                                      */
    const float* g;       /* weight_g ([dim0]) or NULL                                               */
    const float* alpha;   /* device scalar or NULL                                                   */
    const float* bias;    /* [O] or NULL                                                             */
    const float* bias2;   /* [O] or NULL: dst_bias = f*bias + (1-f)*bias2, f = the alpha factor      */
    void* dst;            /* bf16 [rows_total][ks*ks*I]                                              */
    float* dst_bias;      /* fp32 [rows_total] or NULL                                               */
    int32_t dim0, dim1, ks;
    int32_t mode;         /* DASR_PACK_*                                                             */
    int32_t alpha_mode;   /* 0: none, 1: scale by alpha, 2: scale by (1-alpha)                       */
    int32_t shuffle_r;    /* 0, or r: destination row = (i*r+j)*(O/r^2) + c for source row c*r^2+i*r+j */
    int32_t row_offset;   /* destination row offset (stack several convs into one B matrix)          */
    int32_t rows_per_tap; /* DASR_PACK_STYLE only                                                    */
    int64_t dst_plane_stride; /* dasr_set_planes > 1: elements between the planes of dst (its total element count;
                                 several descriptors may fill row ranges of one dst).  Ignored with 1 plane.  */
} dasr_pack_desc;
/* descs: HOST array of n descriptors; scratch: device fp32 [sum of dim0] for the weight-norm scales */
int dasr_pack_weights(const dasr_pack_desc* descs, int n, float* scratch, void* stream);

/* Gradient unpacking: d(packed weight) fp32 -> parameter gradients (weight-norm backward, SEAN alpha blend,
 * PixelShuffle / ConvTranspose index maps).  Mirrors dasr_pack_desc; see pack.cu.                        */
typedef struct {
    const float* dwp;     /* gradient in the packed layout of `mode` (fp32)                              */
    const float* dbias_p; /* packed bias gradient [rows] or NULL                                         */
    const float* v;       /* weight / weight_v (forward value)                                           */
    const float* g;       /* weight_g or NULL                                                            */
    const float* alpha;   /* device scalar or NULL                                                       */
    const float* bias;    /* alpha_mode 2: own bias (b_o)  -- for d alpha                                */
    const float* bias2;   /* alpha_mode 2: the style bias (b_s)                                          */
    float* dv;            /* out: gradient of weight / weight_v, parameter layout                        */
    float* dg;            /* out: gradient of weight_g or NULL                                           */
    float* dbias;         /* out: gradient of bias or NULL                                               */
    float* dbias2;        /* out: gradient of the style bias (alpha_mode 2) or NULL                      */
    float* dalpha;        /* accumulated (atomicAdd): gradient of alpha or NULL                          */
    int32_t dim0, dim1, ks, mode, alpha_mode, shuffle_r, row_offset, rows_per_tap;
    int32_t ipack;        /* packed input channels when padded (encoder.layer1: 32), else 0              */
    int32_t reserved;
} dasr_unpack_desc;
int dasr_unpack_grads(const dasr_unpack_desc* descs, int n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward, memory-bound kernels (backward.cu)
 * ------------------------------------------------------------------------------------------------ */
/* SEAN modulate + double-InstanceNorm backward (normalization.py:56,87-89; sftmd_arch.py:813,820,828,832-833).
 * pass 1: dz = dout * [act_out > 0]; n = (y-mean)*scale; dgb = [dz*n | dz]; dn = dz*(1+gamma); dskip = dz (optional);
 *         part[b][slot][c] = (sum dn, sum dn*n, sum dz*n, sum dz) over the slot's pixels, fp32 [B][slots][nf][4].
 *         All tensors NHWC bf16 [B,HW,nf] but dgb [B,HW,2nf].
 * pass 2: coef = (S1/N, -k*T2/(N*scale)) from part (per CTA); dy = scale*(dn - c1) + c2*n; dbias (optional, fp32
 *         [2nf]) += the bias gradient of the [gamma_o; beta_o] convolution (sum over b, p of dgb).              */
int dasr_sean_bwd_slots(int HW);
int dasr_sean_bwd1(const void* dout, const void* act_out, const void* y, const float* norm, const void* gamma,
                   void* dgb, void* dn, void* dskip, float* part, int B, int HW, int nf, void* stream);
int dasr_sean_bwd2(const void* dn, const void* y, const float* norm, const float* normk, const float* part, void* dy,
                   float* dbias, int B, int HW, int nf, void* stream);
/* out[c] += sum over rows of x[row][c] (x bf16 [rows][C]) -- bias gradients                               */
int dasr_colsum(const void* x, float* out, int64_t rows, int C, void* stream);
/* K-DYN backward: dT[b][k][tap][c] += sum_p dgb[b,p,c] * mask[b,k,p+tap-1] (labels fast path like the forward) */
int dasr_dynconv_bwd(const void* dgb, const uint8_t* labels, const float* masks, const int32_t* flag, float* dT,
                     int B, int K, int H, int W, int nf2, void* stream);
/* Auxiliary input tensor aux NHWC bf16 [B,H,W,DASR_AUX_CH]: channels 0..K-1 = one-hot depth mask (from the label
 * map of dasr_mask_labels), DASR_AUX_DEPTH_HI/LO = the depth split into two bf16 parts (hi + lo = 16 mantissa
 * bits), DASR_AUX_ONE = 1, others 0.  X operand of the two tensor-core kernels below; built once per forward.   */
#define DASR_AUX_CH 32
#define DASR_AUX_DEPTH_HI 16
#define DASR_AUX_DEPTH_LO 17
#define DASR_AUX_ONE 18
#define DASR_AUX_DEPTH_LO2 19   /* third bf16 part: hi + lo + lo2 reproduces the fp32 depth exactly */
int dasr_build_aux(const uint8_t* labels, const float* depth, void* aux, int B, int K, int H, int W, void* stream);
/* K-DYN backward on tcgen05 (one-hot masks): same result as dasr_dynconv_bwd, computed as the per-image weight
 * gradient of a 3x3 convolution over the one-hot channels of aux.  No-op when *flag != 0; the caller then also
 * issues dasr_dynconv_bwd(labels = NULL, masks, flag), which is a no-op when *flag == 0.                        */
int dasr_dynconv_bwd_tc(const void* dgb, const void* aux, const int32_t* flag, float* dT, int B, int K, int H, int W,
                        int nf2, void* stream);
/* mlp_mask backward on tcgen05: same result as dasr_actv_bwd; scratch fp32 [C][9*DASR_AUX_CH] zeroed by the caller */
int dasr_actv_bwd_tc(const void* dA, const void* aux, float* scratch, float* dW, float* db, int B, int H, int W, int C,
                     void* stream);
/* style-table GEMM backward: dWs[n][c] = sum_bk dT[bk][n] stp[bk][c];  dstp[bk][c] = sum_n dT[bk][n] Ws[n][c]    */
int dasr_table_bwd(const float* dT, const void* stp, const void* Ws, float* dWs, float* dstp, int BK, int N, int L,
                   void* stream);
/* the same for nS SEAN instances at once (instance strides: dT BK*N, stp BK*L, Ws / dWs N*L, dstp BK*L)      */
int dasr_table_bwd_batched(const float* dT, const void* stp, const void* Ws, float* dWs, float* dstp, int nS, int BK,
                           int N, int L, void* stream);
/* The two GEMMs of dasr_table_bwd_batched separately (parts: 1 = dWs, 2 = dstp, 3 = both): dWs only reaches parameter
 * gradients, so the training step issues it on a side stream beside the dstp -> A_i_j -> encoder chain.            */
int dasr_table_bwd_parts(const float* dT, const void* stp, const void* Ws, float* dWs, float* dstp, int nS, int BK,
                         int N, int L, int parts, void* stream);
/* A_i_j backward: dA += , da += , dvec += (accumulating over the SEAN instances)                        */
int dasr_style_mix_bwd(const float* dstp, const float* vec, const float* A, float* dA, float* da, float* dvec, int B,
                       int K, int L, void* stream);
/* all nS instances at once: dstp fp32 [nS][B][K][L]; A_ptrs / dA_ptrs / da_ptrs are DEVICE arrays of nS device
 * pointers (A_i_j.weight, its gradient, the gradient of A_i_j.bias); dvec += the sum over the instances          */
int dasr_style_mix_bwd_batched(const float* dstp, const float* vec, const void* A_ptrs, const void* dA_ptrs,
                               const void* da_ptrs, float* dvec, int nS, int B, int K, int L, void* stream);
int dasr_region_pool_bwd(const float* dvec, const float* msel, const float* cnt, void* de5, int B, int P, int C, int K,
                         void* stream);
/* mlp_mask backward: dW[c][9] += , db[c] += from dA (NHWC bf16 [B,H,W,C], ReLU mask already applied)      */
int dasr_actv_bwd(const void* dA, const float* depth, float* dW, float* db, int B, int H, int W, int C, void* stream);
/* PixelShuffle(r) + LeakyReLU backward (r = 2 | 3): dps/ps_out NHWC bf16 [B,rH,rW,Cq] -> dconv NHWC bf16
 * [B,H,W,r*r*Cq] in the shuffled channel order s*Cq + c, s = r*i + j (DASR_PACK_* shuffle_r)                 */
int dasr_unshuffle_actgrad(const void* dps, const void* ps_out, void* dconv, int B, int H, int W, int Cq, float slope,
                           int r, void* stream);
/* PixelShuffle(r) forward of an NHWC bf16 tensor in that channel order (sftmd_arch.py:904-908; the x3 tail --
 * for r = 2 the shuffle is the store addressing of DASR_EPI_SHUFFLE2): in [B,H,W,r*r*Cq] -> out [B,rH,rW,Cq]  */
int dasr_pixel_shuffle(const void* in, void* out, int B, int H, int W, int Cq, int r, void* stream);
/* clamp backward + horizontal im2row of d(sr): dout, sr NCHW fp32 [B,3,H,W] -> aprime NHWC bf16 [B,H,W,32];
 * dbias[3] += sum of the masked gradient                                                                 */
int dasr_out9_bwd_prep(const float* dout, const float* sr, void* aprime, float* dbias, int B, int H, int W, void* stream);
int dasr_nchw3_to_nhwc32(const float* x, void* out, int B, int H, int W, void* stream);
/* out = d * (act_out > 0 ? 1 : slope)  (ReLU / LeakyReLU backward; n bf16 elements)                       */
int dasr_actgrad(const void* d, const void* act_out, void* out, int64_t n, float slope, void* stream);
/* zero-stuffed copy onto an [Ho,Wo] grid (gradient of a stride-2 conv on the stride-1 grid)               */
int dasr_zero_insert2_to(const void* x, void* out, int B, int H, int W, int C, int Ho, int Wo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout / small ops
 * ------------------------------------------------------------------------------------------------ */
/* encoder.layer1 (weight-normed 3x3 conv Cin=3 -> 32) + LeakyReLU(0.2): NCHW fp32 -> NHWC bf16
 * (sftmd_arch.py:743,772,783).  v: fp32 [32][3][3][3] (weight_v), g: [32] (weight_g, NULL = plain
 * weight), bias [32].                                                                              */
int dasr_conv_first(const float* x_nchw, const float* v, const float* g, const float* bias,
                    void* out_nhwc, int B, int H, int W, void* stream);
/* zero-insertion upsample for the transposed conv (sftmd_arch.py:748): [B,H,W,C] -> [B,2H-1,2W-1,C]   */
int dasr_zero_insert2(const void* x, void* out, int B, int H, int W, int C, void* stream);
/* out = a + b (bf16 NHWC, n elements)  -- feat_add1 (sftmd_arch.py:931); a32 (fp32) replaces a when
 * not NULL (the trunk's fp32 residual stream)                                                        */
int dasr_add(const void* a, const float* a32, const void* b, void* out, int64_t n, void* stream);

/* RegionWiseAvgPooling (sftmd_arch.py:714-733): e5 NHWC bf16 [B,hf,wf,C], masks NCHW fp32 [B,K,H,W]
 * -> depthVec fp32 [B,K,C]                                                                          */
int dasr_region_pool_fwd(const void* e5, const float* masks, float* depth_vec, float* msel, float* cnt, int B,
                         int hf, int wf, int C, int K, int H, int W, void* stream);
/* msel [B,K,hf*wf] / cnt [B,K]: optional (NULL) copies of the thresholded masks and their sums for the backward */

/* masks NCHW fp32 [B,K,H,W] -> labels u8 [B,H,W] (k if one-hot at k, 255 if all zero); *flag_not_onehot
 * (device int, caller zeroes) is set when some pixel is neither                                      */
int dasr_mask_labels(const float* masks, uint8_t* labels, int32_t* flag_not_onehot, int B, int K, int H,
                     int W, void* stream);

/* SEAN depth branch first layer (normalization.py:37-40,61): actv = ReLU(conv3x3(depth, 1->C) + b),
 * depth NCHW fp32 [B,1,H,W], w fp32 [C][9], out NHWC bf16 [B,H,W,C].  ctas_per_sm: 0 = a persistent grid that
 * fills the device; 1 = one block per SM, compiled for 64 registers, meant to run on a side stream NEXT TO a
 * convolution kernel (the SEAN convolution leaves 16 K registers and ~5 KB of shared memory per SM free)   */
int dasr_actv_fwd(const float* depth, const float* w, const float* bias, void* out, int B, int H, int W,
                  int C, int ctas_per_sm, void* stream);

/* SEAN label mixing (normalization.py:27,80): stp[b][j][:] = sum_i A[j][i] depth_vec[b][i][:] + a[j]
 * depth_vec fp32 [B,K,L] -> stp bf16 [B*K][L].  The style table (normalization.py:81-85 restated,
 * SURVEY 8a-7b)  T[b][k][tap][o] = alpha_x * sum_c W_x[o][c][tap] stp[b][k][c]  is then ONE 1x1
 * dasr_conv_fwd over the "image" [1,1,B*K,L] with weights packed by DASR_PACK_STYLE
 * (rows = tap*2nf + o, gamma rows o in [0,nf), beta rows o in [nf,2nf)) -> T bf16 [B][K][9][2nf].     */
int dasr_style_mix(const float* depth_vec, const float* A, const float* a, void* stp, int B, int K,
                   int L, void* stream);

/* The same for ALL nS SEAN instances of the network in one launch: A_ptrs / a_ptrs are DEVICE arrays of nS device
 * pointers (A_i_j.weight [K][K], A_i_j.bias [K]); stp bf16 [nS][B*K][L].  The nS style tables are then one
 * dasr_conv_fwd with per-image weights (w_img_rows) over the "batch" of nS instances.                      */
int dasr_style_mix_batched(const float* depth_vec, const void* A_ptrs, const void* a_ptrs, void* stp, int nS, int B,
                           int K, int L, void* stream);

/* K-DYN, the depth-guided dynamic convolution apply step:
 *   gb_s[b,p,:] = sum_tap T[b][label(p+tap)][tap][:]          (one-hot masks: labels given, *flag == 0)
 *   gb_s[b,p,:] = sum_{k,tap} mask[b,k,p+tap] T[b][k][tap][:]   (general masks: labels == NULL or *flag != 0;
 *                                                              flag is the device int of dasr_mask_labels)
 * equal to [mlp_gamma_s(style_map) ; mlp_beta_s(style_map)] without bias (bias is merged into the conv
 * bias by dasr_pack_weights).  table bf16 [B][K][9][2nf]; out NHWC bf16 [B,H,W,2nf].                  */
int dasr_dynconv_fwd(const void* table, const uint8_t* labels, const float* masks, const int32_t* flag,
                     void* out, int B, int K, int H, int W, int nf2, void* stream);

/* K-DYN folded into the SEAN GEMM (dasr_conv_args.dyn_x / dyn_w).  dasr_build_mask16: masks NCHW fp32 [B,K,H,W]
 * (K <= 16) -> NHWC bf16 [B,H,W,16] (one-hot masks are exact in bf16; other values are rounded like every other
 * activation).  dasr_table_to_dynweights: the n = (instances x images) tables bf16 [n][K][9][2nf] -> GEMM-B
 * weights bf16 [n][2nf][9*16] (column tap*16 + k, zero for k >= K).  group (0 = n): images per weight matrix
 * (one SEAN instance) -- only matters with dasr_set_planes > 1, where the planes of wdyn sit inside every
 * group ([n/group][plane][group][2nf][9*16]) so that one instance's slice is a self-contained dyn_w operand.      */
int dasr_build_mask16(const float* masks, void* mask16, int B, int K, int H, int W, void* stream);
int dasr_table_to_dynweights(const void* table, void* wdyn, int n, int K, int nf2, int group, void* stream);

/* InstanceNorm statistics (sftmd_arch.py:813,820 + normalization.py:17,56 = IN applied twice):
 * stats [B][nslots][C][2] (partial sum, sumsq over H*W; summed here in slot order) ->
 * norm [B][C][2] = (mean, (v+eps)^-1/2 (v/(v+eps)+eps)^-1/2)                                          */
int dasr_instats_finalize(const float* stats, float* norm, float* normk, int B, int C, int HW, int nslots,
                          void* stream);
/* normk [B][C] (optional, NULL): k = 1/a + eps/(a^2 r) with a = v+eps, r = v/a+eps, used by dasr_sean_bwd2 */

/* ------------------------------------------------------------------------------------------------
 * Training step behind the generator (train.cu)
 * ------------------------------------------------------------------------------------------------ */
#define DASR_LOSS_KMAX 16      /* most depth masks the loss kernels keep in registers                  */
#define DASR_LOSS_ROW 36       /* floats per partial-sum row: [0] sum|SR-HR|, [1..16] sum SmoothL1 of mask k,
                                  [17..32] sum of mask k over the HR grid, [33..35] padding              */
/* K-LOSS forward.  Replaces nn.L1Loss (codes/models/F_model_depthCond.py:52,164) and
 * dynamic_weight_mask_loss.forward (codes/models/modules/mask_loss.py:64-90) with ONE pass over SR/HR.
 * sr, hr: NCHW fp32 [B,C,Ho,Wo] (C <= 4, Wo % 4 == 0); labels u8 [B,h,w] + flag from dasr_mask_labels (one-hot fast
 * path) and/or masks NCHW fp32 [B,K,h,w] (general path, taken when labels == NULL or *flag != 0); the masks are
 * resized to the HR grid like F.interpolate(mode='nearest').  part: fp32 [dasr_loss_rows()][DASR_LOSS_ROW] scratch;
 * sums: fp32 [DASR_LOSS_ROW] = column sums of part, rows added in order (bit-reproducible).  A data-parallel
 * caller that wants the reference's single-process "ratio of batch sums" all-reduces `sums` before finalize.   */
int dasr_loss_rows(int B, int Ho, int Wo);
int dasr_loss_fwd(const float* sr, const float* hr, const uint8_t* labels, const float* masks, const int32_t* flag,
                  float* part, float* sums, int B, int C, int K, int h, int w, int Ho, int Wo, void* stream);
/* sums -> out fp32 [4 + 4*DASR_LOSS_KMAX]: [0] total = l_pix + l_dyn, [1] l_pix = w_pix * sums[0] / n_elems,
 * [2] l_dyn = w_dyn * sum_k softmax(wdyn)_k * loss_k, [3] w_pix / n_elems, then four arrays of DASR_LOSS_KMAX:
 * loss_k = sums[1+k] / (C * sums[17+k]) (0/0 = NaN for an empty mask, like the reference), softmax(wdyn)_k,
 * the backward coefficient w_dyn * softmax_k / (C * sums[17+k]), and d l_dyn / d wdyn_k.
 * wdyn: device fp32 [K] (dynamic_weight_mask_loss.trainable_weight, mask_loss.py:62) or NULL (= zeros).
 * n_elems = number of SR elements the L1 mean runs over (B*C*Ho*Wo; the global count under data parallelism). */
int dasr_loss_finalize(const float* sums, const float* wdyn, float* out, int K, int C, double n_elems, float w_pix,
                       float w_dyn, void* stream);
/* dsr = gp * out[3] * sign(SR-HR) + gd * sum_k coef_k * m_k * clamp(m_k*SR - m_k*HR, -1, 1)
 * -- autograd of the two criteria (F_model_depthCond.py:191).  g: device fp32 [3] = upstream gradients of
 * (total, l_pix, l_dyn), or NULL (= 1,0,0); gp = use_pix*(g[0]+g[1]), gd = use_dyn*(g[0]+g[2]).
 * dwdyn (optional): fp32 [K] = gd * d l_dyn / d wdyn.                                                     */
int dasr_loss_bwd(const float* sr, const float* hr, const uint8_t* labels, const float* masks, const int32_t* flag,
                  const float* out, const float* g, float use_pix, float use_dyn, float* dsr, float* dwdyn, int B,
                  int C, int K, int h, int w, int Ho, int Wo, void* stream);
/* K-ADAM: one torch.optim.Adam step (F_model_depthCond.py:99-101,192; amsgrad off) over flat fp32 buffers of n
 * elements: g' = grad_scale*g + wd*p; m = lerp(m, g', 1-beta1); v = beta2*v + (1-beta2)*g'^2;
 * p -= lr/(1-beta1^step) * m / (sqrt(v)/sqrt(1-beta2^step) + eps).  step counts from 1.  16-byte aligned buffers.
 * dev_scalars (optional): device fp32 [2] = { lr/(1-beta1^step), sqrt(1-beta2^step) } read by the kernel instead of
 * the values derived from lr/step -- lets a captured CUDA graph of the training step be replayed for every step. */
int dasr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                   double eps, double weight_decay, int64_t step, double grad_scale, const float* dev_scalars,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input / output steps either side of the generator (io.cu)
 * ------------------------------------------------------------------------------------------------ */
/* getDepthMask (codes/data/LQGTker_Depth_dataset.py:204-226) on the device: depth NCHW fp32 [B,1,H,W] -> labels u8
 * [B,H,W] (bin index, 255 = in no bin -- the pixel(s) at the image maximum, whose upper edge is exclusive) and,
 * when masks != NULL, the reference's one-hot fp32 planes [B,K,H,W].  fixed_range = depthFixedRange (bins over
 * [0,1]); otherwise per-image (min,max), written to range_out [B][2] when not NULL.  Same fp32 roundings as torch. */
int dasr_depth_masks(const float* depth, uint8_t* labels, float* masks, float* range_out, int B, int K, int H, int W,
                     int fixed_range, void* stream);
/* F.interpolate(x, size=(Ho,Wo), mode='nearest') with integer ratios, NCHW fp32 planes [planes][H][W] ->
 * [planes][Ho][Wo]: what a SEAN instance above LR resolution applies to the depth map and the depth masks
 * (codes/models/modules/normalization.py:58-59; which_ResBlk_depth containing nb-2 / nb-1 at x8 / x4).            */
int dasr_nearest_up(const float* in, float* out, int planes, int H, int W, int Ho, int Wo, void* stream);
/* tensor2img (codes/utils/util.py:566-590) per frame: sr NCHW fp32 [B,3,H,W] RGB -> img u8 [B,H,W,3] BGR,
 * round_half_even((clamp(x, lo, hi) - lo) / (hi - lo) * 255)                                                     */
int dasr_tensor2img(const float* sr, uint8_t* img, int B, int H, int W, float lo, float hi, void* stream);

/* Validation metrics.
 * dasr_sqdiff_u8: out[f] = exact integer sum of (a-b)^2 over frame f ([F,H,W,C] uint8) inside a `crop`-pixel border --
 * the numerator of util.calculate_psnr (codes/utils/util.py:646-653) as train.py:251-257 calls it.
 * dasr_ssim: pytorch_ssim.ssim (codes/pytorch_ssim/__init__.py:17-38,65-72; 11x11 Gaussian window, sigma 1.5, zero
 * padding, C1 = 0.01^2, C2 = 0.03^2) of NCHW fp32 frames [F,C,H,W]; out[f] = mean SSIM of frame f.
 * part: fp32 scratch [F*C*dasr_ssim_tiles(H,W)].                                                                  */
int dasr_sqdiff_u8(const uint8_t* a, const uint8_t* b, unsigned long long* out, int F, int H, int W, int C, int crop,
                   void* stream);
/* the same plus psnr[f] = 20 log10(255 / sqrt(ssd[f] / n)) in double (+inf for identical frames)                  */
int dasr_psnr_u8(const uint8_t* a, const uint8_t* b, unsigned long long* ssd, double* psnr, int F, int H, int W, int C,
                 int crop, void* stream);
int dasr_ssim_tiles(int H, int W);
int dasr_ssim(const float* img1, const float* img2, float* part, float* out, int F, int C, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DASR_H_ */

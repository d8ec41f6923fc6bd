"""ctypes binding of libdasr_b200.so (C ABI declared in include/dasr.h).

The product path has NO CPU fallback: if the shared library is missing, or a tensor is not a contiguous CUDA
tensor of the expected dtype, these wrappers raise.  Every launch goes to torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DASR_LIB_PATH: developer override (tools/prof_stalls.py loads the -DDASR_PROFILE build)
LIB_PATH = os.environ.get("DASR_LIB_PATH") or os.path.join(_HERE, "libdasr_b200.so")

# --- enums (include/dasr.h)
EPI_STORE, EPI_STATS, EPI_SEAN, EPI_SHUFFLE2, EPI_NCHW_F32 = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
PACK_CONV, PACK_CONVT, PACK_STYLE, PACK_ROWTAPS, PACK_DGRAD, PACK_DGRAD_CONVT, PACK_OUT9_DGRAD = 0, 1, 2, 3, 4, 5, 6


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("B", "H", "W", "Cin", "Cout", "ks", "epi", "act", "subsample", "clamp01", "inner_relu", "kw")] + \
               [("mask_slope", C.c_float), ("w_img_rows", C.c_int32), ("unshuffle", C.c_int32)]


class ConvArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("x", "w", "bias", "out", "resid", "stats", "y", "norm", "gb_s", "actmask",
                                          "gamma_out", "resid_f32", "out_aux_f32", "norm_out", "normk_out", "gen_depth", "gen_w",
                                          "gen_b", "dyn_x", "dyn_w")]


class UnpackDesc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dwp", "dbias_p", "v", "g", "alpha", "bias", "bias2", "dv", "dg", "dbias",
                                          "dbias2", "dalpha")] + \
               [(n, C.c_int32) for n in ("dim0", "dim1", "ks", "mode", "alpha_mode", "shuffle_r", "row_offset",
                                         "rows_per_tap", "ipack", "reserved")]


class WgradDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "H", "W", "Cout", "Cin", "kh", "kw", "ksplit_div")]


class PackDesc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("v", "g", "alpha", "bias", "bias2", "dst", "dst_bias")] + \
               [(n, C.c_int32) for n in ("dim0", "dim1", "ks", "mode", "alpha_mode", "shuffle_r", "row_offset",
                                         "rows_per_tap")] + [("dst_plane_stride", C.c_int64)]


_lib = None


def load() -> C.CDLL:
    """Load the shared library once; fail loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libdasr_b200.so is missing (%s). Build it with `python __graft_entry__.py build` -- the B200 path has "
            "no CPU/PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.dasr_last_error.restype = C.c_char_p
    lib.dasr_version.restype = C.c_int
    lib.dasr_launch_count.restype = C.c_int64
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    sigs = {
        "dasr_check_device": [],
        "dasr_set_planes": [i32],
        "dasr_set_sean_pair": [i32],
        "dasr_get_planes": [],
        "dasr_conv_fwd": [C.POINTER(ConvDesc), C.POINTER(ConvArgs), vp],
        "dasr_conv_out9": [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "dasr_conv_out9_frames": [vp, vp, vp, vp, i32, i32, i32, C.c_float, C.c_float, vp],
        "dasr_conv_wgrad": [C.POINTER(WgradDesc), vp, vp, vp, vp, vp],
        "dasr_pack_weights": [C.POINTER(PackDesc), i32, vp, vp],
        "dasr_conv_first": [vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "dasr_zero_insert2": [vp, vp, i32, i32, i32, i32, vp],
        "dasr_add": [vp, vp, vp, vp, i64, vp],
        "dasr_conv_stats_slots": [C.POINTER(ConvDesc)],
        "dasr_conv_gen_ok": [i32, i32],
        "dasr_region_pool_fwd": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
        "dasr_unpack_grads": [C.POINTER(UnpackDesc), i32, vp],
        "dasr_sean_bwd_slots": [i32],
        "dasr_sean_bwd1": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "dasr_sean_bwd2": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "dasr_colsum": [vp, vp, i64, i32, vp],
        "dasr_dynconv_bwd": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "dasr_table_bwd": [vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "dasr_style_mix_bwd": [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "dasr_table_bwd_batched": [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "dasr_table_bwd_parts": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "dasr_style_mix_bwd_batched": [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "dasr_region_pool_bwd": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "dasr_actv_bwd": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "dasr_unshuffle_actgrad": [vp, vp, vp, i32, i32, i32, i32, C.c_float, i32, vp],
        "dasr_pixel_shuffle": [vp, vp, i32, i32, i32, i32, i32, vp],
        "dasr_out9_bwd_prep": [vp, vp, vp, vp, i32, i32, i32, vp],
        "dasr_nchw3_to_nhwc32": [vp, vp, i32, i32, i32, vp],
        "dasr_actgrad": [vp, vp, vp, i64, C.c_float, vp],
        "dasr_zero_insert2_to": [vp, vp, i32, i32, i32, i32, i32, i32, vp],
        "dasr_mask_labels": [vp, vp, vp, i32, i32, i32, i32, vp],
        "dasr_actv_fwd": [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "dasr_style_mix": [vp, vp, vp, vp, i32, i32, i32, vp],
        "dasr_dynconv_fwd": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "dasr_instats_finalize": [vp, vp, vp, i32, i32, i32, i32, vp],
        "dasr_style_mix_batched": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "dasr_build_mask16": [vp, vp, i32, i32, i32, i32, vp],
        "dasr_table_to_dynweights": [vp, vp, i32, i32, i32, i32, vp],
        "dasr_build_aux": [vp, vp, vp, i32, i32, i32, i32, vp],
        "dasr_dynconv_bwd_tc": [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "dasr_actv_bwd_tc": [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "dasr_depth_masks": [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "dasr_tensor2img": [vp, vp, i32, i32, i32, C.c_float, C.c_float, vp],
        "dasr_nearest_up": [vp, vp, i32, i32, i32, i32, i32, vp],
        "dasr_sqdiff_u8": [vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "dasr_psnr_u8": [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "dasr_ssim_tiles": [i32, i32],
        "dasr_ssim": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "dasr_loss_rows": [i32, i32, i32],
        "dasr_loss_fwd": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
        "dasr_loss_finalize": [vp, vp, vp, i32, i32, C.c_double, C.c_float, C.c_float, vp],
        "dasr_loss_bwd": [vp, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
        "dasr_adam_step": [vp, vp, vp, vp, i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, i64,
                           C.c_double, vp, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


EXPORTED = ["dasr_last_error", "dasr_version", "dasr_launch_count", "dasr_check_device", "dasr_set_planes", "dasr_get_planes", "dasr_set_sean_pair", "dasr_conv_fwd", "dasr_conv_stats_slots", "dasr_conv_gen_ok", "dasr_conv_out9", "dasr_conv_out9_frames", "dasr_conv_wgrad", "dasr_pack_weights",
            "dasr_conv_first", "dasr_zero_insert2", "dasr_add", "dasr_region_pool_fwd", "dasr_mask_labels",
            "dasr_actv_fwd", "dasr_style_mix", "dasr_dynconv_fwd", "dasr_instats_finalize", "dasr_unpack_grads",
            "dasr_sean_bwd_slots", "dasr_sean_bwd1", "dasr_sean_bwd2", "dasr_colsum",
            "dasr_dynconv_bwd", "dasr_table_bwd", "dasr_style_mix_bwd", "dasr_region_pool_bwd", "dasr_actv_bwd",
            "dasr_unshuffle_actgrad", "dasr_pixel_shuffle", "dasr_out9_bwd_prep", "dasr_nchw3_to_nhwc32", "dasr_actgrad",
            "dasr_zero_insert2_to", "dasr_loss_rows", "dasr_loss_fwd", "dasr_loss_finalize", "dasr_loss_bwd",
            "dasr_adam_step", "dasr_table_bwd_batched", "dasr_table_bwd_parts", "dasr_style_mix_bwd_batched", "dasr_build_aux", "dasr_build_mask16", "dasr_table_to_dynweights", "dasr_depth_masks", "dasr_tensor2img", "dasr_nearest_up", "dasr_sqdiff_u8", "dasr_psnr_u8", "dasr_ssim_tiles", "dasr_ssim", "dasr_style_mix_batched", "dasr_dynconv_bwd_tc", "dasr_actv_bwd_tc"]
AUX_CH = 32
LOSS_KMAX, LOSS_ROW = 16, 36


# Flat gradient buffers produced by Engine._finish_backward: (weakref to the fp32 buffer, {id(param): offset}).
# autograd hands the views to ``.grad`` without copying, so FusedAdam can recognise them by address and step the
# whole network with one launch straight from the buffer.
_FLAT_GRADS = []


def register_flat_grad(flat: torch.Tensor, layout: dict) -> None:
    import weakref
    _FLAT_GRADS[:] = [(r, l) for r, l in _FLAT_GRADS if r() is not None][-7:]
    _FLAT_GRADS.append((weakref.ref(flat), layout))


def flat_grads():
    return [(r(), l) for r, l in _FLAT_GRADS if r() is not None]


def flat_pad(numel: int) -> int:
    """Elements one parameter occupies in the flat gradient / optimiser buffers (slices stay 16-byte aligned)."""
    return (numel + 3) // 4 * 4


_replayed_launches = 0


def note_replayed_launches(n: int) -> None:
    """Kernels of this library executed by a CUDA-graph replay (counted at capture time; a replay does not pass
    through the C entry points that bump dasr_launch_count)."""
    global _replayed_launches
    _replayed_launches += int(n)


def launch_count() -> int:
    """Kernels of libdasr_b200.so launched so far in this process, directly or through CUDA-graph replays."""
    return int(load().dasr_launch_count()) + _replayed_launches


# ------------------------------------------------------------------------------------------------ act tensors
# "act" tensors (NHWC bf16 activations, packed weights).  With dasr_set_planes(3) -- the fp32-split precise mode the
# parity tests use, see include/dasr.h -- every act tensor is 3 consecutive bf16 planes; the host code keeps working
# with the plane-0 view (same shape as in the product configuration) and the kernels find the other planes at a
# stride of the tensor's own element count.
def planes() -> int:
    return int(load().dasr_get_planes())


def set_planes(n: int) -> None:
    check(load().dasr_set_planes(int(n)))


def act_empty(*shape, device) -> torch.Tensor:
    n = planes()
    if n == 1:
        return torch.empty(*shape, device=device, dtype=torch.bfloat16)
    return torch.empty(n, *shape, device=device, dtype=torch.bfloat16)[0]


def act_zeros(*shape, device) -> torch.Tensor:
    n = planes()
    if n == 1:
        return torch.zeros(*shape, device=device, dtype=torch.bfloat16)
    return torch.zeros(n, *shape, device=device, dtype=torch.bfloat16)[0]


def act_like(t: torch.Tensor) -> torch.Tensor:
    return act_empty(*t.shape, device=t.device)


def act_planes(t: torch.Tensor) -> torch.Tensor:
    """[planes, *shape] view of an act tensor allocated by act_empty / act_zeros (tests / probes)."""
    n = planes()
    return torch.as_strided(t, (n,) + tuple(t.shape), (t.numel(),) + tuple(t.stride()), t.storage_offset())


def act_value(t: torch.Tensor) -> torch.Tensor:
    """fp32 value of an act tensor: the sum of its planes, lowest plane first (exact)."""
    pl = act_planes(t).float()
    v = pl[-1].clone()
    for k in range(pl.shape[0] - 2, -1, -1):
        v += pl[k]
    return v


def act_from(x: torch.Tensor) -> torch.Tensor:
    """act tensor holding the fp32 tensor ``x`` (bf16-rounded with one plane, exact with three)."""
    t = act_empty(*x.shape, device=x.device)
    pl = act_planes(t)
    r = x.float().clone()
    for k in range(pl.shape[0]):
        pl[k] = r.to(torch.bfloat16)
        r -= pl[k].float()
    return t


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("libdasr_b200: %s (status %d)" % (load().dasr_last_error().decode(), rc))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor], dtype=None) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libdasr_b200 needs CUDA tensors (got %s); there is no CPU fallback" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("libdasr_b200 needs contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError("expected dtype %s, got %s" % (dtype, t.dtype))
    return t.data_ptr()


# ------------------------------------------------------------------------------------------------ wrappers
def conv_fwd(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, out: torch.Tensor, *, Cout: int, ks: int,
             epi: int = EPI_STORE, act: int = ACT_NONE, subsample: int = 1, clamp01: int = 0, inner_relu: int = 0,
             resid=None, stats=None, y=None, norm=None, gb_s=None, resid_f32=None, out_aux_f32=None, kw: int = 0,
             actmask=None, mask_slope: float = 0.0, gamma_out=None, w_img_rows: int = 0, dyn_x=None, dyn_w=None, norm_out=None, normk_out=None, gen_depth=None, gen_w=None, gen_b=None,
             shape=None, unshuffle: int = 0) -> torch.Tensor:
    """x: NHWC bf16 [B,H,W,Cin] (or None with gen_depth / shape = (B,H,W,Cin): the kernel generates its A operand)."""
    B, H, W, Cin = x.shape if x is not None else shape
    d = ConvDesc(B, H, W, Cin, Cout, ks, epi, act, subsample, clamp01, inner_relu, kw, mask_slope, w_img_rows,
                 unshuffle)
    a = ConvArgs(ptr(x, torch.bfloat16) if x is not None else None, ptr(w, torch.bfloat16), ptr(bias, torch.float32), ptr(out), ptr(resid),
                 ptr(stats), ptr(y), ptr(norm), ptr(gb_s), ptr(actmask, torch.bfloat16),
                 ptr(gamma_out, torch.bfloat16), ptr(resid_f32, torch.float32), ptr(out_aux_f32, torch.float32),
                 ptr(norm_out, torch.float32), ptr(normk_out, torch.float32),
                 ptr(gen_depth, torch.float32), ptr(gen_w, torch.float32), ptr(gen_b, torch.float32),
                 ptr(dyn_x, torch.bfloat16), ptr(dyn_w, torch.bfloat16))
    check(load().dasr_conv_fwd(C.byref(d), C.byref(a), stream_ptr()))
    return out


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, kh: int = 3, kw: int = 3,
               db: Optional[torch.Tensor] = None, ksplit_div: int = 0) -> torch.Tensor:
    """dw[Cout][kh*kw*Cin] (fp32, zeroed by the caller) += weight gradient; dy / x NHWC bf16.
    db (optional, fp32 [Cout]) += the bias gradient (column sums of dy), fused into the same kernel.
    ksplit_div > 1: that many times fewer split-K CTAs (side-stream launches, see include/dasr.h)."""
    B, H, W, Cout = dy.shape
    Cin = x.shape[3]
    d = WgradDesc(B, H, W, Cout, Cin, kh, kw, int(ksplit_div))
    check(load().dasr_conv_wgrad(C.byref(d), ptr(dy, torch.bfloat16), ptr(x, torch.bfloat16), ptr(dw, torch.float32),
                                 ptr(db, torch.float32) if db is not None else None, stream_ptr()))
    return dw


def conv_stats_slots(B, H, W, Cin, Cout, ks=3) -> int:
    d = ConvDesc(B, H, W, Cin, Cout, ks, EPI_STATS, 0, 1, 0, 0, 0, 0.0, 0)
    n = load().dasr_conv_stats_slots(C.byref(d))
    if n <= 0:
        check(n)
    return n


def pack_weights(descs: Sequence[PackDesc], scratch: torch.Tensor) -> None:
    arr = (PackDesc * len(descs))(*descs)
    check(load().dasr_pack_weights(arr, len(descs), ptr(scratch, torch.float32), stream_ptr()))


def pack_desc(v, dst, *, g=None, alpha=None, alpha_mode=0, bias=None, bias2=None, dst_bias=None, mode=PACK_CONV,
              shuffle_r=0, row_offset=0, rows_per_tap=0, dst_stride=None) -> PackDesc:
    """dst_stride: element count of the WHOLE destination matrix when ``dst`` is a row slice of it (plane stride of
    the fp32-split mode; ignored otherwise)."""
    return PackDesc(ptr(v, torch.float32), ptr(g, torch.float32), ptr(alpha, torch.float32), ptr(bias, torch.float32),
                    ptr(bias2, torch.float32), ptr(dst, torch.bfloat16), ptr(dst_bias, torch.float32),
                    v.shape[0], v.shape[1], v.shape[2], mode, alpha_mode, shuffle_r, row_offset, rows_per_tap,
                    int(dst_stride if dst_stride is not None else dst.numel()))

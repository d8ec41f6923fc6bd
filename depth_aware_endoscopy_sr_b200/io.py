"""Input preparation and output conversion on the device.

    depth_masks(depth, num, fixed_range)   ``LQGTker_Depth_dataset.getDepthMask`` (codes/data/LQGTker_Depth_dataset.py
                                           :204-226) for a whole batch: returns the reference's one-hot fp32
                                           ``DepthMaskList [B,num,h,w]`` and the u8 label map the kernels consume
    tensor2img(sr, min_max)                ``utils.util.tensor2img`` (codes/utils/util.py:566-590) per frame:
                                           ``[B,3,H,W]`` fp32 RGB -> ``[B,H,W,3]`` uint8 BGR on the device, so only a
                                           quarter of the bytes crosses PCIe and no CPU pass is needed

    psnr(sr_u8, gt_u8, crop)               ``utils.util.calculate_psnr`` on the uint8 frames of ``tensor2img`` with the
                                           border crop of codes/train.py:251-257, per frame
    ssim(img1, img2)                       ``pytorch_ssim.ssim`` (codes/pytorch_ssim/__init__.py), per frame or averaged

All run in libdasr_b200.so; CPU tensors raise.
"""
from __future__ import annotations

import torch

from . import _lib as L


def depth_masks(depth: torch.Tensor, num: int = 10, fixed_range: bool = False, want_masks: bool = True):
    """depth [B,1,h,w] fp32 CUDA -> (masks fp32 [B,num,h,w] or None, labels u8 [B,h,w])."""
    depth = depth.contiguous().float()
    B, _, h, w = depth.shape
    labels = torch.empty(B, h, w, device=depth.device, dtype=torch.uint8)
    masks = torch.empty(B, num, h, w, device=depth.device, dtype=torch.float32) if want_masks else None
    L.check(L.load().dasr_depth_masks(L.ptr(depth), L.ptr(labels), L.ptr(masks), None, B, num, h, w,
                                      1 if fixed_range else 0, L.stream_ptr()))
    return masks, labels


def tensor2img(sr: torch.Tensor, min_max=(0.0, 1.0)) -> torch.Tensor:
    """sr [B,3,H,W] (or [3,H,W]) fp32 CUDA, RGB -> uint8 [B,H,W,3] (or [H,W,3]), BGR, still on the device."""
    squeeze = sr.dim() == 3
    x = (sr.unsqueeze(0) if squeeze else sr).contiguous().float()
    B, C, H, W = x.shape
    if C != 3:
        raise RuntimeError("tensor2img (B200) converts 3-channel frames")
    img = torch.empty(B, H, W, 3, device=x.device, dtype=torch.uint8)
    L.check(L.load().dasr_tensor2img(L.ptr(x), L.ptr(img), B, H, W, float(min_max[0]), float(min_max[1]),
                                     L.stream_ptr()))
    return img[0] if squeeze else img


def psnr(sr_u8: torch.Tensor, gt_u8: torch.Tensor, crop: int = 0) -> torch.Tensor:
    """uint8 frames [F,H,W,C] (or [H,W,C]) on the device -> PSNR per frame (float64 tensor on the device),
    20*log10(255/sqrt(mse)) over the region inside a ``crop``-pixel border; inf for identical frames."""
    a = (sr_u8.unsqueeze(0) if sr_u8.dim() == 3 else sr_u8).contiguous()
    b = (gt_u8.unsqueeze(0) if gt_u8.dim() == 3 else gt_u8).contiguous()
    if a.dtype != torch.uint8 or b.dtype != torch.uint8 or a.shape != b.shape:
        raise RuntimeError("psnr: two uint8 frame tensors of the same shape are needed")
    F_, H, W, C = a.shape
    acc = torch.empty(F_, device=a.device, dtype=torch.int64)
    out = torch.empty(F_, device=a.device, dtype=torch.float64)
    L.check(L.load().dasr_psnr_u8(L.ptr(a), L.ptr(b), L.ptr(acc), L.ptr(out), F_, H, W, C, int(crop), L.stream_ptr()))
    return out


def ssim(img1: torch.Tensor, img2: torch.Tensor, size_average: bool = True) -> torch.Tensor:
    """fp32 frames [F,C,H,W] in [0,1] on the device -> mean SSIM (scalar) or per-frame SSIM [F]."""
    a, b = img1.contiguous().float(), img2.contiguous().float()
    if a.shape != b.shape or a.dim() != 4:
        raise RuntimeError("ssim: two [F,C,H,W] tensors of the same shape are needed")
    F_, C, H, W = a.shape
    lib = L.load()
    part = torch.empty(F_ * C * lib.dasr_ssim_tiles(H, W), device=a.device, dtype=torch.float32)
    out = torch.empty(F_, device=a.device, dtype=torch.float32)
    L.check(lib.dasr_ssim(L.ptr(a), L.ptr(b), L.ptr(part), L.ptr(out), F_, C, H, W, L.stream_ptr()))
    if not size_average:
        return out
    if F_ == 1:
        return out[0]
    return out.mean()      # frames have equal sizes: the mean of the per-frame means is the global mean

"""Input preparation and output conversion on the device.

    depth_masks(depth, num, fixed_range)   ``LQGTker_Depth_dataset.getDepthMask`` (codes/data/LQGTker_Depth_dataset.py
                                           :204-226) for a whole batch: returns the reference's one-hot fp32
                                           ``DepthMaskList [B,num,h,w]`` and the u8 label map the kernels consume
    tensor2img(sr, min_max)                ``utils.util.tensor2img`` (codes/utils/util.py:566-590) per frame:
                                           ``[B,3,H,W]`` fp32 RGB -> ``[B,H,W,3]`` uint8 BGR on the device, so only a
                                           quarter of the bytes crosses PCIe and no CPU pass is needed

Both run in libdasr_b200.so; CPU tensors raise.
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

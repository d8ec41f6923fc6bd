"""Deterministic synthetic weights and inputs shared by tests, bench.py and smoke().

There is no network for datasets or checkpoints, so every benchmark / parity case is driven by
tensors generated here from a seed (CPU generator => identical on every box).

* ``depth_masks``  bins every depth map into K equal-width one-hot masks over its own range (the input convention
  of the reference data pipeline, codes/data/LQGTker_Depth_dataset.py:204-226; pinned to the reference function by
  tests/test_io_golden_cpu.py).
* ``synthetic_inputs`` follows SURVEY.md section 8(d): LQ in [0,1), depth in [0.01,10), GT in [0,1).
* ``fill_state_dict`` fills a DepthNet ``state_dict`` *layout* (names + shapes) with seeded values.  It is
  deliberately harsher than the default init: ``weight_g != ||weight_v||`` so that the weight-norm path
  matters, the SEAN blend scalars are drawn from U[0,1) like the reference (normalization.py:31-32), and the
  output convolution is centred on 0.5 so that the SR image fills [0,1] (with torch's default init the
  reference's output sits in [0, 0.14] and the final clamp would hide most errors).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch


def depth_masks(depth: torch.Tensor, num: int = 10) -> torch.Tensor:
    """depth [B,1,h,w] fp32 -> one-hot fp32 masks [B,num,h,w]: ``num`` equal-width bins over every image's own
    [min, max) -- the input convention of the reference's data pipeline.  A vectorised generator for synthetic inputs;
    the checker for the device kernel ``io.depth_masks`` is the golden vector recorded from the reference's own
    ``getDepthMask`` (tests/golden/io_golden.npz) and its restatement in oracle/depthnet_oracle.py."""
    d = depth[:, 0].float()
    lo = d.amin(dim=(1, 2), keepdim=True)
    step = (d.amax(dim=(1, 2), keepdim=True) - lo) / num
    k = torch.arange(num, dtype=torch.float32).view(1, num, 1, 1)
    start = lo.unsqueeze(1) + step.unsqueeze(1) * k
    end = lo.unsqueeze(1) + step.unsqueeze(1) * (k + 1)
    dd = d.unsqueeze(1)
    return ((dd >= start) & (dd < end)).float()


def synthetic_inputs(batch: int, h: int, w: int, scale: int = 8, num_masks: int = 10, seed: int = 0,
                     with_gt: bool = False):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    lq = torch.rand(batch, 3, h, w, generator=g)
    depth = 0.01 + 9.99 * torch.rand(batch, 1, h, w, generator=g)
    masks = depth_masks(depth, num_masks)
    if with_gt:
        gt = torch.rand(batch, 3, h * scale, w * scale, generator=g)
        return lq, depth, masks, gt
    return lq, depth, masks


def _seed_for(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF


def fill_state_dict(layout: "OrderedDict[str, torch.Size]", seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """layout: name -> shape (e.g. ``{k: v.shape for k, v in net.state_dict().items()}``)."""
    sd = OrderedDict()
    for name, shape in layout.items():
        g = torch.Generator(device="cpu")
        g.manual_seed(_seed_for(name, seed))
        shape = tuple(shape)
        leaf = name.rsplit(".", 1)[-1]
        if leaf in ("alpha_beta", "alpha_gamma"):
            t = torch.rand(shape, generator=g)
        elif leaf == "weight_g":
            # companion weight_v decides the scale; use U[0.6,1.4] * ||v|| computed below
            t = 0.6 + 0.8 * torch.rand(shape, generator=g)
        elif leaf == "bias":
            t = (torch.rand(shape, generator=g) - 0.5) * 0.2
        else:  # weight / weight_v: uniform, variance 0.64 / fan_in (activations neither vanish nor explode)
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 0.8 * math.sqrt(3.0 / max(fan_in, 1))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        if name == "conv_output.weight":
            t = t * 0.5
        if name == "conv_output.bias":
            t = t + 0.5  # centre the SR image inside [0,1] so that the clamp does not hide errors
        sd[name] = t.float()
    # weight_g := factor * ||weight_v|| (norm over all dims except 0, as torch weight_norm dim=0)
    for name in list(sd.keys()):
        if name.endswith(".weight_g"):
            v = sd[name[:-1] + "v"]
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(sd[name].shape)
            sd[name] = sd[name] * nrm
    return sd

"""One training step of the reference (``F_Model_depthCond.optimize_parameters``, codes/models/F_model_depthCond.py
:158-192, with ``pixel_criterion: l1`` + ``dynamic_loss`` as in options/train/*.yml) on the B200 kernels:

    zero_grad -> SR = netG(LQ, Depth, DepthMaskList) -> total = l_pix_w * L1 + dynamic depth-mask loss
    -> backward -> [flat gradient all-reduce] -> Adam step

``TrainStep`` keeps everything on the device: the step returns the loss vector as a device tensor (one D2H read for
the whole log instead of the reference's 23 ``.item()`` syncs per step).

``graph=True`` records the whole step (weight packing, ~770 kernels of forward / loss / backward, the NCCL gradient
all-reduce and Adam) ONCE into a CUDA graph and replays it per step: the step is launch-bound when issued from
Python (18 ms of host time against 17 ms of kernels at batch 16).  The first ``warmup`` calls run eagerly (they are
real training steps); inputs are copied into the graph's static buffers; the only host work per replay is the
upload of Adam's two step-dependent scalars.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import loss as _loss
from . import parallel as _par
from .optim import FusedAdam


class TrainStep:
    def __init__(self, net: nn.Module, num_masks: int = 10, lr: float = 1e-3, betas=(0.9, 0.99), weight_decay: float = 0.0,
                 l_pix_w: float = 1.0, dynamic_weight: float = 10.0, distributed: bool = False, mode: str = "ddp",
                 process_group=None, graph: bool = False, warmup: int = 2):
        self.net = net
        dev = next(net.parameters()).device
        self.dynamic_loss = _loss.dynamic_weight_mask_loss(dict(dynamic_criterion="smoothl1", dynamic_weight=dynamic_weight),
                                                           num_trainable_para=num_masks).to(dev)
        self.l_pix_w, self.l_dyn_w = float(l_pix_w), float(dynamic_weight)
        self.mode = mode
        self.group = process_group
        self.distributed = distributed and _par.world(process_group)[1] > 1
        self.model = _par.FlatDataParallel(net, process_group=process_group, mode=mode) if self.distributed else net
        if self.distributed:
            _par.broadcast_parameters_(self.dynamic_loss, 0, process_group)
        # optimiser over netG's trainable parameters + the dynamic-loss weights (F_model_depthCond.py:88-101)
        params = [p for p in net.parameters() if p.requires_grad] + list(self.dynamic_loss.parameters())
        self.optimizer = FusedAdam(params, lr=lr, betas=betas, weight_decay=weight_decay, capturable=graph)
        self.graph = bool(graph)
        if self.graph and int(warmup) < 1:
            # the first call re-homes the parameters into FusedAdam's flat buffers and allocates its device scalars:
            # that must not happen inside the capture
            raise ValueError("TrainStep(graph=True) needs warmup >= 1 eager step before the capture")
        self._warmup = int(warmup)
        self._calls = 0
        self._g = None
        self._static_in = None
        self._static_out = None
        self.launches_per_step = 0
        self._hook = _par.loss_sums_hook(process_group) if (self.distributed and mode == "global") else None
        self._nscale = float(_par.world(process_group)[1]) if self._hook is not None else 1.0

    def __call__(self, lq: torch.Tensor, depth: torch.Tensor, masks: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
        """Returns the device vector [total, l_pix, l_dyn, w_pix/n, loss_k.., softmax_k.., ...] (loss.O_* offsets)."""
        if not self.graph:
            return self._step(lq, depth, masks, gt)
        self._calls += 1
        if self._calls <= self._warmup:
            # eager steps on a side stream (torch's rule for allocations that a later capture will replay)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.optimizer.advance()
                out = self._step(lq, depth, masks, gt)
            torch.cuda.current_stream().wait_stream(s)
            return out
        if self._g is None:
            self._static_in = tuple(t.detach().clone().contiguous() for t in (lq, depth, masks, gt))
            eng = self.net.engine()
            eng.always_pack = True
            self._g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            # captured on a high-priority stream: the CTAs of the main chain are dispatched ahead of the side-stream
            # work (actv, weight gradients), which keeps the default priority (7.54 -> 7.48 ms; DASR_TRAIN_PRIO=0: off)
            cap = torch.cuda.Stream(priority=-1) if os.environ.get("DASR_TRAIN_PRIO", "1") == "1" else None
            with torch.cuda.graph(self._g, stream=cap):
                self._static_out = self._step(*self._static_in)
            self.launches_per_step = L.launch_count() - n0      # kernels of this library inside one replay
        for dst, src in zip(self._static_in, (lq, depth, masks, gt)):
            if dst.shape != src.shape:
                raise RuntimeError("TrainStep(graph=True) was captured for input shape %s, got %s" % (
                    tuple(dst.shape), tuple(src.shape)))
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.optimizer.advance()
        self._g.replay()
        L.note_replayed_launches(self.launches_per_step)
        return self._static_out

    def _step(self, lq, depth, masks, gt):
        self.optimizer.zero_grad(set_to_none=True)
        sr = self.model(lq, depth, masks)
        vec = _loss.loss_vector(sr, gt, masks, self.dynamic_loss.trainable_weight, w_pix=self.l_pix_w,
                                w_dyn=self.l_dyn_w, sums_hook=self._hook, n_scale=self._nscale)
        vec[_loss.O_TOTAL].backward()
        if self.distributed and self.mode == "ddp":
            # per-rank losses: average the gradients of the 10 loss weights like every other gradient.  In "global"
            # mode the loss partial sums were all-reduced BEFORE dasr_loss_finalize, so every rank already holds the
            # gradient of the global-batch loss w.r.t. the loss weights (identical on all ranks): nothing to reduce.
            _par.sync_extra_grads_(self.dynamic_loss.parameters(), self.group, average=True)
        self.optimizer.step()
        return vec.detach()

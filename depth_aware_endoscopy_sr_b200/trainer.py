"""One training step of the reference (``F_Model_depthCond.optimize_parameters``, codes/models/F_model_depthCond.py
:158-192, with ``pixel_criterion: l1`` + ``dynamic_loss`` as in options/train/*.yml) on the B200 kernels:

    zero_grad -> SR = netG(LQ, Depth, DepthMaskList) -> total = l_pix_w * L1 + dynamic depth-mask loss
    -> backward -> [flat gradient all-reduce] -> Adam step

``TrainStep`` keeps everything on the device: the step returns the loss vector as a device tensor (one D2H read for
the whole log instead of the reference's 23 ``.item()`` syncs per step).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import loss as _loss
from . import parallel as _par
from .optim import FusedAdam


class TrainStep:
    def __init__(self, net: nn.Module, num_masks: int = 10, lr: float = 1e-3, betas=(0.9, 0.99), weight_decay: float = 0.0,
                 l_pix_w: float = 1.0, dynamic_weight: float = 10.0, distributed: bool = False, mode: str = "ddp",
                 process_group=None):
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
        self.optimizer = FusedAdam(params, lr=lr, betas=betas, weight_decay=weight_decay)
        self._hook = _par.loss_sums_hook(process_group) if (self.distributed and mode == "global") else None
        self._nscale = float(_par.world(process_group)[1]) if self._hook is not None else 1.0

    def __call__(self, lq: torch.Tensor, depth: torch.Tensor, masks: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
        """Returns the device vector [total, l_pix, l_dyn, w_pix/n, loss_k.., softmax_k.., ...] (loss.O_* offsets)."""
        self.optimizer.zero_grad(set_to_none=True)
        sr = self.model(lq, depth, masks)
        vec = _loss.loss_vector(sr, gt, masks, self.dynamic_loss.trainable_weight, w_pix=self.l_pix_w,
                                w_dyn=self.l_dyn_w, sums_hook=self._hook, n_scale=self._nscale)
        vec[_loss.O_TOTAL].backward()
        if self.distributed:
            _par.sync_extra_grads_(self.dynamic_loss.parameters(), self.group, average=(self.mode == "ddp"))
        self.optimizer.step()
        return vec.detach()

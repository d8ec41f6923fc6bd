"""Training criteria of the reference on the K-LOSS kernels (C ABI: dasr_loss_fwd / _finalize / _bwd).

Mirrors, with the same names, constructor arguments and return values:

    L1Loss                      nn.L1Loss as ``F_Model_depthCond.cri_pix``      (codes/models/F_model_depthCond.py:52,164)
    dynamic_weight_mask_loss    codes/models/modules/mask_loss.py:49-90 (``loss_type: smoothl1``)

and adds ``training_loss`` -- both criteria in ONE pass over SR/HR (what ``optimize_parameters`` computes at
F_model_depthCond.py:163-190 with ``pixel_criterion: l1`` + ``dynamic_loss``).  There is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L

_KM = L.LOSS_KMAX
# layout of the vector dasr_loss_finalize writes
O_TOTAL, O_PIX, O_DYN, O_CPIX, O_LOSSK, O_SOFTMAX, O_COEF, O_DW = 0, 1, 2, 3, 4, 4 + _KM, 4 + 2 * _KM, 4 + 3 * _KM


def mask_labels(masks: torch.Tensor):
    """fp32 one-hot masks [B,K,h,w] -> (u8 label map [B,h,w], device flag that is non-zero if they are not one-hot)."""
    B, K, h, w = masks.shape
    labels = torch.empty(B, h, w, device=masks.device, dtype=torch.uint8)
    flag = torch.zeros(1, device=masks.device, dtype=torch.int32)
    L.check(L.load().dasr_mask_labels(L.ptr(masks, torch.float32), L.ptr(labels), L.ptr(flag), B, K, h, w, L.stream_ptr()))
    return labels, flag


class _LossFn(torch.autograd.Function):
    """(sr, hr, masks, wdyn) -> the finalize vector; differentiable w.r.t. sr and wdyn through entries 0..2."""

    @staticmethod
    def forward(ctx, sr, hr, masks, wdyn, use_pix, use_dyn, w_pix, w_dyn, sums_hook, n_scale):
        if not sr.is_cuda:
            raise RuntimeError("the B200 loss kernels need CUDA tensors; there is no CPU fallback")
        lib = L.load()
        sr_c = sr.detach().contiguous().float()
        hr_c = hr.detach().contiguous().float()
        masks_c = masks.detach().contiguous().float()
        B, C, Ho, Wo = sr_c.shape
        K, h, w = masks_c.shape[1], masks_c.shape[2], masks_c.shape[3]
        if hr_c.shape != sr_c.shape or masks_c.shape[0] != B:
            raise RuntimeError("loss: SR %s, HR %s and masks %s disagree" % (tuple(sr.shape), tuple(hr.shape),
                                                                             tuple(masks.shape)))
        s = L.stream_ptr()
        labels, flag = mask_labels(masks_c)
        rows = lib.dasr_loss_rows(B, Ho, Wo)
        if rows <= 0:
            L.check(rows)
        part = torch.empty(rows, L.LOSS_ROW, device=sr.device, dtype=torch.float32)
        sums = torch.empty(L.LOSS_ROW, device=sr.device, dtype=torch.float32)
        L.check(lib.dasr_loss_fwd(L.ptr(sr_c), L.ptr(hr_c), L.ptr(labels), L.ptr(masks_c), L.ptr(flag), L.ptr(part),
                                  L.ptr(sums), B, C, K, h, w, Ho, Wo, s))
        if sums_hook is not None:      # data parallel, "global batch" semantics: all-reduce the 33 partial sums
            sums_hook(sums)
        out = torch.empty(4 + 4 * _KM, device=sr.device, dtype=torch.float32)
        wd = None if wdyn is None else wdyn.detach().contiguous().float()
        L.check(lib.dasr_loss_finalize(L.ptr(sums), L.ptr(wd), L.ptr(out), K, C, float(sr_c.numel()) * n_scale,
                                       float(w_pix) * (1.0 if use_pix else 0.0),
                                       float(w_dyn) * (1.0 if use_dyn else 0.0), s))
        ctx.save_for_backward(sr_c, hr_c, masks_c, labels, flag, out)
        ctx.flags = (float(bool(use_pix)), float(bool(use_dyn)), wdyn is not None)
        ctx.mark_non_differentiable(labels)
        return out

    @staticmethod
    def backward(ctx, dout):
        sr, hr, masks, labels, flag, out = ctx.saved_tensors
        use_pix, use_dyn, has_w = ctx.flags
        B, C, Ho, Wo = sr.shape
        K, h, w = masks.shape[1], masks.shape[2], masks.shape[3]
        dsr = torch.empty_like(sr)
        dw = torch.empty(K, device=sr.device, dtype=torch.float32) if has_w else None
        g = dout.contiguous().float()
        L.check(L.load().dasr_loss_bwd(L.ptr(sr), L.ptr(hr), L.ptr(labels), L.ptr(masks), L.ptr(flag), L.ptr(out),
                                       L.ptr(g), use_pix, use_dyn, L.ptr(dsr), L.ptr(dw), B, C, K, h, w, Ho, Wo,
                                       L.stream_ptr()))
        return dsr, None, None, dw, None, None, None, None, None, None


def loss_vector(sr, hr, masks, wdyn=None, *, use_pix=True, use_dyn=True, w_pix=1.0, w_dyn=10.0, sums_hook=None,
                n_scale=1.0) -> torch.Tensor:
    """The raw finalize vector (see include/dasr.h, dasr_loss_finalize)."""
    return _LossFn.apply(sr, hr, masks, wdyn, use_pix, use_dyn, w_pix, w_dyn, sums_hook, n_scale)


def training_loss(sr, hr, masks, wdyn, l_pix_w: float = 1.0, l_dyn_w: float = 10.0, *, sums_hook=None, n_scale=1.0):
    """total = l_pix_w * L1(sr, hr) + l_dyn_w * sum_k softmax(wdyn)_k loss_k  (F_model_depthCond.py:163-190).
    Returns (total, l_pix, l_dyn, loss_k [K], softmax(wdyn) [K]); only ``total`` needs to be back-propagated."""
    K = masks.shape[1]
    v = loss_vector(sr, hr, masks, wdyn, w_pix=l_pix_w, w_dyn=l_dyn_w, sums_hook=sums_hook, n_scale=n_scale)
    return v[O_TOTAL], v[O_PIX], v[O_DYN], v[O_LOSSK:O_LOSSK + K].detach(), v[O_SOFTMAX:O_SOFTMAX + K].detach()


class L1Loss(nn.Module):
    """Drop-in for ``nn.L1Loss()`` as the reference's ``cri_pix`` (mean reduction, NCHW fp32 CUDA images).
    The masks argument of the kernel is not needed here: a single all-ones mask is used."""

    def forward(self, sr: torch.Tensor, hr: torch.Tensor) -> torch.Tensor:
        B, _, Ho, Wo = sr.shape
        ones = torch.ones(B, 1, 1, 1, device=sr.device, dtype=torch.float32)
        return loss_vector(sr, hr, ones, None, use_pix=True, use_dyn=False, w_pix=1.0, w_dyn=0.0)[O_PIX]


class dynamic_weight_mask_loss(nn.Module):
    """Same constructor and return tuple as the reference class (codes/models/modules/mask_loss.py:49-90):
    ``opt`` = the yml ``train.dynamic_loss`` block (``dynamic_criterion`` must be ``smoothl1``, ``dynamic_weight``
    the scale), ``num_trainable_para`` = depthMaskNum.  forward -> (loss_list, weighted_loss_list, weighted_loss,
    softmax_weight)."""

    def __init__(self, opt, device=None, num_trainable_para: int = 10):
        # (``device`` is accepted and ignored exactly like the reference's constructor, mask_loss.py:50: the module is
        # moved with .to(device) by its owner)
        super().__init__()
        loss_type = opt["dynamic_criterion"]
        if loss_type != "smoothl1":
            raise NotImplementedError("Loss type [{:s}] for depth loss is not recognized.".format(str(loss_type)))
        if num_trainable_para > _KM:
            raise NotImplementedError("at most %d depth masks" % _KM)
        self.loss_type_mask = loss_type
        self.l_mask_w = opt["dynamic_weight"]
        self.num_trainable_parameters = num_trainable_para
        self.trainable_weight = nn.Parameter(torch.ones(num_trainable_para))

    def forward(self, sr_img, hr_img, depthMaskList):
        ch = depthMaskList.shape[1]
        assert self.num_trainable_parameters == ch, "The number of trainable parameters for dynamic loss is not enought."
        v = loss_vector(sr_img, hr_img, depthMaskList, self.trainable_weight, use_pix=False, use_dyn=True, w_pix=0.0,
                        w_dyn=float(self.l_mask_w))
        lk = v[O_LOSSK:O_LOSSK + ch].detach()
        sw = v[O_SOFTMAX:O_SOFTMAX + ch].detach()
        loss_list = [lk[i] for i in range(ch)]
        weighted = [sw[i] * lk[i] for i in range(ch)]
        return loss_list, weighted, v[O_DYN], sw

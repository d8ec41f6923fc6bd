"""Training path of DepthNet on libdasr_b200.so: forward with saved activations + hand-written backward.

``depthnet_apply`` is what ``DepthNet.forward`` calls when gradients are enabled.  It is ONE
``torch.autograd.Function`` over (LQ, depth, masks, *parameters): the forward runs the same kernel schedule as
``Engine.infer`` while recording a tape of backward closures; the backward replays the tape in reverse and
returns one gradient per parameter (``None`` for parameters the reference never touches either, e.g.
``depth-residual14.*`` -- SURVEY.md headline fact 5).  It restates autograd of the reference graph
(``total_loss.backward()``, codes/models/F_model_depthCond.py:191) with these kernels:

    data gradients     dasr_conv_fwd over DASR_PACK_DGRAD weights (flipped / transposed), with the ReLU /
                       LeakyReLU mask and the residual accumulation fused into the epilogue
    weight gradients   dasr_conv_wgrad (tcgen05, MN-major operands) into one flat fp32 buffer in the packed
                       layout, then dasr_unpack_grads (weight-norm backward, SEAN alpha blend, index maps)
    SEAN / IN          dasr_sean_bwd1 / _bwd2;   K-DYN  dasr_dynconv_bwd + dasr_table_bwd +
                       dasr_style_mix_bwd;   pooling  dasr_region_pool_bwd;   mlp_mask  dasr_actv_bwd
    tail               dasr_unshuffle_actgrad (PixelShuffle + LeakyReLU), dasr_out9_bwd_prep (clamp + im2row)

All parameter gradients of one backward live in ONE flat fp32 buffer (``Engine.last_flat_grad``), which is what
the data-parallel wrapper all-reduces (parallel.py).
"""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib as L

BF16 = torch.bfloat16


class _T:
    """Activation on the tape: NHWC bf16 tensor + how its producer's activation is undone in the backward."""
    __slots__ = ("data", "act", "pending", "grad", "masked", "shuffle_r", "unshuffled")

    def __init__(self, data, act="none"):
        self.data = data
        self.act = act          # 'none' | 'relu' | 'lrelu' : mask applied lazily by the producer's backward
        self.pending = 0        # consumers that have not back-propagated yet
        self.grad = None
        self.masked = False     # the accumulated grad already includes the activation mask
        self.shuffle_r = 0      # 2: output of a PixelShuffle(2) + LeakyReLU convolution (EPI_SHUFFLE2)
        self.unshuffled = False  # grad was written space-to-depth, masked, by its (single) producer's epilogue

    @property
    def slope(self):
        return 0.0 if self.act == "relu" else 0.2


def _dbg_mask(eng, key, t):
    """tests / probes (Engine.debug): the activation pattern of the NHWC act tensor ``t`` as an NCHW bool tensor --
    which units of this ReLU / LeakyReLU are on.  The backward kernels switch on exactly this (stored value > 0)."""
    if eng.debug is not None:
        eng.debug["mask:" + key] = (L.act_value(t) > 0).permute(0, 3, 1, 2).contiguous()


class Tape:
    def __init__(self, eng):
        self.eng = eng
        self.lib = L.load()
        self.ops = []
        self.s = L.stream_ptr()
        # weight gradients are leaves of the backward: with ``Engine.wgrad_overlap`` they are issued on a side stream
        # and fill the wave tails / launch gaps of the data-gradient chain (begin_backward / join_backward)
        self.wg_side = None
        self.wg_main = None
        self._held = []
        self._wg_next = 0

    # ------------------------------------------------------------------ gradient bookkeeping
    def use(self, t: _T) -> _T:
        t.pending += 1
        return t

    def accum(self, t: _T, g: torch.Tensor):
        """Add a contribution (already a finished tensor) to t.grad."""
        t.pending -= 1
        if t.grad is None:
            t.grad = g
        else:
            out = L.act_like(g)
            L.check(self.lib.dasr_add(L.ptr(t.grad), None, L.ptr(g), L.ptr(out), g.numel(), self.s))
            t.grad = out

    def take(self, t: _T) -> torch.Tensor:
        """Final gradient w.r.t. the PRE-activation value of t (applies the lazy activation mask)."""
        g = t.grad
        if g is None:
            raise RuntimeError("tape: tensor without gradient")
        t.grad = None
        if t.act in ("relu", "lrelu") and not t.masked:
            out = L.act_like(g)
            L.check(self.lib.dasr_actgrad(L.ptr(g), L.ptr(t.data), L.ptr(out), g.numel(), t.slope, self.s))
            g = out
        return g

    def dgrad_into(self, t: _T, dy: torch.Tensor, wname: str, *, ks=3, kw=0, subsample=1):
        """t.grad (+)= conv(dy, dgrad-packed weights).  Fuses the accumulation and -- when this is the last
        contribution -- the activation mask of t into the convolution epilogue."""
        eng = self.eng
        pk = eng._packed[wname]
        t.pending -= 1
        last = t.pending == 0
        B, H, W, _ = dy.shape
        if subsample == 2:
            out = L.act_empty(B, (H + 1) // 2, (W + 1) // 2, pk.cout, device=dy.device)
        else:
            out = L.act_empty(B, H, W, pk.cout, device=dy.device)
        mask = None
        slope = 0.0
        if last and t.act in ("relu", "lrelu") and not t.masked:
            mask, slope = t.data, t.slope
            t.masked = True
        unshuffle = 0
        if (last and t.shuffle_r == 2 and t.grad is None and subsample == 1 and H % 2 == 0 and W % 2 == 0):
            # the only gradient contribution of a PixelShuffle(2) + LeakyReLU output: write it space-to-depth with
            # the LeakyReLU mask applied -- exactly the tensor the shuffle convolution's backward consumes
            # (replaces the dasr_unshuffle_actgrad pass: 3 x 80 us per training step at B = 16)
            out = L.act_empty(B, H // 2, W // 2, 4 * pk.cout, device=dy.device)
            mask, slope, unshuffle = t.data, 0.2, 2
            t.unshuffled = True
        L.conv_fwd(dy, pk.w, eng._zero_bias, out, Cout=pk.cout, ks=ks, kw=kw, subsample=subsample, resid=t.grad,
                   actmask=mask, mask_slope=slope, unshuffle=unshuffle)
        t.grad = out

    # ------------------------------------------------------------------ weight-gradient helpers
    def wgrad(self, dy: torch.Tensor, x: torch.Tensor, wname: str, kh=3, kw=3, bias=False):
        """Weight gradient (and, with ``bias``, the bias gradient in the same kernel)."""
        dw = self.eng._dw_view(wname)
        db = self.eng._db_view(wname) if bias else None
        if self.wg_side is None:
            L.conv_wgrad(dy, x, dw, kh, kw, db=db)
            return
        side = self.wg_side[self._wg_next % len(self.wg_side)]
        self._wg_next += 1
        ev = torch.cuda.Event()
        ev.record(self.wg_main)             # dy (and the zeroed flat buffers) are ready at this point of the main stream
        side.wait_event(ev)
        with torch.cuda.stream(side):
            L.conv_wgrad(dy, x, dw, kh, kw, db=db, ksplit_div=self.eng.wgrad_ksplit_div)
        # the caching allocator orders reuse on the main stream only: keep both operands alive until the join
        self._held.append((dy, x))

    @contextlib.contextmanager
    def leaf(self, *tensors):
        """Run a leaf chain of the backward (kernels whose results only reach parameter gradients) on the next side
        stream; yields the raw stream to launch on.  ``tensors``: every operand allocated on the main stream."""
        if self.wg_side is None or not self.eng.leaf_overlap:
            yield self.s
            return
        side = self.wg_side[self._wg_next % len(self.wg_side)]
        self._wg_next += 1
        ev = torch.cuda.Event()
        ev.record(self.wg_main)
        side.wait_event(ev)
        self._held.append(tensors)
        with torch.cuda.stream(side):
            yield side.cuda_stream

    def sync_leaves(self):
        """Main stream waits for everything issued on the side streams so far (before a consumer on the main stream)."""
        if self.wg_side is not None:
            for side in self.wg_side[:max(1, min(self._wg_next, len(self.wg_side)))]:
                ev = torch.cuda.Event()
                ev.record(side)
                self.wg_main.wait_event(ev)

    def begin_backward(self, device):
        self.s = L.stream_ptr()
        # kernel-by-kernel issue is host-bound (the cross-stream events only add host time: 13.7 -> 17 ms per step), so
        # the side streams are used inside a CUDA-graph capture (TrainStep(graph=True)) or when forced
        eng = self.eng
        if eng.wgrad_overlap and (eng.overlap_eager or torch.cuda.is_current_stream_capturing()):
            self.wg_main = torch.cuda.current_stream(device)
            self.wg_side = self.eng._wgrad_streams(device)
            self._wg_next = 0

    def join_backward(self):
        """Main stream waits for the weight gradients of the side stream (before unpack / all-reduce)."""
        if self.wg_side is not None:
            self.sync_leaves()
            self._held.clear()
            self.wg_side = None

    def bias_grad(self, dy: torch.Tensor, wname: str):
        db = self.eng._db_view(wname)
        C_ = dy.shape[-1]
        L.check(self.lib.dasr_colsum(L.ptr(dy), L.ptr(db), dy.numel() // C_, C_, self.s))


def _conv_train(tp: Tape, x: _T, name: str, *, act="none", subsample=1, shuffle=False, convt_src: Optional[_T] = None,
                need_dgrad=True, bias_grad=True) -> _T:
    """Generic convolution of the trunk / tail / encoder with its backward closure.
    x is the tensor the kernel reads (for the transposed conv: the zero-stuffed tensor; convt_src is then the
    tensor that receives the data gradient, through the subsampling epilogue)."""
    eng = tp.eng
    lib = tp.lib
    acts = {"none": L.ACT_NONE, "relu": L.ACT_RELU, "lrelu": L.ACT_LRELU}
    r = 2 if shuffle is True else int(shuffle)        # PixelShuffle factor (0: none)
    if r == 3:
        # x3 tail: plain conv into the shuffled channel order (+ activation), then PixelShuffle(3) as a copy
        u = eng._conv(x.data, name, act=acts[act])
        B_, H_, W_, C_ = u.shape
        out = L.act_empty(B_, 3 * H_, 3 * W_, C_ // 9, device=u.device)
        L.check(lib.dasr_pixel_shuffle(L.ptr(u), L.ptr(out), B_, H_, W_, C_ // 9, 3, tp.s))
        del u
        o = _T(out, "handled")
    elif shuffle:
        out = eng._conv(x.data, name, epi=L.EPI_SHUFFLE2, act=acts[act])
        o = _T(out, "handled")
        o.shuffle_r = 2
    else:
        out = eng._conv(x.data, name, act=acts[act], subsample=subsample)
        o = _T(out, act)
    if act != "none":
        _dbg_mask(eng, name, o.data)
    target = convt_src if convt_src is not None else x
    if need_dgrad:
        tp.use(target)

    def backward():
        s = tp.s
        g = o.grad
        if shuffle and o.unshuffled:
            o.grad = None
            dy = g                 # already space-to-depth and masked (Tape.dgrad_into)
        elif shuffle:
            o.grad = None
            B, H2, W2, Cq = o.data.shape
            dy = L.act_empty(B, H2 // r, W2 // r, r * r * Cq, device=g.device)
            L.check(lib.dasr_unshuffle_actgrad(L.ptr(g), L.ptr(o.data), L.ptr(dy), B, H2 // r, W2 // r, Cq, 0.2, r, s))
        else:
            dy = tp.take(o)
        if subsample == 2:      # gradient on the stride-1 grid of the forward kernel
            B, Ho, Wo, Co = dy.shape
            H, W = x.data.shape[1], x.data.shape[2]
            full = L.act_empty(B, H, W, Co, device=dy.device)
            L.check(lib.dasr_zero_insert2_to(L.ptr(dy), L.ptr(full), B, Ho, Wo, Co, H, W, s))
            dy = full
        tp.wgrad(dy, x.data, name, bias=bias_grad)
        if need_dgrad:
            if convt_src is not None:
                tp.dgrad_into(convt_src, dy, name + ".dg", subsample=2)
            else:
                tp.dgrad_into(x, dy, name + ".dg")

    tp.ops.append(backward)
    return o


def _sean_train(tp: Tape, n: str, sean, cur: _T, conv_name: str, ctx, tables, dT_all, scr_all, *, first: bool,
                resid: Optional[_T], actv_pre=None) -> _T:
    """conv -> IN -> IN -> SEAN modulate -> ReLU (first) | + resid -> ReLU (second), with the backward closure.
    ``ctx``: Engine.mask_context of this feature-map resolution (depth map, masks, labels, flag, aux, mask image)."""
    eng, lib, s = tp.eng, tp.lib, tp.s
    x = cur.data
    B, H, W, nf = x.shape
    nf2 = 2 * nf
    K, lat = sean.label_nc, sean.len_latent
    dev = x.device
    depth, masks, flag, aux, mask16 = ctx["depth"], ctx["masks"], ctx["flag"], ctx["aux"], ctx["mask16"]
    # ---- forward (same kernels as Engine._dgb, plus the tensors the backward needs)
    grp = eng._sean_group[n]
    sidx = grp.index[n]
    # all instances of a width group were computed in two launches; K-DYN runs inside the SEAN GEMM as a K extension
    _stp, table, wdyn = eng.sean_tables(tables, n)
    nslots = L.conv_stats_slots(B, H, W, nf, nf)
    stats = torch.empty(B, nslots, nf, 2, device=dev, dtype=torch.float32)
    norm = torch.empty(B, nf, 2, device=dev, dtype=torch.float32)
    normk = torch.empty(B, nf, device=dev, dtype=torch.float32)
    y = eng._conv(x, conv_name, epi=L.EPI_STATS, stats=stats)      # coefficients are finalised inside the SEAN conv
    if actv_pre is not None:
        actv = actv_pre          # produced on the side stream (the caller made this stream wait for it)
    else:
        actv = L.act_empty(B, H, W, nf2, device=dev)
        L.check(lib.dasr_actv_fwd(L.ptr(depth), L.ptr(sean.mlp_mask[0].weight), L.ptr(sean.mlp_mask[0].bias),
                                  L.ptr(actv), B, H, W, nf2, 0, s))
    gamma = L.act_empty(B, H, W, nf, device=dev)
    if first:
        out = eng._conv(actv, n + ".gb_o", epi=L.EPI_SEAN, inner_relu=1, y=y, stats=stats, norm_out=norm, normk_out=normk,
                        dyn_x=mask16, dyn_w=wdyn, gamma_out=gamma)
    else:
        out = eng._conv(actv, n + ".gb_o", epi=L.EPI_SEAN, act=L.ACT_RELU, y=y, stats=stats, norm_out=norm, normk_out=normk,
                        dyn_x=mask16, dyn_w=wdyn, resid=resid.data, gamma_out=gamma)
    del table, stats
    o = _T(out, "handled")
    _dbg_mask(eng, n + ".actv", actv)
    _dbg_mask(eng, n + ".out", out)
    tp.use(cur)
    if resid is not None:
        tp.use(resid)

    def backward():
        HW = H * W
        dout = o.grad
        o.grad = None
        if eng.debug is not None:      # tests / probes: gradient w.r.t. this SEAN's output and its forward value
            eng.debug[n + ".dout"] = L.act_value(dout)
            eng.debug[n + ".out"] = L.act_value(out)
        slots = lib.dasr_sean_bwd_slots(HW)
        dgb = L.act_empty(B, H, W, nf2, device=dev)
        dn = L.act_empty(B, H, W, nf, device=dev)
        dskip = L.act_empty(B, H, W, nf, device=dev) if resid is not None else None
        part = torch.empty(B, slots, nf, 4, device=dev, dtype=torch.float32)
        L.check(lib.dasr_sean_bwd1(L.ptr(dout), L.ptr(out), L.ptr(y), L.ptr(norm), L.ptr(gamma), L.ptr(dgb), L.ptr(dn),
                                   L.ptr(dskip), L.ptr(part), B, HW, nf, s))
        if resid is not None:
            tp.accum(resid, dskip)
        dy = L.act_empty(B, H, W, nf, device=dev)
        # pass 2 also adds the [gamma_o; beta_o] bias gradient (column sums of dgb, accumulated by pass 1)
        L.check(lib.dasr_sean_bwd2(L.ptr(dn), L.ptr(y), L.ptr(norm), L.ptr(normk), L.ptr(part), L.ptr(dy),
                                   L.ptr(eng._db_view(n + ".gb_o")), B, HW, nf, s))
        # ---- gamma_o / beta_o convolution and mlp_mask
        tp.wgrad(dgb, actv, n + ".gb_o")
        pkd = eng._packed[n + ".gb_o.dg"]
        dA = L.act_empty(B, H, W, nf2, device=dev)
        gw, gb = eng._grad_view(n + ".mlp_mask.0.weight"), eng._grad_view(n + ".mlp_mask.0.bias")
        scr = scr_all[grp][sidx]     # zeroed once per step for all instances
        dT = dT_all[grp][sidx]
        # everything below only reaches parameter gradients (mlp_mask, the style tables): a leaf chain
        with tp.leaf(dgb, actv, dA) as s2:
            L.conv_fwd(dgb, pkd.w, eng._zero_bias, dA, Cout=nf2, ks=3, actmask=actv, mask_slope=0.0)
            L.check(lib.dasr_actv_bwd_tc(L.ptr(dA), L.ptr(aux), L.ptr(scr), L.ptr(gw), L.ptr(gb), B, H, W, nf2, s2))
        # ---- style branch: K-DYN backward into this instance's slice of dT_all; the table GEMM / A_i_j backward of
        # all instances run batched once every block has back-propagated (bwd_pool)
        # one-hot masks: tensor-core kernel; otherwise (device flag) the exact general-mask kernel -- each is a
        # no-op in the other's case, so no host synchronisation is needed to choose
        with tp.leaf(dgb) as s2:
            L.check(lib.dasr_dynconv_bwd_tc(L.ptr(dgb), L.ptr(aux), L.ptr(flag), L.ptr(dT), B, K, H, W, nf2, s2))
            L.check(lib.dasr_dynconv_bwd(L.ptr(dgb), None, L.ptr(masks), L.ptr(flag), L.ptr(dT), B, K, H, W, nf2, s2))
        # ---- the block convolution in front of the norm (its bias gradient is exactly zero: IN removes the mean)
        tp.wgrad(dy, x, conv_name)
        tp.dgrad_into(cur, dy, conv_name + ".dg")

    tp.ops.append(backward)
    return o


def _forward_train(eng, lq, depth, masks):
    net = eng.net
    lib = L.load()
    tp = Tape(eng)
    s = tp.s
    B, _, h, w = lq.shape
    K = masks.shape[1]
    dev = lq.device
    enc = net.encoder

    # ---- encoder
    f0d = L.act_empty(B, h, w, 32, device=dev)
    L.check(lib.dasr_conv_first(L.ptr(lq), L.ptr(enc.layer1.weight_v), L.ptr(enc.layer1.weight_g), L.ptr(enc.layer1.bias),
                                L.ptr(f0d), B, h, w, s))
    f0 = _T(f0d, "lrelu")
    _dbg_mask(eng, "encoder.layer1", f0d)

    def bwd_first():
        dy = tp.take(f0)
        lq32 = L.act_empty(B, h, w, 32, device=dev)
        L.check(lib.dasr_nchw3_to_nhwc32(L.ptr(lq), L.ptr(lq32), B, h, w, s))
        tp.wgrad(dy, lq32, "encoder.layer1", bias=True)

    tp.ops.append(bwd_first)

    vec = dvec = tables = dT_all = scr_all = None
    ctxs = {}        # Engine.mask_context per feature-map resolution of the depth-guided blocks
    if not net.isBaseline:
        e2 = _conv_train(tp, f0, "encoder.layer2", act="lrelu", subsample=2)
        e3 = _conv_train(tp, e2, "encoder.layer3", act="lrelu", subsample=2)
        h3, w3 = e3.data.shape[1], e3.data.shape[2]
        zd = L.act_empty(B, 2 * h3 - 1, 2 * w3 - 1, 128, device=dev)
        L.check(lib.dasr_zero_insert2(L.ptr(e3.data), L.ptr(zd), B, h3, w3, 128, s))
        z = _T(zd, "none")
        e4 = _conv_train(tp, z, "encoder.layer4", act="lrelu", convt_src=e3)
        e5 = _conv_train(tp, e4, "encoder.layer5", subsample=2)
        lat = e5.data.shape[3]
        P = e5.data.shape[1] * e5.data.shape[2]
        vec = torch.empty(B, K, lat, device=dev, dtype=torch.float32)
        msel = torch.empty(B, K, P, device=dev, dtype=torch.float32)
        cnt = torch.empty(B, K, device=dev, dtype=torch.float32)
        L.check(lib.dasr_region_pool_fwd(L.ptr(e5.data), L.ptr(masks), L.ptr(vec), L.ptr(msel), L.ptr(cnt), B,
                                         e5.data.shape[1], e5.data.shape[2], lat, K, h, w, s))
        dvec = torch.zeros(B, K, lat, device=dev, dtype=torch.float32)
        tp.use(e5)

        def bwd_pool():
            # style branch of ALL SEAN instances: dWs = dT^T stp, dstp = dT Ws, then the A_i_j backward (batched per
            # width group)
            tp.sync_leaves()             # dT_all is complete once the K-DYN leaf chains have run
            for grp in eng._sean_groups:
                nS = len(grp.names)
                N = grp.ws_rows
                dWs_all = eng._dw_flat[grp.wg_tables_off:grp.wg_tables_off + nS * N * lat]
                dstp_all = torch.empty(nS, B * K, lat, device=dev, dtype=torch.float32)
                # the weight gradient dWs only reaches parameter gradients: a leaf (side stream); the data gradient dstp
                # feeds the A_i_j / encoder chain on the main stream
                with tp.leaf(dT_all[grp], tables[grp][0]) as s2:
                    L.check(lib.dasr_table_bwd_parts(L.ptr(dT_all[grp]), L.ptr(tables[grp][0]), None, L.ptr(dWs_all), None,
                                                     nS, B * K, N, lat, 1, s2))
                L.check(lib.dasr_table_bwd_parts(L.ptr(dT_all[grp]), None, L.ptr(grp.ws_all), None, L.ptr(dstp_all), nS,
                                                 B * K, N, lat, 2, tp.s))
                L.check(lib.dasr_style_mix_bwd_batched(L.ptr(dstp_all), L.ptr(vec), L.ptr(grp.A_ptrs), L.ptr(grp.dA_ptrs),
                                                       L.ptr(grp.da_ptrs), L.ptr(dvec), nS, B, K, lat, tp.s))
            de5 = L.act_like(e5.data)
            L.check(lib.dasr_region_pool_bwd(L.ptr(dvec), L.ptr(msel), L.ptr(cnt), L.ptr(de5), B, P, lat, K, s))
            tp.accum(e5, de5)

        tp.ops.append(bwd_pool)
        # (inside the captured step the style-table chain runs on its own side stream beside the head convolutions)
        tside = None
        if eng.tables_overlap and torch.cuda.is_current_stream_capturing():
            tside = eng._side_streams.get(("tables", dev.index))
            if tside is None:
                tside = eng._side_streams[("tables", dev.index)] = torch.cuda.Stream(device=dev)
        tables = eng.style_tables(vec, side=tside)
        dT_all = {g: torch.zeros(len(g.names), B * K, g.ws_rows, device=dev, dtype=torch.float32)
                  for g in eng._sean_groups}
        scr_all = {g: torch.zeros(len(g.names), 2 * g.nf, 9 * L.AUX_CH, device=dev, dtype=torch.float32)
                   for g in eng._sean_groups}
        ctxs[(h, w)] = eng.mask_context(depth, masks, h, w, training=True)

    # ---- head + trunk
    h1 = _conv_train(tp, f0, "head.0", act="lrelu")
    fea_bef = _conv_train(tp, h1, "head.2", act="lrelu")
    order = net.block_order()

    # actv of every SEAN instance depends on the depth map only (and is kept for the backward pass): all of them are
    # produced on the engine's side stream, one block per SM, beside the convolutions of the main stream
    # (Engine._ActvPrefetch is the inference counterpart with rotating buffers)
    actv_pre = {}
    if eng.actv_overlap and not net.isBaseline:
        dgb_blocks = [i for i, _pos in order if i in net.which_ResBlk_depth and eng.block_scale(i) == 1]
        side = eng._side_streams.get(dev.index)
        if side is None:
            side = eng._side_streams[dev.index] = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        for i in dgb_blocks:
            blk = net.block(i)
            nf2 = 2 * blk.norm1.norm_nc
            pair = [L.act_empty(B, h, w, nf2, device=dev) for _ in range(2)]
            with torch.cuda.stream(side):
                for sean, buf in zip((blk.norm1, blk.norm2), pair):
                    L.check(lib.dasr_actv_fwd(L.ptr(depth), L.ptr(sean.mlp_mask[0].weight), L.ptr(sean.mlp_mask[0].bias),
                                              L.ptr(buf), B, h, w, nf2, 1, side.cuda_stream))
                ev = torch.cuda.Event()
                ev.record(side)
            actv_pre[i] = (pair, ev)

    def run_block(i, x):
        if i in net.which_ResBlk_depth:
            ready = tables.pop("ready", None)
            if ready is not None:       # join the style-table side stream in front of the first SEAN instance
                torch.cuda.current_stream(dev).wait_event(ready)
            res = (x.data.shape[1], x.data.shape[2])
            if res not in ctxs:
                # a depth-guided block above LR resolution: depth map and masks resized to its feature map
                # (normalization.py:58-59)
                ctxs[res] = eng.mask_context(depth, masks, res[0], res[1], training=True)
            ctx = ctxs[res]
            p = "depth-residual%d" % (i + 1)
            blk = net.block(i)
            pre = (None, None)
            if i in actv_pre:
                pre, ev = actv_pre.pop(i)
                torch.cuda.current_stream(dev).wait_event(ev)
            a = _sean_train(tp, p + ".norm1", blk.norm1, x, p + ".conv1.0", ctx, tables, dT_all, scr_all, first=True,
                            resid=None, actv_pre=pre[0])
            return _sean_train(tp, p + ".norm2", blk.norm2, a, p + ".conv2.0", ctx, tables, dT_all, scr_all, first=False,
                               resid=x, actv_pre=pre[1])
        p = "classic-residual%d" % (i + 1)
        f = _conv_train(tp, x, p + ".block.0", act="relu")
        # relu(x + conv(f)): the residual add is the conv epilogue; its backward = lazy ReLU mask, then both paths
        pk = eng._packed[p + ".block.2"]
        outd = eng._conv(f.data, p + ".block.2", act=L.ACT_RELU, resid=x.data)
        o = _T(outd, "relu")
        _dbg_mask(eng, p + ".out", outd)
        tp.use(f)
        tp.use(x)

        def bwd_res():
            dy = tp.take(o)
            tp.wgrad(dy, f.data, p + ".block.2", bias=True)
            tp.accum(x, dy)
            tp.dgrad_into(f, dy, p + ".block.2.dg")

        tp.ops.append(bwd_res)
        return o

    x = fea_bef
    for i, pos in order:
        if pos == "trunk":
            x = run_block(i, x)
    addd = L.act_like(x.data)
    L.check(lib.dasr_add(L.ptr(x.data), None, L.ptr(fea_bef.data), L.ptr(addd), addd.numel(), s))
    add = _T(addd, "none")
    xa, xb = tp.use(x), tp.use(fea_bef)

    def bwd_add():
        g = tp.take(add)
        if eng.debug is not None:
            eng.debug["feat_add1.dout"] = L.act_value(g)
        tp.accum(xa, g)
        tp.accum(xb, g)

    tp.ops.append(bwd_add)
    x = add

    # ---- tail
    if net.scale == 8:
        x = _conv_train(tp, x, "upscale1.0", act="lrelu", shuffle=True)
        x = _conv_train(tp, x, "upscale1.3", act="lrelu")
    x = run_block(order[-2][0], x)
    if net.scale >= 4:
        x = _conv_train(tp, x, "upscale2.0", act="lrelu", shuffle=True)
        x = _conv_train(tp, x, "upscale2.3", act="lrelu")
    x = run_block(order[-1][0], x)
    u3 = _conv_train(tp, x, "upscale3.0", act="lrelu", shuffle=3 if net.scale == 3 else True)
    Bo, Ho, Wo, _ = u3.data.shape
    sr = torch.empty(B, 3, Ho, Wo, device=dev, dtype=torch.float32)
    pk = eng._packed["conv_output"]
    L.check(lib.dasr_conv_out9(L.ptr(u3.data), L.ptr(pk.w), L.ptr(pk.bias), L.ptr(sr), Bo, Ho, Wo, 3, 1, s))
    tp.use(u3)

    def bwd_out(dsr):
        ap = L.act_empty(Bo, Ho, Wo, 32, device=dev)
        L.check(lib.dasr_out9_bwd_prep(L.ptr(dsr), L.ptr(sr), L.ptr(ap), L.ptr(eng._grad_view("conv_output.bias")), Bo,
                                       Ho, Wo, s))
        tp.wgrad(ap, u3.data, "conv_output", kh=9, kw=1)
        tp.dgrad_into(u3, ap, "conv_output.dg", ks=9, kw=1)

    tp.bwd_out = bwd_out
    return sr, tp


class _DepthNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, lq, depth, masks, *params):
        if not lq.is_cuda:
            raise RuntimeError("DepthNet (B200) needs CUDA tensors; there is no CPU fallback")
        lq = lq.detach().contiguous().float()
        depth = depth.detach().contiguous().float()
        masks = masks.detach().contiguous().float()
        eng.pack(training=True, force=eng.always_pack)
        sr, tp = _forward_train(eng, lq, depth, masks)
        eng.wait_pack_bwd()       # the dgrad weight copies were packed beside this forward (Engine.pack)
        ctx.eng = eng
        ctx.tape = tp
        ctx.n_params = len(params)
        return sr

    @staticmethod
    def backward(ctx, dsr):
        eng, tp = ctx.eng, ctx.tape
        if tp is None:
            raise RuntimeError("DepthNet backward called twice (activations are released after the first pass)")
        ctx.tape = None
        dsr = dsr.contiguous().float()
        eng.wait_pack_bwd()
        tp.begin_backward(dsr.device)
        eng._begin_backward(dsr.device)
        tp.bwd_out(dsr)
        for op in reversed(tp.ops):
            op()
        tp.join_backward()
        grads = eng._finish_backward()
        # break the tape <-> closure reference cycles now: left to Python's cyclic GC, the tail activations the closures
        # hold (0.32 GB per step at B=16) pile up for several steps and the caching allocator keeps growing
        tp.ops = None
        tp.bwd_out = None
        return (None, None, None, None) + tuple(grads)


def depthnet_apply(eng, lq, depth, masks):
    params = list(eng.net.parameters())
    return _DepthNetFn.apply(eng, lq, depth, masks, *params)

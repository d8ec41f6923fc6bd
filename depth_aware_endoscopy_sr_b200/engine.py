"""Host-side schedule of the DepthNet forward pass on libdasr_b200.so (C ABI: include/dasr.h).

The engine owns (a) the packed bf16 GEMM-B copies of all convolution weights -- weight-norm, SEAN alpha
folding, PixelShuffle row permutation and the ConvTranspose flip are applied by ``dasr_pack_weights`` -- and
(b) the order of kernel launches that restates ``DepthNet.forward`` (reference codes/models/modules/
sftmd_arch.py:912-950) and ``SEAN.forward`` (normalization.py:52-92):

    once         st'    = A_i_j(depthVec) for all SEAN instances             dasr_style_mix_batched
                 T      = alpha * W_s . st'   (tables of per-image dynamic filters)  ONE dasr_conv_fwd (1x1, per-image weights)
                 wdyn   = T as GEMM-B weights; mask16 = bf16 mask image       dasr_table_to_dynweights / dasr_build_mask16
    per SEAN     actv   = ReLU(conv3x3(depth, 1->2nf))                       dasr_actv_fwd
    per DGB conv y      = conv3x3(x) + b ; per-tile sum / sumsq per channel   dasr_conv_fwd EPI_STATS
                 out    = act(IN(IN(y)) * (1 + gamma) + beta [+ x])          dasr_conv_fwd EPI_SEAN over actv:
                          [gamma_o|beta_o] as the GEMM, K-DYN (dynamic 3x3 conv of mask16 with wdyn) as a K extension
                          of the same GEMM, the InstanceNorm finalize in the prologue, blend + modulation as epilogue

Activations are NHWC bf16; network input / output are NCHW fp32 like the reference.  Everything is launched on
torch's current stream; nothing here computes on the host or with torch ops.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import torch

from . import _lib as L

BF16 = torch.bfloat16


class _Packed:
    """bf16 GEMM-B weights + fp32 bias of one convolution (device tensors owned by the engine)."""
    __slots__ = ("w", "bias", "cout", "cin", "ks")

    def __init__(self, w, bias, cout, cin, ks):
        self.w, self.bias, self.cout, self.cin, self.ks = w, bias, cout, cin, ks


class _SeanGroup:
    """The SEAN instances of one width (norm_nc, latent size) in execution order: their style-table operands
    (``ws_all`` [n * 9*2nf, latent]), A_i_j pointer tables and weight-gradient slices are contiguous, so the per-image
    filter tables of the whole group (and their backward) are ONE launch each."""

    def __init__(self, nf: int, lat: int):
        self.nf, self.lat = nf, lat
        self.ws_rows = 9 * 2 * nf
        self.names: List[str] = []
        self.seans: list = []
        self.index: Dict[str, int] = {}
        self.ws_all = self.A_ptrs = self.a_ptrs = self.dA_ptrs = self.da_ptrs = None
        self.wg_tables_off = 0


class Engine:
    def __init__(self, net):
        self.net = net
        self._key = None
        self._key_bwd = None
        self.last_flat_grad = None
        # Residual stream of the trunk.  False (default): bf16 like every other activation -- measured on the golden
        # cases the fp32 copy changes the output error by < 5e-4 (0.0040 vs 0.0037 max-abs at the reference init; the
        # error is set by the bf16 GEMM operands) and costs 5 % of the step (128 B of epilogue traffic per pixel and
        # block).  True: carry an fp32 copy next to the bf16 GEMM operand (SEAN epilogue resid_f32 / out_aux_f32).
        self.fp32_residual = False
        # Option (inference): generate actv = ReLU(mlp_mask(depth)) inside the SEAN conv (GEN kernel: a fourth
        # warpgroup writes the swizzled A stages, registers redistributed with setmaxnreg) instead of a separate
        # kernel + a 67 MB tensor per SEAN instance.  Measured at B=64: 138 us against 92 + 34 us for the two-kernel
        # form -- the generator competes with the epilogue warps for issue slots and sits on the MMA's critical
        # path -- so it is OFF by default; it saves 3.5 GB of activation memory at B=64.
        self.fuse_actv = os.environ.get("DASR_FUSE_ACTV", "0") == "1"
        # actv of the next SEAN instances is produced on a low-priority side stream while the trunk convolutions run
        # (it depends on the depth map only); three rotating buffers.  DASR_ACTV_OVERLAP=0 keeps one stream.
        self.actv_overlap = os.environ.get("DASR_ACTV_OVERLAP", "1") == "1"
        # inference: the style-table chain on its own side stream beside the head convolutions (style_tables)
        self.tables_overlap = os.environ.get("DASR_TABLES_OVERLAP", "1") == "1"
        # side-stream actv: blocks produced ahead of the main stream, resident actv blocks per SM
        # (ahead: 1 -> 6.07 ms per B = 64 step, 2 -> 5.62, 3 -> 5.58, 4 .. 13 like 3; two resident blocks: 5.78)
        self.actv_ahead = max(1, int(os.environ.get("DASR_ACTV_AHEAD", "3")))
        self.actv_ctas = max(1, int(os.environ.get("DASR_ACTV_CTAS", "1")))
        # captured training step: the data-gradient weight copies are packed on a side stream beside the forward
        self.pack_overlap = os.environ.get("DASR_PACK_OVERLAP", "1") == "1"
        self._pack_bwd_ready = None
        self._side_streams = {}
        # Training backward: the weight gradients (leaves of the backward) are issued round-robin on side streams, so
        # their CTAs fill the wave tails of the data-gradient chain (256 tiles on 148 SMs at B=16) and their launch /
        # prologue / flush latency is off the critical path; beside other kernels a gradient is split over half as
        # many CTAs (half the partial-dW flushes and SM-microseconds, twice the latency -- hidden by the four
        # streams).  Measured at B=16: 7.72 (one stream) -> 7.44 (1 side stream) -> 7.25 ms (4 streams, split / 2);
        # split / 3 and / 4 are slower again (7.66, 7.78).  DASR_WGRAD_OVERLAP=0 keeps one stream.
        # Round 2, with the leaf chains on the side streams too (below): more streams take a third of the split-K CTAs
        # per launch without exposing the latency -- 6 streams, split / 2: 6.50 ms; 16 streams, split / 3: 6.39 ms
        # (12 / 3: 6.42, 24 / 4: 6.40, 16 / 4: 6.51).
        self.wgrad_overlap = os.environ.get("DASR_WGRAD_OVERLAP", "1") == "1"
        self.wgrad_streams = max(1, int(os.environ.get("DASR_WGRAD_STREAMS", "16")))
        self.wgrad_ksplit_div = max(1, int(os.environ.get("DASR_WG_KSPLIT_DIV", "3")))
        # the other leaf chains of a SEAN backward (gamma_o/beta_o data gradient -> mlp_mask gradient, K-DYN backward)
        # go to the side streams as well: the critical chain of an instance is then sean_bwd1 -> sean_bwd2 -> one
        # 64 -> 64 data gradient.  7.31 -> 6.83 ms (4 streams), 6.77 ms (6 streams).  DASR_LEAF_OVERLAP=0: main stream.
        self.leaf_overlap = os.environ.get("DASR_LEAF_OVERLAP", "1") == "1"
        self.overlap_eager = os.environ.get("DASR_OVERLAP_EAGER", "0") == "1"    # side streams outside a graph capture too
        self._wg_streams = {}
        self.use_graphs = os.environ.get("DASR_INFER_GRAPH", "1") != "0"   # replay inference from a CUDA graph (see infer)
        self.max_graphs = 3
        # Up to the bench shape (B=64 at 64x64, or eight 135x240 frames of a 1080p stream): measured at B=64, kernel by
        # kernel 5.63 ms per step with 4.3 ms of host time to issue it (the device waits on the host at every short
        # kernel of the encoder / table stage), replayed 5.53 ms including the 0.06 ms copy of the result out of the
        # graph's static buffer (tools/hostbound.py).  A recorded shape pins its activations (3.5 GB at B=64, at most
        # ``max_graphs`` shapes); larger batches stay kernel by kernel.
        self.graph_max_pixels = 64 * 64 * 64
        self._graphs = {}
        self._graph_state = None
        self.always_pack = False    # CUDA-graph capture of a training step: repack inside every forward
        self.grad_sync = None       # callable(flat fp32 grad buffer) installed by parallel.FlatDataParallel
        self._packed: Dict[str, _Packed] = {}
        self._descs: List[L.PackDesc] = []
        self._scratch = None
        self._zero_bias = None
        self._device = None
        # bench.py's instrumented pass: when a list, every launch appends
        # {family, bound, flops, bytes, e0, e1} with CUDA events recorded on the launching stream
        self.profile = None
        self.debug = None           # tests / probes: a dict that receives intermediate gradients of the training backward

    def _timed(self, family, bound, flops, nbytes, fn):
        if self.profile is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        self.profile.append(dict(family=family, bound=bound, flops=float(flops), bytes=float(nbytes), e0=e0, e1=e1))
        return r

    # ------------------------------------------------------------------------------------------ packing
    def _named(self):
        return dict(self.net.named_parameters()), dict(self.net.named_buffers())

    def _build_pack_plan(self, device):
        """Allocate destination buffers and descriptors for every convolution on the path."""
        net = self.net
        params, bufs = self._named()

        def P(name):
            return params[name] if name in params else bufs[name]

        self._packed = {}
        self._descs = []
        self._descs_bwd = []      # data-gradient (flipped / transposed) copies, packed only when training
        self._wg = {}             # name -> (offset, rows, K) into the flat fp32 weight-gradient buffer
        self._bg = {}             # name -> (offset, rows) into the flat fp32 packed-bias-gradient buffer
        self._unpack_specs = []   # recipes for dasr_unpack_grads (built lazily, see _build_unpack)
        self._wg_total = 0
        self._bg_total = 0
        self._used_params = set()
        rows_total = 0
        rows_bwd = 0
        # every active SEAN instance, in execution order, grouped by width (nf = 64 in the trunk, 32 for the blocks
        # behind upscale1 / upscale2 when which_ResBlk_depth selects them): the style-table GEMM operands of a group
        # live in ONE buffer so that its table GEMMs are a single launch with per-image weights
        # (dasr_conv_desc.w_img_rows)
        sean_names = []
        for i, _pos in net.block_order():
            if i in net.which_ResBlk_depth:
                sean_names += ["depth-residual%d.norm%d" % (i + 1, j) for j in (1, 2)]
        self._sean_names = sean_names
        self._sean_groups = []
        self._sean_group = {}
        for n in sean_names:
            m = getattr(net.block(int(n.split(".")[0][len("depth-residual"):]) - 1), n.split(".")[1])
            g = next((g for g in self._sean_groups if g.nf == m.norm_nc and g.lat == m.len_latent), None)
            if g is None:
                g = _SeanGroup(m.norm_nc, m.len_latent)
                self._sean_groups.append(g)
            g.index[n] = len(g.names)
            g.names.append(n)
            g.seans.append(m)
            self._sean_group[n] = g
        for g in self._sean_groups:
            g.ws_all = L.act_zeros(len(g.names) * g.ws_rows, g.lat, device=device)
            g.A_ptrs = torch.tensor([m.A_i_j.weight.data_ptr() for m in g.seans], dtype=torch.int64, device=device)
            g.a_ptrs = torch.tensor([m.A_i_j.bias.data_ptr() for m in g.seans], dtype=torch.int64, device=device)

        def reserve(name, rows, kdim):
            self._wg[name] = (self._wg_total, rows, kdim)
            self._wg_total += rows * kdim
            self._bg[name] = (self._bg_total, rows)
            self._bg_total += rows

        def add_dgrad(name, v, g, *, mode=L.PACK_DGRAD, shuffle_r=0):
            nonlocal rows_bwd
            if mode == L.PACK_DGRAD:
                rows, kdim, co = v.shape[1], v.shape[2] * v.shape[3] * v.shape[0], v.shape[1]
            else:   # PACK_DGRAD_CONVT: [I][O][k][k] -> rows = I
                rows, kdim, co = v.shape[0], v.shape[2] * v.shape[3] * v.shape[1], v.shape[0]
            dst = L.act_zeros(rows, kdim, device=device)
            self._descs_bwd.append(L.pack_desc(v, dst, g=g, mode=mode, shuffle_r=shuffle_r))
            rows_bwd += v.shape[0]
            self._packed[name + ".dg"] = _Packed(dst, None, co, kdim // (v.shape[2] * v.shape[3]), v.shape[2])

        def add(name, v, g, bias, *, mode=L.PACK_CONV, shuffle_r=0, rows_pad=None):
            nonlocal rows_total
            cout = v.shape[0] if mode != L.PACK_CONVT else v.shape[1]
            cin = v.shape[1] if mode != L.PACK_CONVT else v.shape[0]
            ks = v.shape[2]
            rows = rows_pad or cout
            dst = L.act_zeros(rows, ks * ks * cin, device=device)
            dbias = torch.zeros(rows, device=device, dtype=torch.float32)
            self._descs.append(L.pack_desc(v, dst, g=g, bias=bias, dst_bias=dbias, mode=mode, shuffle_r=shuffle_r))
            rows_total += v.shape[0]
            self._packed[name] = _Packed(dst, dbias, cout, cin, ks)
            reserve(name, rows, ks * ks * cin)

        def add_wn(name, dgrad=True, **kw):
            add(name, P(name + ".weight_v"), P(name + ".weight_g"), P(name + ".bias"), **kw)
            mode = kw.get("mode", L.PACK_CONV)
            if dgrad:
                add_dgrad(name, P(name + ".weight_v"), P(name + ".weight_g"),
                          mode=L.PACK_DGRAD_CONVT if mode == L.PACK_CONVT else L.PACK_DGRAD,
                          shuffle_r=kw.get("shuffle_r", 0))
            self._unpack_specs.append(dict(kind="wn", name=name, mode=mode, shuffle_r=kw.get("shuffle_r", 0)))
            self._used_params.update([name + ".weight_v", name + ".weight_g", name + ".bias"])

        # encoder.layer1 runs in its own kernel; only its weight gradient goes through the generic path
        reserve("encoder.layer1", 32, 9 * 32)
        self._unpack_specs.append(dict(kind="wn", name="encoder.layer1", mode=L.PACK_CONV, shuffle_r=0, ipack=32))
        self._used_params.update(["encoder.layer1.weight_v", "encoder.layer1.weight_g", "encoder.layer1.bias"])

        if not net.isBaseline:
            add_wn("encoder.layer2")
            add_wn("encoder.layer3")
            add_wn("encoder.layer4", mode=L.PACK_CONVT)
            add_wn("encoder.layer5")
        add_wn("head.0")
        add_wn("head.2")
        for i, _pos in net.block_order():
            blk = net.block(i)
            if i in net.which_ResBlk_depth:
                p = "depth-residual%d" % (i + 1)
                nf = blk.nf
                for j in (1, 2):
                    cn = "%s.conv%d.0" % (p, j)
                    add(cn, P(cn + ".weight"), None, P(cn + ".bias"))
                    add_dgrad(cn, P(cn + ".weight"), None)
                    self._unpack_specs.append(dict(kind="plain", name=cn))
                    self._used_params.update([cn + ".weight", cn + ".bias"])
                    n = "%s.norm%d" % (p, j)
                    ag, ab = P(n + ".alpha_gamma"), P(n + ".alpha_beta")
                    lat = blk.norm1.len_latent
                    # [gamma_o | beta_o] stacked along N, scaled by (1 - alpha); bias = blend of both branches
                    wo = L.act_zeros(2 * nf, 9 * 2 * nf, device=device)
                    bo = torch.zeros(2 * nf, device=device, dtype=torch.float32)
                    for off, x, al in ((0, "gamma", ag), (nf, "beta", ab)):
                        self._descs.append(L.pack_desc(P("%s.mlp_%s_o.weight" % (n, x)), wo, alpha=al, alpha_mode=2,
                                                       bias=P("%s.mlp_%s_o.bias" % (n, x)),
                                                       bias2=P("%s.mlp_%s_s.bias" % (n, x)), dst_bias=bo,
                                                       row_offset=off))
                        rows_total += nf
                    self._packed[n + ".gb_o"] = _Packed(wo, bo, 2 * nf, 2 * nf, 3)
                    reserve(n + ".gb_o", 2 * nf, 9 * 2 * nf)
                    wod = L.act_zeros(2 * nf, 9 * 2 * nf, device=device)
                    for off, x, al in ((0, "gamma", ag), (nf, "beta", ab)):
                        self._descs_bwd.append(L.pack_desc(P("%s.mlp_%s_o.weight" % (n, x)), wod, alpha=al, alpha_mode=2,
                                                           mode=L.PACK_DGRAD, row_offset=off, rows_per_tap=2 * nf))
                        rows_bwd += nf
                    self._packed[n + ".gb_o.dg"] = _Packed(wod, None, 2 * nf, 2 * nf, 3)
                    self._unpack_specs.append(dict(kind="sean", name=n, nf=nf, lat=lat))
                    for x in ("gamma", "beta"):
                        self._used_params.update(["%s.mlp_%s_o.weight" % (n, x), "%s.mlp_%s_o.bias" % (n, x),
                                                  "%s.mlp_%s_s.weight" % (n, x), "%s.mlp_%s_s.bias" % (n, x),
                                                  "%s.alpha_%s" % (n, x)])
                    self._used_params.update([n + ".mlp_mask.0.weight", n + ".mlp_mask.0.bias", n + ".A_i_j.weight",
                                              n + ".A_i_j.bias"])
                    # style-table GEMM operand: rows = tap * 2nf + [gamma | beta], K = latent, scaled by alpha
                    grp = self._sean_group[n]
                    k0 = grp.index[n] * grp.ws_rows
                    ws = grp.ws_all[k0:k0 + grp.ws_rows]
                    for off, x, al in ((0, "gamma", ag), (nf, "beta", ab)):
                        self._descs.append(L.pack_desc(P("%s.mlp_%s_s.weight" % (n, x)), ws, alpha=al, alpha_mode=1,
                                                       mode=L.PACK_STYLE, row_offset=off, rows_per_tap=2 * nf,
                                                       dst_stride=grp.ws_all.numel()))
                        rows_total += nf
                    self._packed[n + ".table"] = _Packed(ws, None, 9 * 2 * nf, lat, 1)
            else:
                p = "classic-residual%d" % (i + 1)
                add_wn(p + ".block.0")
                add_wn(p + ".block.2")
        # weight gradients of the style-table GEMMs: one contiguous [nS][9*2nf][lat] block per group (batched backward)
        for grp in self._sean_groups:
            grp.wg_tables_off = self._wg_total
            for n in grp.names:
                self._wg[n + ".table"] = (self._wg_total, grp.ws_rows, grp.lat)
                self._wg_total += grp.ws_rows * grp.lat
        if net.scale == 8:
            add_wn("upscale1.0", shuffle_r=2)
            add_wn("upscale1.3")
        if net.scale >= 4:
            add_wn("upscale2.0", shuffle_r=2)
            add_wn("upscale2.3")
        add_wn("upscale3.0", shuffle_r=3 if net.scale == 3 else 2)
        wq = L.act_zeros(9 * 32, 32, device=device)
        bq = torch.zeros(3, device=device, dtype=torch.float32)
        self._descs.append(L.pack_desc(P("conv_output.weight"), wq, bias=P("conv_output.bias"), dst_bias=bq,
                                       mode=L.PACK_ROWTAPS))
        rows_total += 3
        self._packed["conv_output"] = _Packed(wq, bq, 3, 32, 9)
        self._wg["conv_output"] = (self._wg_total, 32, 9 * 32)      # gradient laid out [u*3+co][t*32+ci]
        self._wg_total += 32 * 9 * 32
        wqd = L.act_zeros(32, 9 * 32, device=device)
        self._descs_bwd.append(L.pack_desc(P("conv_output.weight"), wqd, mode=L.PACK_OUT9_DGRAD))
        rows_bwd += 3
        self._packed["conv_output.dg"] = _Packed(wqd, None, 32, 32, 9)
        self._unpack_specs.append(dict(kind="out9"))
        self._used_params.update(["conv_output.weight", "conv_output.bias"])
        self._scratch = torch.zeros(max(rows_total, rows_bwd), device=device, dtype=torch.float32)
        self._scratch_bwd = torch.zeros(max(rows_bwd, 1), device=device, dtype=torch.float32)
        self._dw_flat = None          # allocated on the first backward
        self._unpack_descs = None
        self._zero_bias = torch.zeros(9 * 2 * 64, device=device, dtype=torch.float32)
        self._device = device

    def _state_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.net.parameters())

    def pack(self, force: bool = False, training: bool = False):
        """(Re)pack the weights when any parameter changed (optimizer step, load_state_dict, .to()).
        ``training`` also packs the data-gradient (flipped / transposed) copies."""
        key = self._state_key()
        if force:
            self._key = self._key_bwd = None
        if key == self._key and (not training or self._key_bwd == key):
            return
        device = next(self.net.parameters()).device
        ptrs = tuple(k[0] for k in key)
        if self._device != device or getattr(self, "_ptrs", None) != ptrs or getattr(self, "_planes", 1) != L.planes():
            self._build_pack_plan(device)      # (also when the storage mode changed: plain bf16 <-> fp32-split planes)
            self._ptrs = ptrs
            self._planes = L.planes()
            self._graphs.clear()
            self._key = self._key_bwd = None
        if key != self._key:
            L.pack_weights(self._descs, self._scratch)
            self._key = key
        if training and key != self._key_bwd:
            # the data-gradient copies are first read by the backward: inside the captured training step they are packed
            # on a side stream beside the forward (65 us at x8) and joined by wait_pack_bwd() at the end of the forward
            self._pack_bwd_ready = None
            if self.pack_overlap and torch.cuda.is_current_stream_capturing():
                side = self._side_streams.get(("pack", device.index))
                if side is None:
                    side = self._side_streams[("pack", device.index)] = torch.cuda.Stream(device=device)
                fork = torch.cuda.Event()
                fork.record(torch.cuda.current_stream(device))
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    L.pack_weights(self._descs_bwd, self._scratch_bwd)
                    self._pack_bwd_ready = torch.cuda.Event()
                    self._pack_bwd_ready.record(side)
            else:
                L.pack_weights(self._descs_bwd, self._scratch_bwd)
            self._key_bwd = key

    def wait_pack_bwd(self):
        """The current stream waits for the data-gradient weight copies packed on the side stream (see pack)."""
        ev = getattr(self, "_pack_bwd_ready", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            self._pack_bwd_ready = None

    # ------------------------------------------------------------------------------------------ backward support
    def _dw_view(self, name):
        off, rows, kdim = self._wg[name]
        return self._dw_flat[off:off + rows * kdim].view(rows, kdim)

    def _db_view(self, name):
        off, rows = self._bg[name]
        return self._db_flat[off:off + rows]

    def _grad_view(self, pname):
        off, shape, n = self._goff[pname]
        return self._g_flat[off:off + n].view(shape)

    def _wgrad_streams(self, device):
        st = self._wg_streams.get(device.index)
        if st is None:
            st = self._wg_streams[device.index] = [torch.cuda.Stream(device=device) for _ in range(self.wgrad_streams)]
        return st

    def _begin_backward(self, device):
        """Zero the flat buffers the backward kernels accumulate into (one memset each)."""
        if self._dw_flat is None:
            self._dw_flat = torch.empty(self._wg_total, device=device, dtype=torch.float32)
            self._db_flat = torch.empty(self._bg_total, device=device, dtype=torch.float32)
            self._goff = {}
            off = 0
            for name, p in self.net.named_parameters():
                self._goff[name] = (off, tuple(p.shape), p.numel())
                off += L.flat_pad(p.numel())
            self._g_flat = torch.empty(off, device=device, dtype=torch.float32)
            self._glayout = {id(p): self._goff[name][0] for name, p in self.net.named_parameters()}
            self._build_unpack()
        self._dw_flat.zero_()
        self._db_flat.zero_()
        self._g_flat.zero_()

    def _build_unpack(self):
        params, bufs = self._named()

        def P(name):
            return params[name] if name in params else bufs[name]

        def G(name):
            return self._grad_view(name) if name in params else None

        descs = []

        def U(dwp, v, dv, *, dbias_p=None, g=None, alpha=None, bias=None, bias2=None, dg=None, dbias=None, dbias2=None,
              dalpha=None, mode=L.PACK_CONV, alpha_mode=0, shuffle_r=0, row_offset=0, rows_per_tap=0, ipack=0):
            descs.append(L.UnpackDesc(L.ptr(dwp), L.ptr(dbias_p), L.ptr(v), L.ptr(g), L.ptr(alpha), L.ptr(bias),
                                      L.ptr(bias2), L.ptr(dv), L.ptr(dg), L.ptr(dbias), L.ptr(dbias2), L.ptr(dalpha),
                                      v.shape[0], v.shape[1], v.shape[2], mode, alpha_mode, shuffle_r, row_offset,
                                      rows_per_tap, ipack, 0))

        for sp in self._unpack_specs:
            if sp["kind"] == "wn":
                n = sp["name"]
                U(self._dw_view(n), P(n + ".weight_v"), G(n + ".weight_v"), dbias_p=self._db_view(n),
                  g=P(n + ".weight_g"), dg=G(n + ".weight_g"), dbias=G(n + ".bias"), mode=sp["mode"],
                  shuffle_r=sp["shuffle_r"], ipack=sp.get("ipack", 0))
            elif sp["kind"] == "plain":
                n = sp["name"]    # bias in front of an InstanceNorm: gradient exactly zero (stays zero-filled)
                U(self._dw_view(n), P(n + ".weight"), G(n + ".weight"))
            elif sp["kind"] == "sean":
                n, nf = sp["name"], sp["nf"]
                for off, x in ((0, "gamma"), (nf, "beta")):
                    al, dal = P("%s.alpha_%s" % (n, x)), G("%s.alpha_%s" % (n, x))
                    U(self._dw_view(n + ".gb_o"), P("%s.mlp_%s_o.weight" % (n, x)), G("%s.mlp_%s_o.weight" % (n, x)),
                      dbias_p=self._db_view(n + ".gb_o"), alpha=al, alpha_mode=2, bias=P("%s.mlp_%s_o.bias" % (n, x)),
                      bias2=P("%s.mlp_%s_s.bias" % (n, x)), dbias=G("%s.mlp_%s_o.bias" % (n, x)),
                      dbias2=G("%s.mlp_%s_s.bias" % (n, x)), dalpha=dal, row_offset=off)
                    U(self._dw_view(n + ".table"), P("%s.mlp_%s_s.weight" % (n, x)), G("%s.mlp_%s_s.weight" % (n, x)),
                      alpha=al, alpha_mode=1, dalpha=dal, mode=L.PACK_STYLE, row_offset=off, rows_per_tap=2 * nf)
            elif sp["kind"] == "out9":
                U(self._dw_view("conv_output"), P("conv_output.weight"), G("conv_output.weight"), mode=L.PACK_ROWTAPS)
        self._unpack_descs = (L.UnpackDesc * len(descs))(*descs)
        # gradient pointer tables of the batched A_i_j backward (dasr_style_mix_bwd_batched)
        dev = self._g_flat.device
        for grp in self._sean_groups:
            grp.dA_ptrs = torch.tensor([self._grad_view(n + ".A_i_j.weight").data_ptr() for n in grp.names],
                                       dtype=torch.int64, device=dev)
            grp.da_ptrs = torch.tensor([self._grad_view(n + ".A_i_j.bias").data_ptr() for n in grp.names],
                                       dtype=torch.int64, device=dev)

    def _finish_backward(self):
        """Packed-layout gradients -> parameter gradients; returns one tensor (or None) per parameter, all of them
        views of ONE flat fp32 buffer (kept as ``last_flat_grad`` for the data-parallel all-reduce)."""
        L.check(L.load().dasr_unpack_grads(self._unpack_descs, len(self._unpack_descs), L.stream_ptr()))
        flat = self._g_flat.clone()
        if self.grad_sync is not None:       # data parallel: one all-reduce of the whole flat buffer (parallel.py)
            self.grad_sync(flat)
        L.register_flat_grad(flat, self._glayout)    # lets FusedAdam consume the buffer in place (optim.py)
        self.last_flat_grad = flat
        out = []
        for name, p in self.net.named_parameters():
            if name in self._used_params and p.requires_grad:
                off, shape, n = self._goff[name]
                out.append(flat[off:off + n].view(shape))
            else:
                out.append(None)
        return out

    # ------------------------------------------------------------------------------------------ helpers
    def _conv(self, x, name, *, epi=L.EPI_STORE, act=L.ACT_NONE, subsample=1, out=None, **kw):
        """``x`` None: the kernel generates its A operand (``gen_depth`` / ``shape`` in kw)."""
        pk = self._packed[name]
        B, H, W, cin = x.shape if x is not None else kw["shape"]
        dev = x.device if x is not None else kw["gen_depth"].device
        if out is None:
            if epi == L.EPI_SHUFFLE2:
                out = L.act_empty(B, 2 * H, 2 * W, pk.cout // 4, device=dev)
            elif epi == L.EPI_SEAN:
                out = L.act_empty(B, H, W, pk.cout // 2, device=dev)
            elif subsample == 2:
                out = L.act_empty(B, (H + 1) // 2, (W + 1) // 2, pk.cout, device=dev)
            else:
                out = L.act_empty(B, H, W, pk.cout, device=dev)
        if self.profile is None:
            return L.conv_fwd(x, pk.w, pk.bias, out, Cout=pk.cout, ks=pk.ks, epi=epi, act=act, subsample=subsample,
                              **kw)
        flops = 2.0 * B * H * W * pk.cout * cin * pk.ks * pk.ks           # algorithmic (stride-1 grid)
        if subsample == 2:
            flops /= 4.0
        nbytes = (x.numel() * 2 if x is not None else 0) + out.numel() * out.element_size() + pk.w.numel() * 2
        for t in (kw.get("resid"), kw.get("y"), kw.get("gb_s")):
            if t is not None:
                nbytes += t.numel() * 2
        epi_name = {L.EPI_STORE: "store", L.EPI_STATS: "stats", L.EPI_SEAN: "sean", L.EPI_SHUFFLE2: "shuffle",
                    L.EPI_NCHW_F32: "nchw"}[epi]
        family = "conv%dx%d_%dto%d_%s" % (pk.ks, pk.ks, cin, pk.cout, epi_name)
        return self._timed(family, "tensor", flops, nbytes,
                           lambda: L.conv_fwd(x, pk.w, pk.bias, out, Cout=pk.cout, ks=pk.ks, epi=epi, act=act,
                                              subsample=subsample, **kw))

    def style_tables(self, vec, side=None):
        """The per-image dynamic-filter tables of ALL SEAN instances, three launches per width group:
        stp[s] = A_i_j^(s)(depthVec) (dasr_style_mix_batched), T[s] = alpha^(s) W_s^(s) . stp[s] (one 1x1
        dasr_conv_fwd over the instances of the group with per-image weights) and its GEMM-B form for the K-DYN
        extension.  Returns {group: (stp_all [nS,1,B*K,L], table_all [nS,1,B*K,9*2nf], wdyn_all [nS,B*2nf,9*16])};
        ``sean_tables(tables, n)`` picks instance ``n``.

        ``side``: a stream to run the launches on (forked from the current stream here; the CALLER joins with
        ``tables["ready"]`` before the first consumer).  The chain depends on the encoder output only, so in inference it
        runs beside the head convolutions and the first trunk convolution instead of in front of them.  All buffers
        are allocated on the current stream, like Engine._ActvPrefetch's."""
        lib = L.load()
        B, K, lat = vec.shape
        res = {}
        bufs = []
        for grp in self._sean_groups:
            nS = len(grp.names)
            stp_all = L.act_empty(nS, 1, B * K, lat, device=vec.device)
            table_all = L.act_empty(nS, 1, B * K, grp.ws_rows, device=vec.device)
            # GEMM-B form of every table for the K-DYN extension of the SEAN GEMM: [nS][B*2nf][9*16]
            # (fp32-split planes: the planes of an instance's filters sit inside its own slice -- see include/dasr.h)
            wdyn_all = torch.empty(nS, L.planes(), B * 2 * grp.nf, 9 * 16, device=vec.device, dtype=BF16)
            bufs.append((grp, stp_all, table_all, wdyn_all))
            res[grp] = (stp_all, table_all, wdyn_all[:, 0])

        def launch():
            s = L.stream_ptr()
            for grp, stp_all, table_all, wdyn_all in bufs:
                nS = len(grp.names)
                self._timed("style_mix", "hbm", 0, vec.numel() * 4 + stp_all.numel() * 2,
                            lambda: L.check(lib.dasr_style_mix_batched(L.ptr(vec), L.ptr(grp.A_ptrs), L.ptr(grp.a_ptrs),
                                                                       L.ptr(stp_all), nS, B, K, lat, s)))
                self._timed("style_table_gemm", "tensor", 2.0 * nS * B * K * lat * grp.ws_rows,
                            stp_all.numel() * 2 + table_all.numel() * 2 + grp.ws_all.numel() * 2,
                            lambda: L.conv_fwd(stp_all, grp.ws_all, self._zero_bias, table_all, Cout=grp.ws_rows, ks=1,
                                               w_img_rows=grp.ws_rows))
                self._timed("table_to_dynweights", "hbm", 0, table_all.numel() * 2 + wdyn_all.numel() * 2,
                            lambda: L.check(lib.dasr_table_to_dynweights(L.ptr(table_all), L.ptr(wdyn_all), nS * B, K,
                                                                         2 * grp.nf, B, s)))

        if side is None or not bufs:
            launch()
            return res
        main = torch.cuda.current_stream(vec.device)
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            launch()
            ready = torch.cuda.Event()
            ready.record(side)
        res["ready"] = ready
        return res

    def sean_tables(self, tables, n: str):
        """(stp, table, wdyn) of SEAN instance ``n`` out of ``style_tables``' result."""
        grp = self._sean_group[n]
        t = tables[grp]
        k = grp.index[n]
        return t[0][k], t[1][k], t[2][k]

    def mask_context(self, depth, masks, H: int, W: int, training: bool = False):
        """What the SEAN instances of one resolution read of the depth map and the depth masks: at LR resolution the
        inputs themselves, above it their nearest-neighbour resize (normalization.py:58-59) -- plus the derived
        operands: the bf16 mask image of the K-DYN extension and, for the backward, labels / general-mask flag / aux."""
        lib = L.load()
        s = L.stream_ptr()
        B, K, h, w = masks.shape
        dev = masks.device
        if (H, W) != (h, w):
            d_up = torch.empty(B, 1, H, W, device=dev, dtype=torch.float32)
            m_up = torch.empty(B, K, H, W, device=dev, dtype=torch.float32)
            L.check(lib.dasr_nearest_up(L.ptr(depth), L.ptr(d_up), B, h, w, H, W, s))
            L.check(lib.dasr_nearest_up(L.ptr(masks), L.ptr(m_up), B * K, h, w, H, W, s))
            depth, masks = d_up, m_up
        ctx = dict(depth=depth, masks=masks)
        ctx["mask16"] = torch.empty(B, H, W, 16, device=dev, dtype=BF16)
        L.check(lib.dasr_build_mask16(L.ptr(masks), L.ptr(ctx["mask16"]), B, K, H, W, s))
        if training:
            ctx["labels"] = torch.empty(B, H, W, device=dev, dtype=torch.uint8)
            ctx["flag"] = torch.zeros(1, device=dev, dtype=torch.int32)
            L.check(lib.dasr_mask_labels(L.ptr(masks), L.ptr(ctx["labels"]), L.ptr(ctx["flag"]), B, K, H, W, s))
            ctx["aux"] = torch.empty(B, H, W, L.AUX_CH, device=dev, dtype=BF16)
            L.check(lib.dasr_build_aux(L.ptr(ctx["labels"]), L.ptr(depth), L.ptr(ctx["aux"]), B, K, H, W, s))
        return ctx

    def _sean_actv(self, sean, depth, out=None, ctas_per_sm=0):
        """actv = ReLU(mlp_mask(depth)) of one SEAN instance (normalization.py:37-40,61)."""
        lib = L.load()
        B, _, H, W = depth.shape
        nf2 = 2 * sean.norm_nc
        s = L.stream_ptr()
        actv = out if out is not None else L.act_empty(B, H, W, nf2, device=depth.device)
        self._timed("actv", "hbm", 0, actv.numel() * 2 + depth.numel() * 4,
                    lambda: L.check(lib.dasr_actv_fwd(L.ptr(depth), L.ptr(sean.mlp_mask[0].weight),
                                                      L.ptr(sean.mlp_mask[0].bias), L.ptr(actv), B, H, W, nf2,
                                                      ctas_per_sm, s)))
        return actv

    class _ActvPrefetch:
        """Produces actv of the SEAN instances, in network order, on a side stream, one depth-guided block (two
        instances) at a time and ``ahead`` blocks in front of the main stream.  take() makes the main stream wait for
        the next block's pair and returns its two buffers; done() (after the block's kernels were launched) hands the
        buffers back for block k + ahead.  One cross-stream wait and one record per BLOCK: the four convolutions of a
        block stay an unbroken programmatic-dependent-launch chain."""

        def __init__(self, eng, blocks, depth, nf2, ahead=None):
            dev = depth.device
            ahead = eng.actv_ahead if ahead is None else ahead
            self.eng, self.blocks, self.depth = eng, blocks, depth
            st = eng._side_streams.get(dev.index)
            if st is None:
                st = eng._side_streams[dev.index] = torch.cuda.Stream(device=dev)
            self.side = st
            self.main = torch.cuda.current_stream(dev)
            B, _, H, W = depth.shape
            self.slots = [[L.act_empty(B, H, W, nf2, device=dev) for _ in range(2)]
                          for _ in range(min(ahead, len(blocks)))]
            self.ready = [None] * len(blocks)
            self.k = 0
            fork = torch.cuda.Event()
            fork.record(self.main)
            self.side.wait_event(fork)
            for i in range(len(self.slots)):
                self._issue(i)

        def _issue(self, i):
            with torch.cuda.stream(self.side):
                for sean, buf in zip(self.blocks[i], self.slots[i % len(self.slots)]):
                    # one block per SM: it fits beside the convolution kernels and never keeps their blocks waiting
                    self.eng._sean_actv(sean, self.depth, out=buf, ctas_per_sm=self.eng.actv_ctas)
                ev = torch.cuda.Event()
                ev.record(self.side)
            self.ready[i] = ev

        def take(self):
            self.main.wait_event(self.ready[self.k])
            return self.slots[self.k % len(self.slots)]

        def done(self):
            k = self.k
            self.k += 1
            nxt = k + len(self.slots)
            if nxt < len(self.blocks):
                free = torch.cuda.Event()
                free.record(self.main)
                self.side.wait_event(free)
                self._issue(nxt)

    def _dgb(self, p: str, blk, x, x32, depth, mask16, tables, prefetch=None):
        """Depth_Residual_Block_Mask.forward (sftmd_arch.py:826-834).  ``x`` is the bf16 copy of the block input
        (GEMM operand), ``x32`` its fp32 residual stream (None for the first block: the bf16 tensor is exact).
        Returns (bf16 output, fp32 output)."""
        lib = L.load()
        B, H, W, nf = x.shape
        s = L.stream_ptr()
        nslots = L.conv_stats_slots(B, H, W, nf, nf)
        stats = torch.empty(2, B, nslots, nf, 2, device=x.device, dtype=torch.float32)
        cur = x
        out32 = (torch.empty(B, H, W, nf, device=x.device, dtype=torch.float32)
                 if self.fp32_residual and L.planes() == 1 else None)
        # actv generated inside the SEAN conv (no actv tensor, no actv launch) when the geometry leaves room for it
        gen = self.fuse_actv and nf == 64 and lib.dasr_conv_gen_ok(H, W) == 1 and L.planes() == 1
        pre = prefetch.take() if (prefetch is not None and not gen) else None
        for j, sean in ((1, blk.norm1), (2, blk.norm2)):
            n = "%s.norm%d" % (p, j)
            # conv + per-tile statistics; the double-InstanceNorm coefficients are finalised inside the SEAN conv
            y = self._conv(cur, "%s.conv%d.0" % (p, j), epi=L.EPI_STATS, stats=stats[j - 1])
            wdyn = self.sean_tables(tables, n)[2]     # K-DYN runs inside the SEAN GEMM as a K extension
            if gen:
                actv = None
                akw = dict(shape=(B, H, W, 2 * nf), gen_depth=depth, gen_w=sean.mlp_mask[0].weight,
                           gen_b=sean.mlp_mask[0].bias)
            elif pre is not None:
                actv = pre[j - 1]
                akw = {}
            else:
                actv = self._sean_actv(sean, depth)
                akw = {}
            if j == 1:
                cur = self._conv(actv, n + ".gb_o", epi=L.EPI_SEAN, inner_relu=1, y=y, stats=stats[0], dyn_x=mask16,
                                 dyn_w=wdyn, **akw)
            else:
                cur = self._conv(actv, n + ".gb_o", epi=L.EPI_SEAN, act=L.ACT_RELU, y=y, stats=stats[1], dyn_x=mask16,
                                 dyn_w=wdyn, resid=x if x32 is None else None, resid_f32=x32,
                                 out_aux_f32=out32, **akw)
        if pre is not None:
            prefetch.done()
        return cur, out32

    def block_scale(self, i: int) -> int:
        """Resolution of block ``i``'s feature map relative to the LR input (sftmd_arch.py:932-944): the trunk runs at
        LR resolution, block nb-2 behind upscale1 (x2 at scale 8), block nb-1 behind upscale2 (x2 at scale 4, x4 at 8)."""
        net = self.net
        pos = dict(net.block_order()).get(i, "trunk")
        up1 = 2 if net.scale == 8 else 1
        up2 = up1 * (2 if net.scale >= 4 else 1)
        return {"trunk": 1, "up1": up1, "up2": up2}[pos]

    def _classic(self, p: str, x):
        """Classic_Residual_Block.forward (sftmd_arch.py:147-151)."""
        f = self._conv(x, p + ".block.0", act=L.ACT_RELU)
        return self._conv(f, p + ".block.2", act=L.ACT_RELU, resid=x)

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, lq: torch.Tensor, depth: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        net = self.net
        if torch.is_grad_enabled() and any(p.requires_grad for p in net.parameters()):
            from . import autograd as _ag   # training path (forward + backward kernels)
            return _ag.depthnet_apply(self, lq, depth, masks)
        return self.infer(lq, depth, masks)

    @torch.no_grad()
    def infer(self, lq: torch.Tensor, depth: torch.Tensor, masks: torch.Tensor, cap: dict = None,
              clamp: bool = True, frames: bool = False) -> torch.Tensor:
        """Inference.  A forward is ~100 kernel launches; issued from Python that is 2.7 ms of host time, more than
        the kernels of a single 1080p frame need (1.5 ms), so for small batches (``graph_max_pixels``) the schedule is
        replayed from a CUDA graph from the third call with the same input shape on (static input / output buffers;
        the result is returned as a copy): 2.7 -> 0.85 ms per 64x64 frame, 2.7 -> 1.3 ms per 1080p frame.  ``cap`` /
        ``profile`` / an ongoing stream capture use the kernel-by-kernel schedule.
        ``frames``: return uint8 BGR frames [B, sH, sW, 3] -- ``util.tensor2img`` fused into the store of the output
        convolution (dasr_conv_out9_frames) -- instead of the fp32 SR tensor."""
        if (not self.use_graphs or cap is not None or self.profile is not None or not lq.is_cuda or L.planes() > 1
                or lq.shape[0] * lq.shape[2] * lq.shape[3] > self.graph_max_pixels
                or torch.cuda.is_current_stream_capturing()):
            return self._infer_eager(lq, depth, masks, cap=cap, clamp=clamp, frames=frames)
        self.pack()
        key = (tuple(lq.shape), tuple(masks.shape), lq.device.index, bool(clamp), self.fp32_residual, bool(frames))
        if self._graph_state != self._key:            # parameters changed: every recorded schedule is stale
            self._graphs.clear()
            self._graph_state = self._key
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= self.max_graphs:   # each graph owns its activation pool: keep a few shapes only
                self._graphs.pop(next(iter(self._graphs)))
            ent = self._graphs[key] = dict(calls=0, graph=None)
        ent["calls"] += 1
        if ent["graph"] is None:
            if ent["calls"] <= 2:
                return self._infer_eager(lq, depth, masks, clamp=clamp, frames=frames)
            ins = [torch.empty(t.shape, device=t.device, dtype=torch.float32) for t in (lq, depth, masks)]
            for d, src in zip(ins, (lq, depth, masks)):
                d.copy_(src)
            g = torch.cuda.CUDAGraph()
            n0 = int(L.load().dasr_launch_count())
            with torch.cuda.graph(g):
                out = self._infer_eager(*ins, clamp=clamp, frames=frames)
            ent.update(graph=g, ins=ins, out=out, launches=int(L.load().dasr_launch_count()) - n0)
        for d, src in zip(ent["ins"], (lq, depth, masks)):
            d.copy_(src)
        ent["graph"].replay()
        L.note_replayed_launches(ent["launches"])
        return ent["out"].clone()

    @torch.no_grad()
    def _infer_eager(self, lq: torch.Tensor, depth: torch.Tensor, masks: torch.Tensor, cap: dict = None,
                     clamp: bool = True, frames: bool = False) -> torch.Tensor:
        """Inference schedule, kernel by kernel.  ``cap`` (tests only) receives intermediate tensors in the engine's
        own layouts (NHWC bf16 activations); ``clamp=False`` returns the pre-clamp output of conv_output."""
        net = self.net
        lib = L.load()
        if not lq.is_cuda:
            raise RuntimeError("DepthNet (B200) needs CUDA tensors; there is no CPU fallback")
        lq = lq.contiguous().float()
        depth = depth.contiguous().float()
        masks = masks.contiguous().float()
        B, _, h, w = lq.shape
        K = masks.shape[1]
        dev = lq.device
        self.pack()
        s = L.stream_ptr()

        # ---- encoder (sftmd_arch.py:771-783)
        enc = net.encoder
        f0 = L.act_empty(B, h, w, 32, device=dev)
        L.check(lib.dasr_conv_first(L.ptr(lq), L.ptr(enc.layer1.weight_v), L.ptr(enc.layer1.weight_g),
                                    L.ptr(enc.layer1.bias), L.ptr(f0), B, h, w, s))
        vec = labels = flag = tables = mask16 = None
        if not net.isBaseline:
            e2 = self._conv(f0, "encoder.layer2", subsample=2, act=L.ACT_LRELU)
            e3 = self._conv(e2, "encoder.layer3", subsample=2, act=L.ACT_LRELU)
            h3, w3 = e3.shape[1], e3.shape[2]
            z = L.act_empty(B, 2 * h3 - 1, 2 * w3 - 1, 128, device=dev)
            L.check(lib.dasr_zero_insert2(L.ptr(e3), L.ptr(z), B, h3, w3, 128, s))
            e4 = self._conv(z, "encoder.layer4", act=L.ACT_LRELU)
            e5 = self._conv(e4, "encoder.layer5", subsample=2)
            lat = e5.shape[3]
            vec = torch.empty(B, K, lat, device=dev, dtype=torch.float32)
            L.check(lib.dasr_region_pool_fwd(L.ptr(e5), L.ptr(masks), L.ptr(vec), None, None, B, e5.shape[1],
                                             e5.shape[2], lat, K, h, w, s))
            # the style-table chain (three small launches, ~130 us at B = 64) runs beside the head convolutions
            tside = None
            if self.tables_overlap and self.profile is None and cap is None:
                tside = self._side_streams.get(("tables", dev.index))
                if tside is None:
                    tside = self._side_streams[("tables", dev.index)] = torch.cuda.Stream(device=dev)
            tables = self.style_tables(vec, side=tside)
            mask16 = self.mask_context(depth, masks, h, w)["mask16"]
            if cap is not None:
                labels = torch.empty(B, h, w, device=dev, dtype=torch.uint8)
                flag = torch.zeros(1, device=dev, dtype=torch.int32)
                L.check(lib.dasr_mask_labels(L.ptr(masks), L.ptr(labels), L.ptr(flag), B, K, h, w, s))
                cap.update(e5=e5, depthVec=vec, labels=labels, flag=flag)

        # ---- head + trunk (sftmd_arch.py:920-931)
        fea_bef = self._conv(self._conv(f0, "head.0", act=L.ACT_LRELU), "head.2", act=L.ACT_LRELU)
        x, x32 = fea_bef, None      # x32: fp32 residual stream of the trunk (bf16 copies feed the GEMMs)
        order = net.block_order()

        prefetch = None
        # (the instrumented pass of bench.py times kernels one by one: no overlap there unless forced by a probe)
        if (self.actv_overlap and (self.profile is None or self.actv_overlap == "force") and cap is None
                and not self.fuse_actv):
            # (the instances at LR resolution; a depth-guided block behind upscale1 / upscale2 produces its own actv)
            blocks = [(net.block(i).norm1, net.block(i).norm2) for i, _pos in order
                      if i in net.which_ResBlk_depth and self.block_scale(i) == 1]
            if blocks:
                prefetch = Engine._ActvPrefetch(self, blocks, depth, 2 * blocks[0][0].norm_nc)

        def run_block(i, x, x32):
            if i in net.which_ResBlk_depth:
                p = "depth-residual%d" % (i + 1)
                ready = tables.pop("ready", None)
                if ready is not None:       # join the style-table side stream in front of the first SEAN instance
                    torch.cuda.current_stream(dev).wait_event(ready)
                if x.shape[1] != h or x.shape[2] != w:
                    # a depth-guided block above LR resolution: depth map and masks resized to its feature map
                    hr = self.mask_context(depth, masks, x.shape[1], x.shape[2])
                    return self._dgb(p, net.block(i), x, x32, hr["depth"], hr["mask16"], tables)
                return self._dgb(p, net.block(i), x, x32, depth, mask16, tables, prefetch=prefetch)
            return self._classic("classic-residual%d" % (i + 1), x), None

        for i, pos in order:
            if pos == "trunk":
                x, x32 = run_block(i, x, x32)
                if cap is not None:
                    cap["block%d.out" % (i + 1)] = x
        if cap is not None:
            cap["fea_bef"] = fea_bef
        add = L.act_like(x)
        self._timed("add", "hbm", 0, x.numel() * (2 + 2 + (4 if x32 is not None else 2)),
                    lambda: L.check(lib.dasr_add(L.ptr(x), L.ptr(x32), L.ptr(fea_bef), L.ptr(add), x.numel(), s)))
        x = add

        # ---- tail (sftmd_arch.py:932-950)
        if net.scale == 8:
            x = self._conv(x, "upscale1.0", epi=L.EPI_SHUFFLE2, act=L.ACT_LRELU)
            x = self._conv(x, "upscale1.3", act=L.ACT_LRELU)
        x, _ = run_block(order[-2][0], x, None)
        if net.scale >= 4:
            x = self._conv(x, "upscale2.0", epi=L.EPI_SHUFFLE2, act=L.ACT_LRELU)
            x = self._conv(x, "upscale2.3", act=L.ACT_LRELU)
        x, _ = run_block(order[-1][0], x, None)
        if net.scale == 3:
            # 64 -> 288 convolution (shuffled channel order, LeakyReLU commutes with the permutation), then
            # PixelShuffle(3) as a copy kernel (sftmd_arch.py:904-908)
            u = self._conv(x, "upscale3.0", act=L.ACT_LRELU)
            x = L.act_empty(B, 3 * u.shape[1], 3 * u.shape[2], u.shape[3] // 9, device=dev)
            L.check(lib.dasr_pixel_shuffle(L.ptr(u), L.ptr(x), B, u.shape[1], u.shape[2], u.shape[3] // 9, 3, s))
        else:
            x = self._conv(x, "upscale3.0", epi=L.EPI_SHUFFLE2, act=L.ACT_LRELU)
        if net.min != 0.0 or net.max != 1.0:
            raise NotImplementedError("the fused output epilogue clamps to [0,1] (the only range define_G builds)")
        if cap is not None:
            cap["feat_up3"] = x
        pk = self._packed["conv_output"]
        Bo, Ho, Wo, _ = x.shape
        if frames:
            # conv_output + clamp + util.tensor2img in one kernel: no fp32 frames in HBM
            out = torch.empty(B, Ho, Wo, 3, device=dev, dtype=torch.uint8)
            self._timed("conv_out9", "tensor", 2.0 * Bo * Ho * Wo * 3 * 32 * 81, x.numel() * 2 + out.numel(),
                        lambda: L.check(lib.dasr_conv_out9_frames(L.ptr(x), L.ptr(pk.w), L.ptr(pk.bias), L.ptr(out), Bo, Ho,
                                                                  Wo, float(net.min), float(net.max), s)))
            return out
        out = torch.empty(B, 3, Ho, Wo, device=dev, dtype=torch.float32)
        # algorithmic bytes: read feat_up3 once + write the fp32 frames
        self._timed("conv_out9", "tensor", 2.0 * Bo * Ho * Wo * 3 * 32 * 81, x.numel() * 2 + out.numel() * 4,
                    lambda: L.check(lib.dasr_conv_out9(L.ptr(x), L.ptr(pk.w), L.ptr(pk.bias), L.ptr(out), Bo, Ho, Wo, 3,
                                                       1 if clamp else 0, s)))
        return out

"""Adam on the K-ADAM kernel (C ABI: dasr_adam_step) -- the optimiser step of the reference's training loop
(``torch.optim.Adam(optim_params, lr, weight_decay, betas)``, codes/models/F_model_depthCond.py:99-101,192).

``FusedAdam`` is a ``torch.optim.Optimizer`` with Adam's constructor, ``param_groups`` (so the reference's
``lr_scheduler`` classes and ``update_learning_rate`` keep working), ``step`` / ``zero_grad`` / ``state_dict``.
It re-homes every parameter of a group into ONE flat fp32 buffer (the ``.data`` of each parameter becomes a view),
keeps ``exp_avg`` / ``exp_avg_sq`` flat as well, and steps the whole group with a single kernel launch.  When all
gradients are views of one flat buffer in parameter order (what ``DepthNet``'s backward produces,
``Engine.last_flat_grad``) they are consumed in place with one launch; every other parameter (e.g. the 10 weights
of the dynamic loss) is stepped with one launch of its own.

Inside the engine's flat buffer a parameter the network never uses has a zero gradient.  With zero Adam state that
is a no-op (m = v = 0 -> update 0), which matches torch skipping ``grad is None`` parameters (SURVEY.md section 7,
"unused parameters").
"""
from __future__ import annotations

from typing import Iterable

import torch

from . import _lib as L


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, capturable: bool = False):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay or 0.0))
        self._flat = {}      # group index -> dict(p, m, v, g, offsets, step)
        # capturable: the step-dependent scalars (lr / bias corrections) are read from device memory, so step()
        # can be recorded once in a CUDA graph; call advance() before every step() / graph replay
        self.capturable = capturable

    # ------------------------------------------------------------------ flat storage
    def _ensure_flat(self, gi, group):
        st = self._flat.get(gi)
        params = [p for p in group["params"] if p.requires_grad]
        if st is not None and all(p.data_ptr() == st["p"].data_ptr() + 4 * off
                                  for p, off in zip(params, st["offsets"])):
            return st
        if not params:
            return None
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam (B200) needs CUDA parameters; there is no CPU fallback")
        offsets, n = [], 0
        for p in params:
            if p.dtype != torch.float32 or p.device != dev:
                raise RuntimeError("FusedAdam: all parameters of a group must be fp32 on one device")
            offsets.append(n)
            n += L.flat_pad(p.numel())               # keep every slice 16-byte aligned (same rule as the engine)
        flat = torch.zeros(n, device=dev, dtype=torch.float32)
        old = st
        for p, off in zip(params, offsets):
            flat[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = flat[off:off + p.numel()].view(p.shape)
        st = dict(p=flat, m=torch.zeros_like(flat), v=torch.zeros_like(flat), g=None, offsets=offsets, params=params,
                  step=0 if old is None else old["step"],
                  pstep=[0] * len(params) if old is None or len(old["pstep"]) != len(params) else old["pstep"])
        if old is not None and old["m"].numel() == n:     # parameters were moved (.to / load): keep the moments
            st["m"].copy_(old["m"])
            st["v"].copy_(old["v"])
        self._flat[gi] = st
        return st

    def _segments(self, st, weight_decay=0.0):
        """[(offset, n, grad tensor)] covering the flat parameter buffer.  Parameters whose gradients are views of
        ONE flat buffer with this optimiser's layout (``Engine.last_flat_grad``, registered with its layout
        {id(param): offset} in ``_lib.flat_grads()``; zero where the network leaves a gradient None) are stepped as a single segment straight
        from that buffer; any other parameter is its own segment; ``grad is None`` outside a flat buffer is skipped
        like torch.optim.Adam does."""
        params, offsets = st["params"], st["offsets"]
        segs, covered = [], set()
        for base, lay in L.flat_grads():
            idx = [i for i, p in enumerate(params) if id(p) in lay and i not in covered]
            if not idx or idx != list(range(idx[0], idx[-1] + 1)):
                continue
            i0, i1 = idx[0], idx[-1]
            g0 = lay[id(params[i0])]
            n = offsets[i1] + params[i1].numel() - offsets[i0]
            same_layout = all(lay[id(params[i])] - offsets[i] == g0 - offsets[i0] for i in idx)
            in_place = all(p.grad is None or p.grad.data_ptr() == base.data_ptr() + 4 * lay[id(p)]
                           for p in params[i0:i1 + 1])
            if same_layout and in_place and any(p.grad is not None for p in params[i0:i1 + 1]) \
                    and (base.data_ptr() + 4 * g0) % 16 == 0 and g0 + n <= base.numel() and base.device == st["p"].device:
                # A parameter the network never uses has a zero slice in the flat buffer.  Without weight decay a zero
                # gradient on zero Adam state is a no-op, so the whole range is ONE launch; with weight decay the
                # kernel's g += wd * p would decay it (torch.optim.Adam skips grad-None parameters entirely), so
                # the range is split into the runs of parameters that do have a gradient.
                if weight_decay:
                    runs, cur = [], []
                    for i in idx:
                        if params[i].grad is not None:
                            cur.append(i)
                        elif cur:
                            runs.append(cur)
                            cur = []
                    if cur:
                        runs.append(cur)
                else:
                    runs = [idx]
                for run in runs:
                    r0, r1 = run[0], run[-1]
                    gr = lay[id(params[r0])]
                    nr = offsets[r1] + params[r1].numel() - offsets[r0]
                    segs.append((offsets[r0], nr, base[gr:gr + nr], list(run)))
                covered.update(range(i0, i1 + 1))
        for i, p in enumerate(params):
            if i in covered or p.grad is None:
                continue
            g = p.grad.detach()
            if g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() % 16:
                g = g.float().contiguous().clone()
            segs.append((offsets[i], p.numel(), g.reshape(-1), [i]))
        return segs

    def advance(self):
        """Host half of a step in capturable mode: bump the step counter and upload {lr/(1-b1^t), sqrt(1-b2^t)}."""
        for gi, group in enumerate(self.param_groups):
            st = self._ensure_flat(gi, group)
            if st is None:
                continue
            st["step"] += 1
            st["pstep"] = [st["step"]] * len(st["params"])
            b1, b2 = group["betas"]
            t = st["step"]
            host = torch.tensor([group["lr"] / (1.0 - b1 ** t), (1.0 - b2 ** t) ** 0.5], dtype=torch.float32)
            if "scal" not in st:
                st["scal"] = torch.empty(2, device=st["p"].device, dtype=torch.float32)
            st["scal"].copy_(host, non_blocking=True)

    # ------------------------------------------------------------------ Optimizer API
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        for gi, group in enumerate(self.param_groups):
            st = self._ensure_flat(gi, group)
            if st is None:
                continue
            segs = self._segments(st, group["weight_decay"])
            if not segs:
                continue
            b1, b2 = group["betas"]
            pstep = st["pstep"]
            scal = None
            if self.capturable:
                if "scal" not in st:
                    raise RuntimeError("FusedAdam(capturable=True): call advance() before step()")
                scal = st["scal"]
            else:
                st["step"] += 1
            for off, n, g, members in segs:
                # torch keeps one step counter per parameter (a parameter without gradient does not advance);
                # the members of a flat segment always step together
                if not self.capturable:
                    for i in members:
                        pstep[i] += 1
                L.check(lib.dasr_adam_step(L.ptr(st["p"][off:off + n]), L.ptr(g), L.ptr(st["m"][off:off + n]),
                                           L.ptr(st["v"][off:off + n]), n, float(group["lr"]), float(b1), float(b2),
                                           float(group["eps"]), float(group["weight_decay"]), pstep[members[0]], 1.0,
                                           L.ptr(scal), L.stream_ptr()))
            # the kernel wrote through raw pointers: advance the version counters like an in-place torch op
            # would (weight caches such as Engine.pack key on ``p._version``)
            torch.autograd.graph.increment_version(st["params"])
        return loss

    # ------------------------------------------------------------------ checkpoint interchange with torch.optim.Adam
    def _param_indices(self):
        """{id(param): index} in the numbering of Optimizer.state_dict() (group order, every parameter counted)."""
        out, n = {}, 0
        for group in self.param_groups:
            for p in group["params"]:
                out[id(p)] = n
                n += 1
        return out

    def state_dict(self):
        """The layout of ``torch.optim.Adam.state_dict()``: ``state[index] = {step, exp_avg, exp_avg_sq}`` per
        parameter that has been stepped -- a reference ``*.state`` file written by this optimiser resumes under
        torch.optim.Adam and vice versa (codes/models/base_model.py resume_training -> optimizer.load_state_dict)."""
        sd = super().state_dict()
        index = self._param_indices()
        state = {}
        for st in self._flat.values():
            for j, (p, off) in enumerate(zip(st["params"], st["offsets"])):
                if st["pstep"][j] <= 0:
                    continue
                n = p.numel()
                state[index[id(p)]] = dict(step=torch.tensor(float(st["pstep"][j])),
                                           exp_avg=st["m"][off:off + n].view(p.shape).clone(),
                                           exp_avg_sq=st["v"][off:off + n].view(p.shape).clone())
        sd["state"] = state
        return sd

    def load_state_dict(self, state_dict):
        """Accepts this class's own state_dict, a ``torch.optim.Adam`` state_dict (per-parameter moments are copied
        into the flat buffers) and the ``flat`` layout of earlier versions of this class."""
        flat = state_dict.get("flat", {})
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "flat"})
        for gi, group in enumerate(self.param_groups):
            st = self._ensure_flat(gi, group)
            if st is None:
                continue
            if gi in flat:
                st["step"] = int(flat[gi]["step"])
                st["pstep"] = list(flat[gi].get("pstep", [st["step"]] * len(st["params"])))
                st["m"].copy_(flat[gi]["exp_avg"])
                st["v"].copy_(flat[gi]["exp_avg_sq"])
                continue
            for j, (p, off) in enumerate(zip(st["params"], st["offsets"])):
                ps = self.state.get(p)
                if not ps:
                    continue
                n = p.numel()
                st["m"][off:off + n].copy_(ps["exp_avg"].reshape(-1))
                st["v"][off:off + n].copy_(ps["exp_avg_sq"].reshape(-1))
                st["pstep"][j] = int(float(ps["step"]))
            st["step"] = max(st["pstep"]) if st["pstep"] else 0
        self.state.clear()       # the moments live in the flat buffers

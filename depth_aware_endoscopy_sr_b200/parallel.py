"""Data parallelism of the hot path: one process per GPU, frames / images sharded by rank.

* Inference shards by frame and needs no collective: ``shard_frames``.
* Training partitions the image batch; the only exchange step is the gradient all-reduce.  ``FlatDataParallel``
  takes the place of the ``DistributedDataParallel`` wrapper the reference puts around ``netG`` when ``opt['dist']``
  (codes/models/F_model_depthCond.py:32-33): same ``.module`` attribute and call signature, but the gradients of one
  backward are ONE flat fp32 buffer (``Engine.last_flat_grad``, 59 MB at x8) that is all-reduced with a single
  NCCL call over NVLink/NVSwitch before autograd hands the views to ``.grad`` -- no per-parameter hooks, no
  buckets, and parameters the network never uses (``depth-residual14.*``, which make stock DDP raise) are simply
  zero slices of that buffer.
* ``install_ddp(reference_module)`` swaps the ``DistributedDataParallel`` symbol inside the reference's
  ``models.F_model_depthCond`` for this wrapper so ``codes/train.py --launcher pytorch`` runs unmodified.

Loss semantics under data parallelism.  ``mode="ddp"`` (default) is what the reference's own DDP would compute:
every rank forms its local loss, gradients are averaged.  ``mode="global"`` reproduces a single-process run over
the concatenated batch exactly: the dynamic depth-mask loss is a ratio of *batch sums* (mask_loss.py:80-83), so the
33 partial sums are all-reduced before the ratio is formed (``loss_sums_hook``) and the gradients are summed.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn


def world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_frames(n_frames: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    """Frame indices this rank processes: frame f -> rank f mod world (no collective on the data path)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return list(range(rank, n_frames, world_size))


def allreduce_flat_(flat: torch.Tensor, group=None, average: bool = True) -> torch.Tensor:
    """In-place all-reduce of one flat buffer.  NCCL averages inside the collective (ReduceOp.AVG); gloo (the CPU
    tests) sums and divides."""
    rank, ws = world(group)
    if ws == 1:
        return flat
    backend = dist.get_backend(group)
    if average and backend == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(ws)
    return flat


def sync_extra_grads_(params: Iterable[torch.nn.Parameter], group=None, average: bool = True) -> None:
    """All-reduce the gradients of the few parameters that live outside the generator (the 10 weights of the dynamic
    loss) through one small flat buffer."""
    ps = [p for p in params if p.grad is not None]
    if not ps or world(group)[1] == 1:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    allreduce_flat_(flat, group, average)
    off = 0
    for p in ps:
        p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
        off += p.numel()


def loss_sums_hook(group=None):
    """``sums_hook`` for ``loss.training_loss``: all-reduce (sum) the partial sums of the loss so that the ratio of
    batch sums is the global one."""
    def hook(sums: torch.Tensor):
        if world(group)[1] > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return hook


def broadcast_parameters_(module: nn.Module, src: int = 0, group=None) -> None:
    """Every rank starts from rank ``src``'s weights (what DistributedDataParallel does at construction)."""
    if world(group)[1] == 1:
        return
    with torch.no_grad():
        ts = list(module.parameters()) + list(module.buffers())
        for t in ts:
            dist.broadcast(t.data, src=src, group=group)
        # written through .data: advance the version counters like an in-place op would, so that weight caches keyed
        # on them (Engine.pack, the recorded inference graphs) are refreshed on the ranks that received new values
        if ts:
            torch.autograd.graph.increment_version(ts)


class FlatDataParallel(nn.Module):
    """Drop-in for ``DistributedDataParallel(netG, device_ids=[...])`` around a B200 ``DepthNet``."""

    def __init__(self, module: nn.Module, device_ids=None, output_device=None, process_group=None, mode: str = "ddp",
                 broadcast: bool = True, **_ignored):
        super().__init__()
        if mode not in ("ddp", "global"):
            raise ValueError("mode must be 'ddp' or 'global'")
        self.module = module
        self.process_group = process_group
        self.mode = mode
        if broadcast:
            broadcast_parameters_(module, 0, process_group)
        eng = module.engine() if hasattr(module, "engine") else None
        if eng is None:
            raise TypeError("FlatDataParallel wraps the B200 DepthNet (a module with .engine())")
        eng.grad_sync = self._sync

    def _sync(self, flat: torch.Tensor) -> None:
        allreduce_flat_(flat, self.process_group, average=(self.mode == "ddp"))

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def install_ddp(model_module=None):
    """Replace the ``DistributedDataParallel`` name inside the reference's ``models.F_model_depthCond`` (the wrapper
    it applies at line 33) with ``FlatDataParallel``.  Call before ``create_model(opt)``; see INTEGRATION.md."""
    if model_module is None:
        import models.F_model_depthCond as model_module  # the reference package, must be on sys.path
    model_module.DistributedDataParallel = FlatDataParallel
    return model_module

"""Host-side logic of the N > 1 path on CPU: two gloo ranks (no GPU needed).

Covers frame sharding, the single flat-buffer gradient all-reduce behind ``FlatDataParallel`` (installed as the
engine's ``grad_sync``), the small all-reduce of the dynamic-loss weights, the "global batch" loss-sum hook and the
parameter broadcast -- everything of parallel.py that is not a kernel."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeEngine:
    def __init__(self):
        self.grad_sync = None


class _FakeNet(nn.Module):
    """Stands in for DepthNet on CPU: has .engine() and parameters; 'backward' hands a flat buffer to grad_sync."""

    def __init__(self, seed):
        super().__init__()
        torch.manual_seed(seed)
        self.a = nn.Parameter(torch.randn(5, 3))
        self.b = nn.Parameter(torch.randn(7))
        self._e = _FakeEngine()

    def engine(self):
        return self._e

    def forward(self, x):
        return x * 2


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from depth_aware_endoscopy_sr_b200 import parallel as par
        res = {}
        res["frames"] = par.shard_frames(11)
        net = _FakeNet(seed=rank)                     # different weights per rank before wrapping
        ddp = par.FlatDataParallel(net, device_ids=[0])
        res["a_after_broadcast"] = net.a.detach().clone()
        assert ddp.module is net and net.engine().grad_sync is not None
        assert torch.equal(ddp(torch.ones(2)), torch.full((2,), 2.0))
        flat = torch.arange(24, dtype=torch.float32) * (rank + 1)
        net.engine().grad_sync(flat)                  # what Engine._finish_backward does with the flat gradient
        res["flat_avg"] = flat.clone()
        glob = par.FlatDataParallel(net, mode="global", broadcast=False)
        flat = torch.ones(8) * (rank + 1)
        net.engine().grad_sync(flat)
        res["flat_sum"] = flat.clone()
        w = nn.Parameter(torch.zeros(10))
        w.grad = torch.full((10,), float(rank))
        par.sync_extra_grads_([w])
        res["extra"] = w.grad.clone()
        sums = torch.full((36,), float(rank + 1))
        par.loss_sums_hook()(sums)
        res["sums"] = sums.clone()
        q.put((rank, {k: (v.tolist() if torch.is_tensor(v) else v) for k, v in res.items()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_data_parallel_plumbing():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    out = {r: {k: (torch.tensor(v) if k != "frames" else v) for k, v in d.items()} for r, d in out.items()}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0]["frames"] == [0, 2, 4, 6, 8, 10] and out[1]["frames"] == [1, 3, 5, 7, 9]
    assert torch.equal(out[0]["a_after_broadcast"], out[1]["a_after_broadcast"])
    torch.manual_seed(0)
    assert torch.equal(out[1]["a_after_broadcast"], torch.randn(5, 3))            # rank 0's weights won
    exp = torch.arange(24, dtype=torch.float32) * 1.5
    assert torch.allclose(out[0]["flat_avg"], exp) and torch.allclose(out[1]["flat_avg"], exp)
    assert torch.allclose(out[0]["flat_sum"], torch.full((8,), 3.0))
    assert torch.allclose(out[1]["extra"], torch.full((10,), 0.5))
    assert torch.allclose(out[0]["sums"], torch.full((36,), 3.0))


def test_single_process_helpers_are_identity():
    from depth_aware_endoscopy_sr_b200 import parallel as par
    assert par.world() == (0, 1)
    assert par.shard_frames(5) == [0, 1, 2, 3, 4]
    assert par.shard_frames(5, 1, 2) == [1, 3]
    t = torch.ones(4)
    assert par.allreduce_flat_(t) is t and torch.equal(t, torch.ones(4))
    with pytest.raises(TypeError):
        par.FlatDataParallel(nn.Linear(2, 2))

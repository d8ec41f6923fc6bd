"""Two-GPU tests of the data-parallel training path (NCCL; skipped on a single-GPU box; run with
``gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu``).

* mode="global": the step over two half batches on two ranks equals the single-process step over the whole batch --
  the loss partial sums are all-reduced BEFORE the dynamic-loss ratio is formed and the gradients are summed
  (parallel.py); checked on the flat gradient buffer, the 10 loss-weight gradients and the loss value.
* mode="ddp": per-rank losses, averaged gradients (what the reference's DistributedDataParallel computes); after
  three steps both ranks hold bit-identical parameters.
"""
import os
import socket
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(seed=21):
    import depth_aware_endoscopy_sr_b200 as dasr
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return dasr.DepthNet(which_ResBlk_depth=[0, 1, 2], scale=8, nb=6)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import depth_aware_endoscopy_sr_b200 as dasr
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        full = synthetic_inputs(4, 32, 32, scale=8, seed=33, with_gt=True)
        mine = [t[rank * 2:rank * 2 + 2].contiguous().to(dev) for t in full]
        res = {}
        # ---- global mode: one step on the half batch of this rank
        net = _build().to(dev).train()
        step = dasr.TrainStep(net, num_masks=10, lr=1e-3, betas=(0.9, 0.99), distributed=True, mode="global")
        vec = step(*mine)
        torch.cuda.synchronize()
        res["global_loss"] = float(vec[0].item())
        res["global_flat"] = net.engine().last_flat_grad.detach().double().cpu()
        res["global_dw"] = step.dynamic_loss.trainable_weight.grad.detach().double().cpu()
        if rank == 0:       # the single-process step over the concatenated batch (fresh, identical weights)
            ref = _build().to(dev).train()
            rstep = dasr.TrainStep(ref, num_masks=10, lr=1e-3, betas=(0.9, 0.99))
            rvec = rstep(*[t.to(dev) for t in full])
            torch.cuda.synchronize()
            res["single_loss"] = float(rvec[0].item())
            res["single_flat"] = ref.engine().last_flat_grad.detach().double().cpu()
            res["single_dw"] = rstep.dynamic_loss.trainable_weight.grad.detach().double().cpu()
        # ---- ddp mode: three steps, parameters must stay identical on both ranks
        net2 = _build(seed=22 + rank).to(dev).train()       # different init per rank: the wrapper broadcasts rank 0's
        step2 = dasr.TrainStep(net2, num_masks=10, lr=1e-3, betas=(0.9, 0.99), distributed=True, mode="ddp")
        for _ in range(3):
            step2(*mine)
        torch.cuda.synchronize()
        res["ddp_params"] = torch.cat([p.detach().reshape(-1) for p in net2.parameters()]).cpu()
        res["ddp_w"] = step2.dynamic_loss.trainable_weight.detach().cpu()
        q.put((rank, {k: (v.tolist() if torch.is_tensor(v) else v) for k, v in res.items()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_training_matches_single_process_and_ranks_stay_in_sync():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0, r1 = out[0], out[1]
    g0, g1, gs = (torch.tensor(x, dtype=torch.float64) for x in (r0["global_flat"], r1["global_flat"], r0["single_flat"]))
    assert torch.equal(g0, g1)                                         # the all-reduce left the same buffer on both ranks
    rel = ((g0 - gs).norm() / gs.norm()).item()
    print("global mode: |flat grad (2 ranks) - flat grad (1 process)| / |.| = %.3g ; loss %.6f vs %.6f" % (
        rel, r0["global_loss"], r0["single_loss"]))
    assert rel <= 2e-3                                                 # same per-image bf16 arithmetic, other fp32 sum order
    assert abs(r0["global_loss"] - r0["single_loss"]) <= 1e-5 * abs(r0["single_loss"])
    dw, dws = torch.tensor(r0["global_dw"]), torch.tensor(r0["single_dw"])
    assert torch.equal(dw, torch.tensor(r1["global_dw"]))
    assert (dw - dws).abs().max().item() <= 1e-4 * dws.abs().max().item() + 1e-9      # NOT multiplied by the world size
    assert r0["ddp_params"] == r1["ddp_params"] and r0["ddp_w"] == r1["ddp_w"]          # bit-identical after three steps

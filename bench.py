#!/usr/bin/env python
"""Benchmark of the DepthNet hot path (BASELINE.json: HR frames/s, x8 SR 64x64 -> 512x512).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]

A "step" is one forward pass of the generator over one batch of synthetic frames (BASELINE.json configs[1]:
x8 inference, batch 64 per GPU, 3x64x64 LR + 1x64x64 depth + 10 depth masks -> 3x512x512, bf16 tensor-core
math with fp32 accumulation, random-init weights).  For N > 1 the frames are sharded by rank (no collective on
the data path) and the rate is the whole-job aggregate (weak scaling).

Prints ONE JSON line:
  value      frames/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e        frames/s end to end through the frame-level public call ``net.infer_frames(LQ, Depth)`` with pinned HOST
             buffers: the host->device copies of the LR frames and depth maps and the device->host read of the uint8
             BGR frames are inside the timed region.  This is the reference's test path (codes/test.py:
             getDepthMask in the data pipeline -> netG -> util.tensor2img) with all three steps on the device
             (SURVEY 8(f) rows 1-2): 4 input planes up, one byte per output sample down
  e2e_fp32_tensor  the module-level call ``net(LQ, Depth, DepthMaskList)`` -> fp32 SR tensor (the drop-in nn.Module
             boundary; 14 input planes up, 4 bytes per output sample down), timed the same way; also stored as
             e2e["fp32_tensor"].  Its 201 MB per step and GPU saturate the host's PCIe / memory fabric when several
             GPUs share a host (0.33 efficiency at 8 GPUs, 92 GB/s of frames), which is why the frame-level call is
             the end-to-end figure; e2e_frames repeats e2e under its round-1 name
  stream_1080p  BASELINE.json configs[4]: 135x240 LR frames -> 1080x1920, frames sharded over the ranks, device-resident
             and end to end (uint8 frames)
  roofline   the dominant kernel family of the step, measured live with CUDA events around every launch in two
             instrumented passes that bracket the timed region (3 steps right before it, 3 right after: the board
             heats up over the timed steps and reaches its power cap) -- achieved algorithmic TFLOP/s or GB/s of the
             average launch against MEASURED_PEAKS.json; ms_per_launch_before_after gives both ends
  train      BASELINE.json configs[2] next to the headline: x8 training step (forward + loss + backward + flat
             NCCL gradient all-reduce + Adam), batch 16 per GPU, images/s over the whole job; dp_check = the parameter
             checksum after the timed steps is identical on every rank
  cpu_baseline  the CPU oracle port of the reference's forward (literal form: 256-channel style-map convs),
             timed on this box's host cores on a bounded sample (rank 0, N == 1 only); the same leg checks the CUDA
             output of two frames of the timed batch against that oracle (``parity``) and times the same oracle
             on the GPU with torch's own kernels (``gpu_eager_baseline``: cuDNN / cuBLAS, fp32 and bf16 autocast)

``--impl reference`` times the reference's own CPU algorithm (the oracle port -- the reference is a Python
code base that cannot travel to the GPU box) with all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

# stdout carries exactly ONE JSON line.  NCCL prints its version banner (NCCL_DEBUG=VERSION|WARN|INFO) with a plain
# printf to fd 1, so fd 1 is pointed at stderr for the whole run and the JSON line goes to a duplicate of the
# original stdout.
_JSON_OUT = None


def _claim_stdout():
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCALE = 8
LR = 64
WHICH = tuple(range(14))
METRIC = "hr_frames_per_sec_x8_sr_64to512"
UNIT = "frames/s"


REF_SAMPLE_B = 2       # frames per step of the bounded CPU sample (reference arm AND cpu_baseline leg)


def _config(B, numa=None, n_sets=3):
    """The ``config`` object of BOTH arms (the reference arm times a bounded sample of this workload)."""
    return {"workload": "depthNet_SEAN_depthMask x8 inference, batch %d per GPU, synthetic 3x64x64 LR + 1x64x64 depth + "
                        "10 masks -> 3x512x512, random-init weights" % B,
            "batch_per_gpu": B, "lr": [LR, LR], "scale": SCALE, "sharding": "frames by rank, no collective",
            "l2": "no flush: one step streams ~3.5 GB of activations (>> 126 MB L2) and the input batch rotates over "
                  "%d sets" % n_sets}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tc_burst=float(d["bf16_tflops"]),
                    tc_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md recipe): one
    `nvidia-smi -lms 50` child process streams CSV rows while the benchmark runs; stop() ends exactly that PID."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def _reader(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.strip().split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._reader, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self):
        """Number of rows so far (rows before the mark belong to warm-up)."""
        return len(self.rows)

    def stop(self, since=0):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            if self.thread is not None:
                self.thread.join(timeout=2)
        rows = self.rows[since:] or self.rows
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        reasons = []
        for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        mx = max([float(r[1]) for r in rows if r[1].replace(".", "").isdigit()] or [0.0])
        pw = max([float(r[2]) for r in rows if r[2].replace(".", "").isdigit()] or [0.0])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(rows), "power_w_max": pw or None}


def _make_state(seed=0):
    """The reference's own random init (our DepthNet consumes the RNG like the reference constructor)."""
    import torch
    import depth_aware_endoscopy_sr_b200 as dasr
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=list(WHICH), in_nc=3, out_nc=3, nf=64, nb=16, scale=SCALE,
                            input_para=10, depth_latent_ch=256, depthRangeNum=10, norm_type="weight_norm",
                            use_trainable_params=True, norm_gamma=0, norm_beta=0)
    return net


def cpu_reference_rate(seconds_budget=20.0, batch=REF_SAMPLE_B, warmup=1, min_iters=3, check=None):
    """frames/s of the CPU oracle port (the reference's algorithm, fp32, all host threads) on ``batch`` frames.
    check = (lq, depth, masks, sr_cuda) on the host: the oracle also runs on THESE frames and the max-abs difference
    to the CUDA output is returned (the oracle as the checker)."""
    import torch
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    from oracle import depthnet_oracle as oracle     # the checker, allowed here as the cpu_baseline leg
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    net = _make_state(0)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    lq, depth, masks = synthetic_inputs(batch, LR, LR, scale=SCALE, seed=0)
    times = []
    parity = None
    with torch.no_grad():
        if check is not None:
            ref = oracle.depthnet_forward(sd, check[0], check[1], check[2], scale=SCALE, which=WHICH)
            parity = {"max_abs_vs_oracle": float((check[3] - ref).abs().max()), "frames": int(ref.shape[0]),
                      "tolerance": 1e-2, "what": "SR of the first frames of the timed batch, CUDA vs the CPU oracle"}
        for _ in range(warmup):
            oracle.depthnet_forward(sd, lq, depth, masks, scale=SCALE, which=WHICH)
        t_all = time.time()
        while len(times) < min_iters or (time.time() - t_all) < seconds_budget:
            t0 = time.time()
            oracle.depthnet_forward(sd, lq, depth, masks, scale=SCALE, which=WHICH)
            times.append(time.time() - t0)
            if len(times) >= 200:
                break
    times.sort()
    med = times[len(times) // 2]
    return batch / med, len(times), torch.get_num_threads(), parity


def gpu_eager_rates(dev, B=64, Bt=16):
    """The reference's own formulation executed with torch's GPU kernels (cuDNN / cuBLAS on sm_100) -- the practical
    existing-kernel bar next to the CPU figure.  Same oracle port as the CPU leg: NCHW, 256-channel style-map convs,
    10 python-level mask loops; fp32 (TF32 off) and bf16 autocast + channels_last; inference B frames and one training
    step (forward + loss + backward, no optimizer) at Bt images."""
    import torch
    from oracle import depthnet_oracle as oracle
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    out = {}
    net = _make_state(0)
    sd = {k: v.detach().to(dev) for k, v in net.state_dict().items()}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for mode in ("fp32", "bf16_autocast"):
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            lq, depth, masks = [t.to(dev) for t in synthetic_inputs(B, LR, LR, scale=SCALE, seed=0)]
            if mode != "fp32":
                lq, depth, masks = [t.contiguous(memory_format=torch.channels_last) for t in (lq, depth, masks)]
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode != "fp32" else torch.autocast("cuda", enabled=False)

            def fwd():
                with torch.no_grad(), ctx:
                    return oracle.depthnet_forward(sd, lq, depth, masks, scale=SCALE, which=WHICH)

            for _ in range(3):
                fwd()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5
            e0.record()
            for _ in range(n):
                fwd()
            e1.record()
            torch.cuda.synchronize()
            out["infer_%s_frames_per_s" % mode] = B * n / (e0.elapsed_time(e1) * 1e-3)
            # training step (forward + loss + backward)
            sdt = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            wd = torch.ones(10, device=dev, requires_grad=True)
            lq2, d2, m2, gt2 = [t.to(dev) for t in synthetic_inputs(Bt, LR, LR, scale=SCALE, seed=1, with_gt=True)]

            def step():
                for v in sdt.values():
                    v.grad = None
                with ctx:
                    sr = oracle.depthnet_forward(sdt, lq2, d2, m2, scale=SCALE, which=WHICH)
                total, *_ = oracle.training_loss(sr.float(), gt2, m2, wd)
                total.backward()

            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                step()
            e1.record()
            torch.cuda.synchronize()
            out["train_%s_images_per_s" % mode] = Bt * 3 / (e0.elapsed_time(e1) * 1e-3)
            del sdt
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = tf32
    out["what"] = ("oracle port of the reference executed with torch %s GPU kernels (cuDNN benchmark on): inference B=%d, "
                   "training step (fwd + loss + bwd) B=%d" % (torch.__version__, B, Bt))
    return out


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    from oracle import depthnet_oracle as oracle
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core the box has
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        torch.set_num_threads(os.cpu_count() or 1)
    net = _make_state(0)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    sample_b = REF_SAMPLE_B   # frames per step of the bounded sample (the workload is args.batch frames per step)
    lq, depth, masks = synthetic_inputs(sample_b, LR, LR, scale=SCALE, seed=0)
    with torch.no_grad():
        for _ in range(max(W, 1)):
            oracle.depthnet_forward(sd, lq, depth, masks, scale=SCALE, which=WHICH)
        t0 = time.time()
        for _ in range(K):
            oracle.depthnet_forward(sd, lq, depth, masks, scale=SCALE, which=WHICH)
        dt = time.time() - t0
    fps = sample_b * K / dt
    cores = torch.get_num_threads()
    sample = "%d steps of %d frames (x8, 64x64 LR) out of the 64-frame batch" % (K, sample_b)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(args.batch),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def _bind_to_gpu_numa_node(local):
    """Pin this rank (and therefore its pinned host buffers, first-touch) to the CPUs of the NUMA node its GPU hangs
    off, so that the e2e host<->device copies of 8 ranks do not cross sockets.  Returns a short description."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        with open(base + "/local_cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return "cpus %s of %s" % (cpulist, bus)
        return "unchanged (%s)" % cpulist
    except Exception as e:       # best effort: containers may hide sysfs
        return "unavailable (%s)" % type(e).__name__


def run_b200(args):
    import torch
    import torch.distributed as dist
    from depth_aware_endoscopy_sr_b200 import _lib
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = _bind_to_gpu_numa_node(local) if world > 1 else "single rank"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().dasr_check_device())

    K, W, B = args.steps, max(args.warmup, 3), args.batch
    net = _make_state(0).to(dev).eval()
    eng = net.engine()

    # a few different input batches (rank-specific seeds), resident in HBM and mirrored in pinned host memory
    n_sets = 3
    host_sets, dev_sets = [], []
    for i in range(n_sets):
        lq, depth, masks = synthetic_inputs(B, LR, LR, scale=SCALE, seed=1000 * rank + i)
        host_sets.append(tuple(t.pin_memory() for t in (lq, depth, masks)))
        dev_sets.append(tuple(t.to(dev) for t in (lq, depth, masks)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # ------------------------------------------------ device-resident throughput
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for i in range(W):
            net(*dev_sets[i % n_sets])
        # per-kernel instrumented pass, first half (the second follows the timed region)
        roof_recs = [roofline_records(net, dev_sets)] if rank == 0 else []
        barrier()
        clk_mark = sampler.mark()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            net(*dev_sets[i % n_sets])
        e1.record()
        barrier()
        launches = _lib.launch_count() - n0
        ms_total = e0.elapsed_time(e1)

        # ------------------------------------------------ per-kernel instrumented pass, second half (the two halves
        # bracket the timed region: its thermal / power state at the start and at the end) and the frames the
        # cpu_baseline leg will check against the oracle (before any training step changes the weights)
        roof = None
        if rank == 0:
            roof_recs.append(roofline_records(net, dev_sets))
            roof = roofline_pass(roof_recs)
        nchk = min(2, B)
        sr_chk = net(*[t[:nchk].contiguous() for t in dev_sets[0]]).cpu() if rank == 0 else None

        # ------------------------------------------------ end to end through the public call, host buffers
        # Every step copies ITS inputs from pinned host memory and reads ITS SR frames back to pinned host memory;
        # the copies run on their own streams so that step i's read-back overlaps step i+1's kernels
        # (double-buffered device inputs / host outputs).  The call a user makes is still net(LQ, Depth, Masks).
        out_host = [torch.empty(B, 3, SCALE * LR, SCALE * LR, dtype=torch.float32).pin_memory() for _ in range(2)]
        dev_in = [tuple(torch.empty_like(t, device=dev) for t in host_sets[0]) for _ in range(2)]
        h2d = sum(t.numel() * t.element_size() for t in host_sets[0])
        d2h = out_host[0].numel() * out_host[0].element_size()
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        held = [None, None]
        # The result of a step stays referenced (per slot) until the main stream has waited for its read-back, and is
        # released there: its memory returns to the main stream's pool and is reused deterministically.  With
        # Tensor.record_stream instead, the caching allocator sometimes has to cudaMalloc inside the loop (tools/
        # e2e_jitter.py: up to 33 calls and 33 ms host stalls in 50 steps), which halved this figure in one run of ten.

        def e2e_step(i):
            slot = i % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_done[slot])          # the step that last read this input slot has finished
                for d, h in zip(dev_in[slot], host_sets[i % n_sets]):
                    d.copy_(h, non_blocking=True)
                ev_in[slot].record(s_in)
            cur.wait_event(ev_in[slot])
            cur.wait_event(ev_out[slot])                # the read-back of the result this slot still holds is done
            held[slot] = None
            sr = net(*dev_in[slot])
            ev_done[slot].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[slot])
                out_host[slot].copy_(sr, non_blocking=True)
                ev_out[slot].record(s_out)
            held[slot] = sr

        def e2e_join():
            cur.wait_stream(s_in)
            cur.wait_stream(s_out)

        for i in range(3):
            e2e_step(i)
        e2e_join()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(K):
            e2e_step(i)
        e2e_join()
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)

        # ------------------------------------------------ frame-level public call: net.infer_frames(LQ, Depth)
        # getDepthMask and tensor2img of the reference run on the device: only the LR frame and
        # its depth map go up (4 planes instead of 14), uint8 BGR frames come back (a quarter of the fp32 bytes)
        out_u8 = [torch.empty(B, SCALE * LR, SCALE * LR, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
        h2d_frames = sum(t.numel() * t.element_size() for t in host_sets[0][:2])

        def e2e_u8_step(i):
            slot = i % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_done[slot])
                for d, h in zip(dev_in[slot][:2], host_sets[i % n_sets][:2]):
                    d.copy_(h, non_blocking=True)
                ev_in[slot].record(s_in)
            cur.wait_event(ev_in[slot])
            cur.wait_event(ev_out[slot])
            held[slot] = None
            img = net.infer_frames(dev_in[slot][0], dev_in[slot][1])
            ev_done[slot].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[slot])
                out_u8[slot].copy_(img, non_blocking=True)
                ev_out[slot].record(s_out)
            held[slot] = img

        held[0] = held[1] = None
        for i in range(3):
            e2e_u8_step(i)
        e2e_join()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(K):
            e2e_u8_step(i)
        e2e_join()
        g1.record()
        barrier()
        ms_e2e_u8 = g0.elapsed_time(g1)

        # ------------------------------------------------ BASELINE.json configs[4]: 1080p stream, frames sharded by rank
        stream = stream_pass(args, net, dev, rank, world, barrier, s_in, s_out) if args.stream_steps > 0 else None
        clocks = sampler.stop(clk_mark) if rank == 0 else None      # samples span the timed regions

    train = train_pass(args, net, dev, rank, world, barrier) if args.train_steps > 0 else None

    t = torch.tensor([ms_total, ms_e2e, ms_e2e_u8], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_e2e_u8 = t.tolist()
    value = B * world * K / (ms_total * 1e-3)
    e2e = B * world * K / (ms_e2e * 1e-3)

    cpu = None
    parity = None
    eager = None
    if rank == 0:
        if world == 1 and not args.no_cpu:
            check = tuple(t[:nchk] for t in host_sets[0]) + (sr_chk,)
            fps, iters, cores, parity = cpu_reference_rate(seconds_budget=args.cpu_seconds, check=check)
            cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d forwards of %d frames (x8, 64x64 LR -> 512x512) of the %d-frame batch, median" % (
                       iters, REF_SAMPLE_B, B)}
            if not args.no_eager:
                try:
                    eager = gpu_eager_rates(dev, B=B, Bt=args.train_batch)
                except Exception as e:      # a baseline must not take the benchmark down
                    eager = {"error": "%s: %s" % (type(e).__name__, e)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    frames = {"value": B * world * K / (ms_e2e_u8 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_frames,
              "d2h_bytes_per_step": out_u8[0].numel(), "ms_per_step": ms_e2e_u8 / K,
              "call": "net.infer_frames(LQ, Depth) -> uint8 BGR frames (depth masks and tensor2img on the device)"}
    fp32_tensor = {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": ms_e2e / K, "call": "net(LQ, Depth, DepthMaskList) -> fp32 SR tensor"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": _config(B),
            "e2e": dict(frames, fp32_tensor=fp32_tensor),
            "e2e_fp32_tensor": fp32_tensor, "e2e_frames": frames, "stream_1080p": stream, "host_affinity": numa,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity": parity,
            "gpu_eager_baseline": eager, "train": train}
    _emit(line)


def stream_pass(args, net, dev, rank, world, barrier, s_in, s_out):
    """1080p-HR endoscopy stream: 135x240 LR frames + depth -> 1080x1920, ``--stream-frames`` frames per call and
    rank (frame f of the stream -> rank f mod N, parallel.shard_frames; no collective).  Device-resident rate and the
    end-to-end rate through net.infer_frames with pinned host buffers (LR + depth up, uint8 frames down)."""
    import torch
    import torch.distributed as dist
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    F_, Ks = args.stream_frames, args.stream_steps
    h, w = 135, 240
    sets_h = [tuple(t.pin_memory() for t in synthetic_inputs(F_, h, w, scale=SCALE, seed=7000 + 10 * rank + i)[:2])
              for i in range(2)]
    sets_d = [tuple(t.to(dev) for t in hs) for hs in sets_h]
    out_h = [torch.empty(F_, SCALE * h, SCALE * w, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
    dev_in = [tuple(torch.empty_like(t, device=dev) for t in sets_h[0]) for _ in range(2)]
    cur = torch.cuda.current_stream()
    for i in range(3):
        net.infer_frames(*sets_d[i % 2])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(Ks):
        net.infer_frames(*sets_d[i % 2])
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    held = [None, None]       # results stay referenced until their read-back is done (see the e2e leg of run_b200)

    def step(i):
        slot = i % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_done[slot])
            for d, hh in zip(dev_in[slot], sets_h[i % 2]):
                d.copy_(hh, non_blocking=True)
            ev_in[slot].record(s_in)
        cur.wait_event(ev_in[slot])
        cur.wait_event(ev_out[slot])
        held[slot] = None
        img = net.infer_frames(*dev_in[slot])
        ev_done[slot].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[slot])
            out_h[slot].copy_(img, non_blocking=True)
            ev_out[slot].record(s_out)
        held[slot] = img

    for i in range(3):
        step(i)
    cur.wait_stream(s_in)
    cur.wait_stream(s_out)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(Ks):
        step(i)
    cur.wait_stream(s_in)
    cur.wait_stream(s_out)
    f1.record()
    barrier()
    t = torch.tensor([ms_dev, f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()
    h2d = sum(x.numel() * x.element_size() for x in sets_h[0])
    return {"metric": "hr_frames_per_sec_x8_sr_1080p_stream", "unit": UNIT, "value": F_ * world * Ks / (ms_dev * 1e-3),
            "ms_per_call": ms_dev / Ks, "frames_per_call_per_gpu": F_, "steps": Ks, "lr": [h, w], "hr": [SCALE * h, SCALE * w],
            "sharding": "frame f -> rank f mod %d, no collective" % world,
            "e2e": {"value": F_ * world * Ks / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_call": ms_e2e / Ks,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_h[0].numel(),
                    "call": "net.infer_frames(LQ, Depth) -> uint8 BGR frames"}}


def train_pass(args, net, dev, rank, world, barrier):
    """BASELINE.json configs[2]: x8 training step (L1 + dynamic depth-mask loss, Adam), batch 16 per GPU, data
    parallel over the ranks with ONE flat NCCL gradient all-reduce per step.  Reported next to the headline metric
    as images/s (whole job), timed like the inference leg (CUDA events, barrier on both sides, max over ranks)."""
    import torch
    import torch.distributed as dist
    import depth_aware_endoscopy_sr_b200 as dasr
    from depth_aware_endoscopy_sr_b200 import _lib
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    Bt, Kt = args.train_batch, args.train_steps
    net.train()
    step = dasr.TrainStep(net, num_masks=10, lr=1e-3, betas=(0.9, 0.99), distributed=world > 1, mode="ddp",
                          graph=not args.no_graph)
    sets = [tuple(t.to(dev) for t in synthetic_inputs(Bt, LR, LR, scale=SCALE, seed=5000 + 100 * rank + i, with_gt=True))
            for i in range(2)]
    for i in range(4):
        step(*sets[i % 2])
    barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(Kt):
        vec = step(*sets[i % 2])
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    loss = float(vec[0].item())
    # data-parallel consistency: after the same number of steps every rank must hold the SAME parameters (they start
    # from rank 0's weights and apply the same averaged gradients); a silent divergence would show up here
    chk = torch.stack([torch.cat([p.detach().double().reshape(-1) for p in net.parameters()]).sum(),
                       torch.cat([p.detach().double().abs().reshape(-1) for p in net.parameters()]).sum()])
    dp_check = None
    if world > 1:
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        dp_check = {"param_checksum_identical_on_all_ranks": bool(all(torch.equal(allc[0], c) for c in allc)),
                    "checksum": [float(v) for v in allc[0].tolist()], "ranks": world, "after_steps": Kt + 4}
    net.eval()
    nparam = sum(p.numel() for p in net.parameters())
    return {"metric": "train_images_per_sec_x8_64to512", "value": Bt * world * Kt / (ms * 1e-3), "unit": "images/s",
            "ms_per_step": ms / Kt, "steps": Kt, "batch_per_gpu": Bt, "global_batch": Bt * world, "scaling": "weak",
            "loss": "L1 + dynamic depth-mask (SmoothL1), Adam lr 1e-3 betas (0.9, 0.99)", "last_loss": loss,
            "gradient_allreduce": None if world == 1 else "one NCCL all-reduce (AVG) of the flat fp32 buffer, %.1f MB"
                                                             % (nparam * 4 / 1e6),
            "cuda_graph": not args.no_graph, "dp_check": dp_check,
            "gpu_launches": int(launches)}


def _ncu_traffic(kernel_family):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
    `ncu --set full` capture (profiles/r02_sean_pair_v3_ncu_full_summary.csv, first captured launch), or None."""
    path = {"conv3x3_128to128_sean": os.path.join(ROOT, "profiles", "r02_sean_pair_v3_ncu_full_summary.csv")}.get(kernel_family)
    if path is None or not os.path.exists(path):
        return None
    import csv
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    seen = 0
    with open(path) as f:
        for row in csv.reader(f):
            if row and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and len(row) > 2:
                tot += float(row[2]) * scale.get(row[1], 1.0)
                seen += 1
    return tot if seen == 2 else None


def roofline_records(net, dev_sets, reps=3):
    """Instrumented forwards: CUDA events around every kernel launch of the engine (on the launching stream); returns
    the per-launch records of ``reps`` steps."""
    import torch
    eng = net.engine()
    with torch.no_grad():
        eng.profile = []
        for i in range(2):
            net(*dev_sets[i % len(dev_sets)])
        torch.cuda.synchronize()
        eng.profile = []
        for i in range(reps):
            net(*dev_sets[i % len(dev_sets)])
        torch.cuda.synchronize()
        recs = eng.profile
        eng.profile = None
    return recs, reps


def roofline_pass(passes):
    """``passes``: roofline_records taken right BEFORE and right AFTER the timed region (the board heats up over the 50
    timed steps and runs into its power cap; one pass alone sees only one end of that).  Launch durations are grouped
    by kernel family and averaged over both passes; reports the family that takes the largest share of the step."""
    peaks = _peaks()
    fam = {}
    reps = 0
    for recs, n in passes:
        reps += n
        for r in recs:
            f = fam.setdefault(r["family"], dict(ms=0.0, n=0, flops=0.0, bytes=0.0, bound=r["bound"]))
            f["ms"] += r["e0"].elapsed_time(r["e1"])
            f["n"] += 1
            f["flops"] += r["flops"]
            f["bytes"] += r["bytes"]
    total = sum(f["ms"] for f in fam.values())
    top = max(fam.items(), key=lambda kv: kv[1]["ms"])
    name, f = top
    table = {k: {"share": v["ms"] / total, "ms_per_launch": v["ms"] / v["n"], "launches_per_step": v["n"] // reps,
                 "achieved": (v["flops"] / v["ms"] / 1e9) if v["bound"] == "tensor" else (v["bytes"] / v["ms"] / 1e6),
                 "unit": "TFLOP/s" if v["bound"] == "tensor" else "GB/s"}
             for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    if f["bound"] == "tensor":
        ach, peak, unit = f["flops"] / f["ms"] / 1e9, peaks["tc_sustained"], "TFLOP/s"
    else:
        ach, peak, unit = f["bytes"] / f["ms"] / 1e6, peaks["hbm"], "GB/s"
    per_pass = []
    for recs, n in passes:
        ms = sum(r["e0"].elapsed_time(r["e1"]) for r in recs if r["family"] == name)
        cnt = sum(1 for r in recs if r["family"] == name)
        per_pass.append(ms / max(cnt, 1))
    return {"kernel": name, "bound": "tensor" if f["bound"] == "tensor" else "hbm", "achieved": ach, "peak": peak,
            "unit": unit, "frac": ach / peak, "traffic": _ncu_traffic(name), "peak_source": peaks["src"],
            "ms_per_launch": f["ms"] / f["n"], "ms_per_launch_before_after": per_pass,
            "share_of_step": f["ms"] / total, "families": table}


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step (BASELINE configs[1]: 64)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--train-steps", type=int, default=50, help="timed steps of the training leg (0 = skip it)")
    ap.add_argument("--stream-steps", type=int, default=20, help="timed calls of the 1080p stream leg (0 = skip it)")
    ap.add_argument("--stream-frames", type=int, default=8, help="1080p frames per call and GPU")
    ap.add_argument("--no-eager", action="store_true", help="skip the gpu_eager_baseline leg (torch kernels on the GPU)")
    ap.add_argument("--no-graph", action="store_true", help="issue the training step kernel by kernel (no CUDA graph)")
    ap.add_argument("--train-batch", type=int, default=16, help="images per GPU per training step (BASELINE configs[2])")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py - photometric loss fwd+bwd throughput (Mpix/s of target pixels).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl ours|reference]

A "step" is one fwd+bwd of the loss (`Losses.forward` + `sum(loss).backward()`,
trainer.py:312,264) over one batch of synthetic KITTI-shaped frames.  See
DESIGN.md "Measurement" for the definition of every key of the JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "photometric loss fwd+bwd Mpix/s at 192x640 x3 frames"
# N > 1: the network-gradient all-reduce runs beside the loss; its CTAs need whole SMs, so the collective is capped at
# NCCL_CTAS CTAs and the persistent grid of the loss is sized for the remaining SMs (B200: 148)
NCCL_CTAS = int(os.environ.get("PLB_NCCL_CTAS", "16"))
SM_COUNT_FOR_LOSS = 148 - NCCL_CTAS
UNIT = "Mpix/s"

# name -> (B per GPU, H, W, n_src, n_scales, loss variant)
WORKLOADS = {
    "c1": dict(B=4, H=192, W=640, n_src=2, n_scales=1, variant="live"),
    "c2": dict(B=12, H=192, W=640, n_src=2, n_scales=4, variant="live"),
    "c3": dict(B=8, H=320, W=1024, n_src=3, n_scales=4, variant="live"),
    "c5": dict(B=64, H=192, W=640, n_src=2, n_scales=4, variant="live"),
    # BASELINE.json configs[4] as worded: 4 scales + EDGE-AWARE smoothness (a term the reference does not have: its own
    # smoothness is the 2nd-order one of c5) - the live photometric composition + plb_edge_smooth_loss on the target's pyramid
    "c5e": dict(B=64, H=192, W=640, n_src=2, n_scales=4, variant="live_edge"),
    "headline": dict(B=12, H=192, W=640, n_src=2, n_scales=1, variant="dir0"),
    # the same single-direction kernel at the per-GPU batch of BASELINE.json configs[4] (fixed launch costs amortised)
    "headline64": dict(B=64, H=192, W=640, n_src=2, n_scales=1, variant="dir0"),
    # BASELINE.json configs[1] read literally: the dormant SSIM + per-pixel min-reprojection + automask
    # composition (losses.py:12-84,94-96,154-162) at 4 scales, one direction
    "c2min": dict(B=12, H=192, W=640, n_src=2, n_scales=4, variant="min"),
    # ... with every photometric map clamped at mean + 0.5 std of the whole map, as the reference's
    # compute_photometric_loss does (losses.py:79-82): a grid-wide dependency, i.e. a statistics pass first
    "c2minclip": dict(B=12, H=192, W=640, n_src=2, n_scales=4, variant="minclip"),
}


def workload_name(wl, cfg):
    return "%s: %dx%d batch %d/GPU, 1 target + %d source frames, %d-scale %s loss fwd+bwd" % (
        wl, cfg["H"], cfg["W"], cfg["B"], cfg["n_src"], cfg["n_scales"],
        {"live_edge": "reference-live photometric term (mode='min' = mean over sources, 2 directions, L1) + edge-aware "
                      "1st-order smoothness of the target's disparity pyramid (monodepth2 form; not in the reference)",
         "live": "reference-live Losses.forward (mode='min', which the reference computes as a MEAN over sources, losses.py:226-228; 2 directions, L1 + 2nd-order smoothness)",
         "dir0": "single-direction L1",
         "min": "single-direction SSIM+L1 min-reprojection + automask (dormant path)",
         "minclip": "single-direction SSIM+L1 min-reprojection + automask with the mean + 0.5 std clip of every map (dormant path, "
                    "losses.py:79-82)"}[cfg["variant"]])


def algorithmic_bytes_per_px(cfg):
    """every input byte read once per pass (fwd, bwd), every
    gradient byte written once; all scales of a direction fused into one pass."""
    pyr = 4.0 * sum(0.25 ** s for s in range(cfg["n_scales"]))
    def direction(n_src):
        reads = 12.0 + 12.0 * n_src + pyr
        return 2.0 * reads + pyr
    total = direction(cfg["n_src"])
    if cfg["variant"] in ("live", "live_edge"):
        total += direction(1)
    return total   # "min": same streams as one direction (the automask re-reads the sources it already holds)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML from a thread (a sample every ~2 ms: the
    timed region of a short run is a few tens of milliseconds, less than one `nvidia-smi -lms` period); falls back
    to an `nvidia-smi` child process when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.samples, self.stop_flag, self.t = None, [], False, None

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((float(sm), [n for n, b in bits.items() if r & b]))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map torch's index through CUDA_VISIBLE_DEVICES when it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.nvml = nv
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            sm = sorted(s for s, _ in self.samples)
            reasons = sorted({r for _, rs in self.samples for r in rs})
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference's torch path, host cores
# --------------------------------------------------------------------------------------
def oracle_step(cfg, inp):
    """One fwd+bwd of the reference's torch op sequence (oracle/restated.py, the restatement pinned to the unmodified
    reference by tests/golden) on whatever device `inp` lives on."""
    from oracle import restated as O
    disp = [[d.detach().clone().requires_grad_(True) for d in fr] for fr in inp["disparity"]]
    poses = inp["poses"].detach().clone().requires_grad_(True)
    if cfg["variant"] == "live":
        loss = O.losses_forward(inp["tgt"], inp["ref_imgs"], disp, poses, inp["intrinsics"])
        sum(loss).backward()
        return loss[0].detach() + loss[1].detach()
    depths = O.disp_to_depth(disp)
    if cfg["variant"] == "live_edge":
        loss = O.reprojection_loss(inp["tgt"], inp["ref_imgs"], depths, poses, inp["intrinsics"]) + \
            O.edge_aware_smooth_loss(disp[0], inp["tgt"])
        loss.backward()
        return loss.detach()
    if cfg["variant"] in ("min", "minclip"):
        loss = O.min_reprojection_loss(inp["tgt"], inp["ref_imgs"], depths[0], poses, inp["intrinsics"],
                                       clip_loss=0.5 if cfg["variant"] == "minclip" else None)
    else:
        loss = O.reprojection_loss(inp["tgt"], inp["ref_imgs"], depths[:1], poses, inp["intrinsics"])
    loss.backward()
    return loss.detach()


CPU_ARM_MAX_BATCH = 16


def cpu_reference_run(cfg, steps, warmup, budget_s=25.0):
    """Times the oracle port of the reference's `Losses.forward` + backward (same torch op sequence) on all host
    cores, at the WORKLOAD'S OWN batch (capped at 16 images per step so that a step stays within seconds - the cap
    only bites on c5 / headline64 - and stated in `sample`); same H, W, sources, scales, same synthetic frames."""
    from plb200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = min(CPU_ARM_MAX_BATCH, cfg["B"])
    n_src = cfg["n_src"]
    inp = synth.make_photo_inputs(Bs, cfg["H"], cfg["W"], n_src=n_src, n_scales=cfg["n_scales"], seed=1234,
                                  n_depth_frames=2 if cfg["variant"] in ("live", "live_edge") else 1)
    t_start = time.perf_counter()
    n_warm = max(1, min(warmup, 2))
    for _ in range(n_warm):
        float(oracle_step(cfg, inp))
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        float(oracle_step(cfg, inp))
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break
    mean = sum(times) / len(times)
    mpix = Bs * cfg["H"] * cfg["W"] / 1e6
    return {"value": mpix / mean, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "oracle port of Losses.forward+backward (torch CPU, %d threads), batch %d (workload batch %d) of "
                      "the workload's %dx%d / %d sources / %d scales, %d warm-up + mean of %d steps" % (
                          cores, Bs, cfg["B"], cfg["H"], cfg["W"], n_src, cfg["n_scales"], n_warm, len(times)),
            "ms_per_step": mean * 1e3, "steps": len(times), "warmup": n_warm, "batch": Bs}


def gpu_eager_run(cfg, gpu_sets, dev, iters=10, warmup=3):
    """The real incumbent: the reference's stock torch-eager op
    sequence ON THE SAME B200, same config, same device-resident synthetic frames, CUDA-event timed.  A LIBRARY
    baseline (ATen sm_100 kernels), not our code: `oracle/restated.py` on CUDA tensors is that op sequence, batch
    agnostic where the reference hard-codes batch 4 (geometry/transform.py:110)."""
    st = torch.cuda.current_stream()
    for i in range(warmup):
        oracle_step(cfg, gpu_sets[i % len(gpu_sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(iters):
        oracle_step(cfg, gpu_sets[i % len(gpu_sets)])
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    px = cfg["B"] * cfg["H"] * cfg["W"]
    return {"value": px / 1e6 / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": iters, "warmup": warmup,
            "kind": "torch-eager CUDA (stock ATen kernels, the reference's op sequence) on the same GPU, same config "
                    "and device-resident inputs; library baseline, timed with CUDA events"}


def run_reference_arm(args, cfg):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    r = cpu_reference_run(cfg, args.steps, args.warmup, budget_s=120.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, cfg), "cpu_arm_batch": r["batch"],
                       "note": "CPU arm = oracle port of the reference's torch path (the Python reference cannot travel "
                               "to the GPU box); per-pixel throughput at batch %d of the workload's %d" % (r["batch"], cfg["B"])},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def make_sets(cfg, n_sets, seed, dev):
    from plb200 import synth
    sets = []
    for k in range(n_sets):
        inp = synth.make_photo_inputs(cfg["B"], cfg["H"], cfg["W"], n_src=cfg["n_src"], n_scales=cfg["n_scales"],
                                      seed=seed + 17 * k, n_depth_frames=2 if cfg["variant"] in ("live", "live_edge") else 1)
        sets.append(inp)
    return sets


def set_bytes(inp):
    n = inp["tgt"].numel() * 4 + sum(r.numel() * 4 for r in inp["ref_imgs"]) + inp["poses"].numel() * 4
    n += inp["intrinsics"].numel() * 8 + sum(d.numel() * 4 for fr in inp["disparity"] for d in fr)
    return n


def make_leaves(g):
    return [[d.detach().requires_grad_(True) for d in fr] for fr in g["disparity"]], g["poses"].detach().requires_grad_(True)


def step_fn(criterion, g, cfg, leaves=None):
    """One fwd+bwd through the public API; returns the loss tensors and the grads.  `leaves`: reuse these
    gradient-requiring views of the inputs (a trainer's disparities are network outputs - it does not create leaf
    tensors every step) instead of making new ones."""
    if leaves is None:
        disp, poses = make_leaves(g)
    else:
        disp, poses = leaves
        poses.grad = None
        for fr in disp:
            for d in fr:
                d.grad = None
    if cfg["variant"] in ("live", "live_edge"):
        # live_edge: the criterion is Losses(smoothness="edge") - the edge-aware term inside the same fused call
        loss = criterion.forward(g["tgt"], g["ref_imgs"], disp, poses, g["intrinsics"], None)
        total = loss[0] + loss[1]
    elif cfg["variant"] in ("min", "minclip"):
        from plb200 import ops, _lib
        mam, _ = ops.fused_losses(g["tgt"], g["ref_imgs"], disp[:1], poses, g["intrinsics"], do_smooth=False,
                                  mode=_lib.PHOTO_MIN_REPROJ, clip_loss=0.5 if cfg["variant"] == "minclip" else None)
        total = mam
    else:
        from plb200 import ops
        mam, _ = ops.fused_losses(g["tgt"], g["ref_imgs"], disp[:1], poses, g["intrinsics"], do_smooth=False)
        total = mam
    total.backward()
    return total.detach(), poses.grad, [d.grad for fr in disp for d in fr]


def run_ours(args, cfg):
    from plb200 import synth, _lib
    from losses import Losses
    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        if not args.no_comm:
            os.environ.setdefault("NCCL_MAX_CTAS", str(NCCL_CTAS))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local if world > 1 else 0)
    torch.cuda.set_device(dev)
    criterion = Losses(smoothness="edge", deterministic=True) if cfg["variant"] == "live_edge" else Losses()
    from plb200 import ops as _ops
    n_sets = args.sets
    cpu_sets = make_sets(cfg, n_sets, 1234 + 1000 * rank, dev)
    gpu_sets = [synth.to_device(s, dev) for s in cpu_sets]
    pool_mb = sum(set_bytes(s) for s in cpu_sets) / 1e6
    px_per_step = cfg["B"] * cfg["H"] * cfg["W"]
    use_graph = not args.no_graph
    side = torch.cuda.Stream(device=dev)

    keep = []

    def capture():
        """One CUDA graph per input set, captured from the public-API calls of a step (the launch geometry - the
        persistent grid's size - is part of a graph, so a different sm_limit means a new capture)."""
        if use_graph and cfg["variant"] in ("live", "live_edge"):
            # the public graphed step (Losses.capture(), plb200/graphed.py): static buffers per input set, replayed
            # without copy-in; it owns its backward call, so no guarded relaunch is part of the graph
            n0 = _lib.launch_count()
            steps = [criterion.capture(g["tgt"], g["ref_imgs"], g["disparity"], g["poses"], g["intrinsics"]) for g in gpu_sets]
            per_step = (_lib.launch_count() - n0) // (3 * n_sets)          # two warm-up steps + the captured one
            keep.append(steps)
            return ([st.graph for st in steps],
                    [(st.total, st.grads.poses, [d for fr in st.grads.disparity for d in fr]) for st in steps], per_step)
        side.wait_stream(torch.cuda.current_stream())
        n0 = _lib.launch_count()
        with torch.cuda.stream(side):
            for g in gpu_sets:
                for _ in range(2):
                    step_fn(criterion, g, cfg)
        per_step = (_lib.launch_count() - n0) // (2 * n_sets)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs, outs = [], []
        if use_graph:
            for g in gpu_sets:
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg, stream=side):
                    outs.append(step_fn(criterion, g, cfg))
                graphs.append(cg)
        return graphs, outs, per_step

    graphs, outs, launches_per_step = capture()

    def device_step(i, graphs=graphs, outs=outs):
        if use_graph:
            graphs[i % n_sets].replay()
            return outs[i % n_sets]
        return step_fn(criterion, gpu_sets[i % n_sets], cfg)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- what the loss step itself exchanges in a data-parallel run (N > 1): the all-reduce
    #      of its two logged loss scalars (`trainer.py:264-266` is where DDP sits), issued every step on a side stream
    #      behind that step's kernels and INSIDE the timed region.  (The 66 MB network-gradient all-reduce of the same
    #      training step runs while the NETWORKS' backward pass computes - DDP's bucket hooks - which is outside this
    #      path and many times longer than the 0.2 ms loss; it is measured below as `exchange`, on its own.) ----------
    loss_comm = None
    if world > 1 and not args.no_comm:
        import torch.distributed as dist
        loss_comm = {"buf": torch.zeros(2, dtype=torch.float32, device=dev), "stream": torch.cuda.Stream(device=dev),
                     "ev": torch.cuda.Event()}

    def timed_loop(n, with_comm):
        main = torch.cuda.current_stream()
        for i in range(n):
            o = device_step(i)
            if with_comm and loss_comm is not None:
                loss_comm["ev"].record(main)
                loss_comm["stream"].wait_event(loss_comm["ev"])
                with torch.cuda.stream(loss_comm["stream"]):
                    loss_comm["buf"].copy_(o[0].detach().reshape(1).expand(2))
                    dist.all_reduce(loss_comm["buf"], op=dist.ReduceOp.SUM)
        if with_comm and loss_comm is not None:
            main.wait_stream(loss_comm["stream"])          # the last step's reduced scalars are part of the timed region

    def timed(n, with_comm, loop=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        (loop or timed_loop)(n, with_comm)
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    timed_loop(args.warmup, True)
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    ms_total = timed(args.steps, True)
    clocks = sampler.stop()

    # ---- N > 1, measured on its own: the network-gradient all-reduce of the same training step - the networks are out
    #      of scope, so the payload is a synthetic fp32 arena of the reference configuration's size (DispResNet 14.8 M +
    #      PoseFc 1.64 M parameters = 66 MB, configs/basic_config.yaml:3-10) - in 25 MB buckets on a side stream, loss
    #      scalars fused into the last bucket, launched at the start of a loss step whose persistent grid is sized for
    #      fewer SMs; a step ends when both have finished.  Reported: the all-reduce alone, the step without it, the
    #      step with it. -----------------------------------------------------------------------------------------------
    comm = None
    if world > 1 and not args.no_comm:
        from plb200 import dist as pdist
        _ops.set_sm_limit(SM_COUNT_FOR_LOSS)
        graphs_x, outs_x, _ = capture()
        n_grad = 14_800_000 + 1_640_000
        arena = torch.zeros(n_grad + 2, dtype=torch.float32, device=dev)
        red = pdist.GradBucketReducer(arena, bucket_mb=25.0)
        last_loss = [torch.zeros((), device=dev), torch.zeros((), device=dev)]

        def loop_x(n, with_comm):
            for i in range(n):
                if with_comm:
                    red.launch(losses=last_loss, B_local=cfg["B"], B_global=cfg["B"] * world)
                o = device_step(i, graphs_x, outs_x)
                if with_comm:
                    last_loss[0] = o[0]
                    last_loss[1] = o[0]
                    red.wait()

        loop_x(args.warmup, True)
        ms_nocomm = timed(args.steps, False, loop_x) / args.steps
        barrier()
        ec0, ec1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ec0.record()
        for i in range(args.steps):
            red.launch(losses=last_loss, B_local=cfg["B"], B_global=cfg["B"] * world)
            red.wait()
        ec1.record()
        barrier()
        ms_with = timed(args.steps, True, loop_x) / args.steps
        comm = {"payload_bytes": int(arena.numel() * 4), "buckets": len(red.bounds), "bucket_mb": 25.0,
                "ms_per_step_without_exchange": ms_nocomm, "allreduce_alone_ms": ec0.elapsed_time(ec1) / args.steps,
                "ms_per_step_with_exchange": ms_with, "loss_grid_sms": SM_COUNT_FOR_LOSS, "nccl_max_ctas": NCCL_CTAS}
        del graphs_x, outs_x, arena, red
        _ops.set_sm_limit(0)                       # the kernel-alone and e2e measurements below have the GPU to themselves

    # ---- dominant kernel alone (photo_l1 fused fwd+grad), back-to-back launches -----
    kern_ms = time_photo_kernel(criterion, gpu_sets, cfg, dev, max(20, min(args.steps, 200)))

    # ---- e2e: host (pinned) inputs -> H2D -> public API fwd+bwd -> D2H loss ----------
    e2e = time_e2e(criterion, cpu_sets, cfg, dev, max(5, min(args.steps, 50)), barrier)
    e2e_full = eager_step = None
    if world == 1 and not args.no_cloud:
        # the same with frames at KITTI's decoded size: the GPU does the resize as well (secondary number)
        f = time_e2e(criterion, cpu_sets, cfg, dev, max(5, min(args.steps, 30)), barrier, frame_hw=(375, 1242))
        e2e_full = {"value": px_per_step / 1e6 / (f["ms_per_step"] / 1e3), "unit": UNIT, "ms_per_step": f["ms_per_step"],
                    "h2d_bytes_per_step": f["h2d"], "frames": f["frames"]}
        ems, hms = time_eager(criterion, gpu_sets, cfg, dev, max(20, min(args.steps, 100)))
        eager_step = {"value": px_per_step / 1e6 / (ems / 1e3), "unit": UNIT, "ms_per_step": ems, "host_ms_per_step": hms,
                      "note": "same step issued eagerly through Losses.forward / backward (no CUDA graph), device-resident inputs"}
        if cfg["variant"] in ("live", "live_edge"):
            cms, chms = time_captured(criterion, gpu_sets, dev, max(20, min(args.steps, 100)))
            eager_step["captured"] = {"value": px_per_step / 1e6 / (cms / 1e3), "unit": UNIT, "ms_per_step": cms,
                                      "host_ms_per_step": chms,
                                      "note": "Losses.capture(): the public graphed step, fresh inputs copied into its static "
                                              "buffers (device to device) every step"}

    if world > 1:
        import torch.distributed as dist
        c0 = comm["ms_per_step_without_exchange"] if comm else 0.0
        c1 = comm["allreduce_alone_ms"] if comm else 0.0
        c2 = comm["ms_per_step_with_exchange"] if comm else 0.0
        t = torch.tensor([ms_total, e2e["ms_per_step"], kern_ms, c0, c1, c2], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e["ms_per_step"], kern_ms, c0, c1, c2 = [float(x) for x in t]
        if comm:
            comm["ms_per_step_without_exchange"], comm["allreduce_alone_ms"], comm["ms_per_step_with_exchange"] = c0, c1, c2
            comm["exposed_ms_per_step"] = max(0.0, c2 - c0)
            comm["value_with_exchange"] = world * px_per_step / 1e6 / (c2 / 1e3)
            comm["note"] = ("measured on its own, not part of `value`: in the training step this all-reduce overlaps the "
                            "networks' backward pass (DDP bucket hooks), not the 0.2 ms loss; `value` contains the loss "
                            "step's own collective, the all-reduce of its two logged scalars every step")
    ms_step = ms_total / args.steps
    value = world * px_per_step / 1e6 / (ms_step / 1e3)
    e2e_value = world * px_per_step / 1e6 / (e2e["ms_per_step"] / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        abytes = algorithmic_bytes_per_px(cfg) * px_per_step
        achieved = abytes / (kern_ms / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
        if os.path.isfile(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except ValueError:
                pass
        cpu = cpu_reference_run(cfg, 6, 1) if (world == 1 and not args.no_cpu) else None
        eager = gpu_eager_run(cfg, gpu_sets, dev) if (world == 1 and not args.no_eager) else None
        if eager:
            eager["ours_over_eager"] = value / eager["value"]
        others = None
        if world == 1 and not args.no_cloud:
            # the other photometric compositions on the same frames (kernel-only, same method as `roofline`)
            others = {}
            for name in ("headline", "headline64", "c2min", "c2minclip", "c2"):
                if name == args.workload:
                    continue
                ocfg = WORKLOADS[name]
                osets = [synth.to_device(s_, dev) for s_ in make_sets(ocfg, n_sets if ocfg["B"] <= 16 else 2, 1234, dev)]
                oms = time_photo_kernel(criterion, osets, ocfg, dev, 64)
                ob = algorithmic_bytes_per_px(ocfg) * ocfg["B"] * ocfg["H"] * ocfg["W"]
                others[name] = {"workload": workload_name(name, ocfg), "kernel_ms": oms,
                                "mpix_s": ocfg["B"] * ocfg["H"] * ocfg["W"] / 1e6 / (oms / 1e3),
                                "achieved_gbs": ob / (oms / 1e3) / 1e9, "frac": ob / (oms / 1e3) / 1e9 / peak}
                if not args.no_eager and name != "headline64":
                    oe = gpu_eager_run(ocfg, osets, dev, iters=5, warmup=2)
                    others[name]["gpu_eager_baseline"] = {"mpix_s": oe["value"], "ms_per_step": oe["ms_per_step"],
                                                          "kernel_over_eager": others[name]["mpix_s"] / oe["value"]}
                del osets
        cloud = velo = None
        if world == 1 and not args.no_cloud:
            cloud = time_cloud(dev)
            cloud["frac"] = cloud["achieved_gbs"] / peak
            cloud["f32_pointcloud2"]["frac"] = cloud["f32_pointcloud2"]["achieved_gbs"] / peak
            velo = time_velo(dev)
            velo["frac"] = velo["achieved_gbs"] / peak
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, cfg), "global_batch": cfg["B"] * world,
                       "parallelism": ("batch-sharded x%d, no data-path collective; per step: NCCL all-reduce of the two logged "
                                       "loss scalars on a side stream, inside the timed region (the 66 MB network-gradient "
                                       "all-reduce: `exchange`)" % world) if comm else
                                      "batch-sharded x%d, no data-path collective" % world,
                       "l2": "inputs rotate over %d distinct sets (%.0f MB per GPU) > 126 MB L2" % (n_sets, pool_mb),
                       "step": ("CUDA-graph replay of Losses.capture() (forward + backward over static buffers; the step owns its "
                                "backward call, so the graph holds no guarded relaunch)" if cfg["variant"] in ("live", "live_edge") else
                                "CUDA-graph replay of the public-API forward + backward") if use_graph else "eager public API"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                    "ms_per_step": e2e["ms_per_step"], "frames": e2e["frames"], "frame_bytes_per_step": e2e["frame_bytes"],
                    "fullres": e2e_full},
            "exchange": comm,
            "eager": eager_step,
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "photo_l1_kernel<GRAD> (fused fwd+grad, all directions/scales)",
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": abytes,
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650"},
            "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
            "gpu_eager_baseline": eager,
            "cloud": cloud,
            "velo": velo,
            "other_workloads": others,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def time_photo_kernel(criterion, gpu_sets, cfg, dev, iters):
    """Average duration of the dominant kernel, CUDA events on its own stream."""
    from plb200 import ops
    outs = []
    st = torch.cuda.current_stream()
    for g in gpu_sets:  # pre-build argument structs by running once
        step_fn(criterion, g, cfg)
    torch.cuda.synchronize()
    # launch only the fused photometric kernel (forward + unit-upstream gradients)
    calls = []
    for g in gpu_sets:
        pyr = g["disparity"] if cfg["variant"] in ("live", "live_edge") else g["disparity"][:1]
        from plb200 import _lib
        lcfg = ops.LossConfig(cfg["n_src"], [len(p) for p in pyr], do_smooth=False,
                              mode=_lib.PHOTO_MIN_REPROJ if cfg["variant"] in ("min", "minclip") else _lib.PHOTO_L1_MEAN,
                              clip_loss=0.5 if cfg["variant"] == "minclip" else None)
        g_pyr = [[torch.empty_like(d) for d in p] for p in pyr]
        g_poses = torch.zeros_like(g["poses"])
        out = torch.zeros(2, device=dev)
        calls.append((lcfg, g, pyr, g_pyr, g_poses, out))

    def launch(i):
        lcfg, g, pyr, g_pyr, g_poses, out = calls[i % len(calls)]
        ops._launch_loss(lcfg, g["tgt"], g["ref_imgs"], g["poses"], g["intrinsics"], pyr, True, g_pyr, g_poses,
                         None, None, out, None, False)
    for i in range(5):
        launch(i)
    torch.cuda.synchronize()
    # back-to-back launches replayed from a CUDA graph, so the host-side cost of issuing a launch
    # (ctypes marshalling) cannot hide in the measurement
    per_graph = 4 * len(calls)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(st)
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg, stream=side):
        for i in range(per_graph):
            launch(i)
    reps = max(2, iters // per_graph)
    cg.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        cg.replay()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * per_graph)


def _graph_replay_ms(calls, iters, dev):
    """Each call captured in its own CUDA graph (after one eager run that built its workspaces); the graphs replayed
    round robin, CUDA events around `iters` replays.  Returns ms per replay."""
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    graphs, keep = [], []
    with torch.cuda.stream(side):
        for c in calls:
            c()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    for c in calls:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            keep.append(c())
        graphs.append(g)
    for g in graphs:
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        graphs[i % len(graphs)].replay()
    e1.record()
    torch.cuda.synchronize()
    del graphs, keep
    return e0.elapsed_time(e1) / iters


def time_cloud(dev, B=32, H=375, W=1242, iters=20):
    """Config C4 (BASELINE.json configs[3]): depth -> pseudo-LiDAR back-projection.  The headline numbers are the fp64
    x,y,z,0 parity layout; `f32` is the PointCloud2 wire layout the reference publishes (PseudoLidarPipeline.py:51-54,
    16 B per point) written by the same kernel; `e2e` is project_batch with HOST buffers: pinned depth in, counts and
    only the kept rows of the cloud out.  HBM fraction on the algorithmic bytes of SURVEY.md section 8(d): 4 B/px read
    + 32 (16) B per kept point."""
    import tempfile
    from plb200 import synth
    from utils.PseudoLiDAR import PseudoLiDAR
    with tempfile.TemporaryDirectory() as d:
        pl = PseudoLiDAR(synth.write_kitti_calib(d), 0, device=dev)
    sets = [synth.make_depth_images(B, H, W, seed=40 + k).to(dev) for k in range(3)]   # 3 x 60 MB in, 3 x 477 MB out > L2
    st = torch.cuda.current_stream()
    px = B * H * W
    res = {}
    for layout, want, bpp in (("f64", dict(want_f64=True), 32.0), ("f32", dict(want_f64=False, want_f32=True), 16.0)):
        outs = [pl.project_batch(s, **want) for s in sets]
        torch.cuda.synchronize()
        kept = int(outs[0]["count"].sum())
        del outs
        # the calls issued eagerly through the public API ...
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(iters):
            pl.project_batch(sets[i % len(sets)], **want)
        e1.record(st)
        torch.cuda.synchronize()
        eager_ms = e0.elapsed_time(e1) / iters
        # ... and the two launches replayed from one CUDA graph per input set (like `value`: the host cost of issuing a
        # call - argument marshalling, the 477 MB output allocation - cannot hide in or add to the measurement)
        ms = _graph_replay_ms([lambda s=s: pl.project_batch(s, **want) for s in sets], iters, dev)
        abytes = 4.0 * px + bpp * kept
        res[layout] = {"ms": ms, "mpix_s": px / 1e6 / (ms / 1e3), "kept_points": kept, "algorithmic_bytes": abytes,
                       "achieved_gbs": abytes / (ms / 1e3) / 1e9, "eager_ms": eager_ms}
    # end to end with host buffers (fp64 parity layout): H2D of the depth batch, the two launches, D2H of the counts
    # (the output size is data dependent: one sync) and of the kept rows of every image
    host_in = [s.cpu().pin_memory() for s in sets]
    host_out = torch.empty(B * H * W, 4, dtype=torch.float64).pin_memory()
    dbuf = torch.empty_like(sets[0])

    def e2e_once(i):
        dbuf.copy_(host_in[i % len(host_in)], non_blocking=True)
        r = pl.project_batch(dbuf)
        counts = r["count"].cpu()
        o = 0
        for b in range(B):
            n = int(counts[b])
            host_out[o:o + n].copy_(r["cloud_f64"][b, :n], non_blocking=True)
            o += n
        torch.cuda.synchronize()
        return o
    e2e_once(0)
    t0 = time.perf_counter()
    n_e2e = 5
    for i in range(n_e2e):
        rows = e2e_once(i)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    out = {"workload": "c4: %dx%d depth -> pseudo-LiDAR, batch %d, f64 x,y,z,0 parity layout" % (H, W, B),
           "timing": "CUDA-graph replay of project_batch per input set (3 sets, > L2); eager_ms = the same calls issued eagerly"}
    out.update(res["f64"])
    out["f32_pointcloud2"] = res["f32"]
    out["e2e"] = {"ms": e2e_ms, "mpix_s": px / 1e6 / (e2e_ms / 1e3), "h2d_bytes": 4 * px, "d2h_bytes": 32 * rows + 4 * B,
                  "note": "pinned host depth in, counts + kept rows (f64) out, wall clock incl. the count sync"}
    return out


def time_velo(dev, B=32, N=123577, H=375, W=1242, iters=20):
    """Velodyne sweep -> sparse depth image (Transform.py:69-104), B sweeps of one
    KITTI frame's point count.  Algorithmic bytes: 16 B per point read + 8 B per cell written (fp64 image)."""
    import tempfile
    import numpy as np
    from plb200 import synth
    from Transform.Transform import Transform
    with tempfile.TemporaryDirectory() as d:
        tr = Transform(synth.write_kitti_calib(d), W, H, device=dev)
    one = [torch.from_numpy(synth.make_velodyne_cloud(N, seed=60 + k)) for k in range(4)]
    sets = [torch.stack([one[(k + j) % 4] for j in range(B)]).to(dev) for k in range(3)]
    tr.project_batch(sets[0])
    torch.cuda.synchronize()
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(iters):
        tr.project_batch(sets[i % len(sets)])
    e1.record(st)
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / iters
    ms = _graph_replay_ms([lambda s=s: tr.project_batch(s) for s in sets], iters, dev)     # as time_cloud
    abytes = 16.0 * B * N + 8.0 * B * H * W
    return {"workload": "velodyne -> image: %d sweeps x %d points -> %dx%d f64 depth" % (B, N, H, W), "ms": ms,
            "mpoints_s": B * N / 1e6 / (ms / 1e3), "algorithmic_bytes": abytes, "achieved_gbs": abytes / (ms / 1e3) / 1e9,
            "eager_ms": eager_ms, "timing": "CUDA-graph replay of project_batch per input set; eager_ms = the same calls issued eagerly"}


def time_e2e(criterion, cpu_sets, cfg, dev, iters, barrier, frame_hw=None):
    """Public API with HOST buffers, the way a training loop that owns its staging memory feeds the loss: every step
    copies that step's inputs from pinned host memory to the device - the decoded frames as uint8 [B,h,w,3] (what the
    loader holds before the reference's transform chain), the disparity pyramids, poses and intrinsics - runs the
    loader chain on the device (`FramePrep`: resize + normalise + intrinsics scaling, bit-exact with
    trainer.py:97-103 / dataloaders.py:32-49,95-98), then `Losses.forward` + backward, and reads the step's loss back
    to the host; all inside the timed region.  The copy of step i+1 runs on a copy stream while step i computes (two
    device buffer sets).  `frame_hw`: size of the decoded frames (default: the network resolution - a loader that
    resizes the bytes on the host; (375, 1242) = the GPU also does the resize)."""
    from plb200 import synth
    from plb200.frameprep import FramePrep
    H, W, B, n_src = cfg["H"], cfg["W"], cfg["B"], cfg["n_src"]
    fh, fw = frame_hw or (H, W)
    prep = FramePrep(H, W)

    def host_set(k, s):
        frames = synth.make_frames_u8(B, fh, fw, n_frames=1 + n_src, seed=900 + 31 * k)
        K_dec = synth.kitti_intrinsics(B, fh, fw)          # intrinsics at the decoded size
        return {"frames": torch.cat(frames, 0), "disparity": s["disparity"], "poses": s["poses"], "K": K_dec}

    def flat(g):
        return [g["frames"]] + [d for fr in g["disparity"] for d in fr] + [g["poses"], g["K"]]

    def arena_like(s, **kw):
        """One contiguous byte arena holding every input of a step (256-byte aligned views): the H2D copy of a
        step is ONE transfer."""
        ts = flat(s)
        offs, n = [], 0
        for t in ts:
            offs.append(n)
            n += (t.numel() * t.element_size() + 255) // 256 * 256
        arena = torch.empty(n, dtype=torch.uint8, **kw)
        views = [arena[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape) for o, t in zip(offs, ts)]
        g = {"frames": views[0], "disparity": [], "poses": views[-2], "K": views[-1]}
        k = 1
        for fr in s["disparity"]:
            g["disparity"].append(views[k:k + len(fr)])
            k += len(fr)
        return arena, g

    host_sets = [host_set(k, s) for k, s in enumerate(cpu_sets)]
    pinned = []
    for s in host_sets:
        arena, g = arena_like(s, pin_memory=True)
        for dst, src in zip(flat(g), flat(s)):
            dst.copy_(src)
        pinned.append((arena, g))
    h2d = pinned[0][0].numel()                            # bytes actually copied per step (views are 256-byte aligned)
    frame_bytes = host_sets[0]["frames"].numel()
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()

    bufs = [arena_like(host_sets[0], device=dev), arena_like(host_sets[0], device=dev)]
    planar = [torch.empty((1 + n_src) * B, 3, H, W, dtype=torch.float32, device=dev) for _ in range(2)]
    main_st = torch.cuda.current_stream()
    copy_st = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]     # H2D of the buffer finished
    consumed = [torch.cuda.Event(), torch.cuda.Event()]   # the step that read the buffer finished

    def upload(i):
        k = i % 2
        with torch.cuda.stream(copy_st):
            copy_st.wait_event(consumed[k])
            bufs[k][0].copy_(pinned[i % len(pinned)][0], non_blocking=True)
            copied[k].record(copy_st)

    def run(n):
        for k in range(2):
            consumed[k].record(main_st)
        upload(0)
        last = 0.0
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)                              # overlaps with the compute of step i
            k = i % 2
            main_st.wait_event(copied[k])
            g = bufs[k][1]
            out = prep(g["frames"], g["K"], out=planar[k])
            img = out["planar"]
            step_in = {"tgt": img[:B], "ref_imgs": [img[(1 + j) * B:(2 + j) * B] for j in range(n_src)],
                       "disparity": g["disparity"], "poses": g["poses"], "intrinsics": out["K"]}
            total, _, _ = step_fn(criterion, step_in, cfg)
            consumed[k].record(main_st)
            host_loss.copy_(total, non_blocking=False)     # D2H of the step's result (synchronises)
            last = float(host_loss)
        return last

    run(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(iters)
    e1.record()
    barrier()
    return {"ms_per_step": e0.elapsed_time(e1) / iters, "h2d": h2d, "d2h": 4, "frame_bytes": frame_bytes,
            "frames": "uint8 [%d,%d,%d,3] decoded frames -> FramePrep on the device" % ((1 + n_src) * B, fh, fw)}


def time_eager(criterion, gpu_sets, cfg, dev, iters):
    """The same step issued eagerly through the public API (no CUDA graph): what a trainer that calls
    `criterion.forward` + `backward` every iteration gets, host time included."""
    leaves = [make_leaves(g) for g in gpu_sets]
    for i in range(5):
        step_fn(criterion, gpu_sets[i % len(gpu_sets)], cfg, leaves[i % len(gpu_sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(iters):
        step_fn(criterion, gpu_sets[i % len(gpu_sets)], cfg, leaves[i % len(gpu_sets)])
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / iters
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, host_ms


def time_captured(criterion, gpu_sets, dev, iters):
    """`Losses.capture()` (plb200/graphed.py): the step a trainer gets when it graphs the loss tail - inputs that live
    in other buffers (here: the rotating input sets) are copied into the captured step's static buffers, then one
    graph launch."""
    g0 = gpu_sets[0]
    step = criterion.capture(g0["tgt"], g0["ref_imgs"], g0["disparity"], g0["poses"], g0["intrinsics"])

    def one(i):
        g = gpu_sets[i % len(gpu_sets)]
        step(g["tgt"], g["ref_imgs"], g["disparity"], g["poses"], g["intrinsics"])
    for i in range(5):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(iters):
        one(i)
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / iters
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, host_ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--sets", type=int, default=4, help="distinct input sets rotated between steps (L2 defeat)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cloud", action="store_true", help="skip the secondary pseudo-LiDAR (config C4) timing")
    ap.add_argument("--no-comm", action="store_true", help="N > 1: leave the network-gradient all-reduce out of the step")
    ap.add_argument("--no-eager", action="store_true", help="skip the torch-eager-CUDA incumbent (gpu_eager_baseline)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    import __graft_entry__ as ge
    from plb200 import build as _b
    if not os.path.isfile(_b.LIB):
        ge.build()
    run_ours(args, cfg)


if __name__ == "__main__":
    main()

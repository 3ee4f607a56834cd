#!/usr/bin/env python
"""Benchmark of the ray-rendering training hot path (BASELINE.json metric: training rays/sec, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n-rand R] [--impl ours|reference]

One "step" = render(coarse+fine) + RGB/depth loss + backward over one batch of synthetic LLFF-shaped rays
(config B of SURVEY.md §8: fern_dsnerf.txt shapes, N_rand=4096 -> 2048 RGB + 2048 depth rays, 64+64
samples, coarse D=4 / fine D=8, W=256, use_viewdirs, perturb=1, raw_noise_std=1, NDC), the optimiser
step excluded as in BASELINE.md §3.  For N>1 every rank renders its own N_rand rays (weak scaling) and the
MLP gradients are all-reduced over NCCL inside the step.

Prints ONE JSON line (see the task contract): `value` = rays/s with the ray batch resident in HBM;
`e2e` = the same step through the drop-in API from pinned HOST buffers (H2D of the ray batch + targets and a
non-blocking D2H of the loss every step, read by the host two steps later; `e2e.sync_every_step` = the same with
`loss.item()` after every step); `roofline` for the dominant kernel from CUDA events recorded around every
launch of the timed region; `cpu_baseline` = the oracle port of the reference on the host cores.

`--impl reference` times the reference's own CPU implementation (the oracle port: the reference is Python
and cannot travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, FOCAL = 378, 504, 407.6
N_SAMPLES, N_IMPORTANCE = 64, 64
COARSE_D, FINE_D = 4, 8
SEMANTIC_LAMBDA = 0.01      # configs/fern_dsnerf.txt:56
DEPTH_LAMBDA = 0.01

# algorithmic (unpadded) MACs per point, SURVEY.md §8(d)
MACS_FWD = {8: 593408, 4: 315136}
MACS_DGRAD = {8: 557696, 4: 295552}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled every 10 ms through NVML while the timed region runs (the timed
    region of a default run is ~0.1 s, too short for `nvidia-smi -lms 200`)."""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.sm, self.mask, self.max_mhz, self.power = index, [], 0, None, []
        self._stop = threading.Event()
        self.th = None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[0].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        except Exception as e:       # NVML missing: record nothing rather than fail the bench
            self.err = repr(e)
        return self

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.01)

    def __exit__(self, *a):
        self._stop.set()
        if self.th is not None:
            self.th.join(timeout=1)

    def summary(self):
        sm = sorted(self.sm)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": self.max_mhz,
                "power_w_max": max(self.power) if self.power else None,
                "reasons": sorted(k for k, b in self.REASONS.items() if self.mask & b), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
def make_batch(n_rand, seed):
    """Synthetic LLFF-shaped batch: forward-facing pinhole cameras near the origin with small
    pose jitter, uniformly random pixels; colour targets U(0,1)^3; NDC depth targets U(0.25,1) with ~15 %
    'sky' rays at 1-1e-7.  RGB rays first, depth rays after (run_nerf.py:1409-1411)."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    n_dep = n_rand // 2
    n_rgb = n_rand - n_dep
    px, py = torch.rand(n_rand, generator=g) * (W - 1), torch.rand(n_rand, generator=g) * (H - 1)
    d = torch.stack([(px - W * .5) / FOCAL, -(py - H * .5) / FOCAL, -torch.ones_like(px)], -1)
    ang = (torch.rand(n_rand, 2, generator=g) - .5) * (10. * math.pi / 180.)
    cx, sx, cy, sy = torch.cos(ang[:, 0]), torch.sin(ang[:, 0]), torch.cos(ang[:, 1]), torch.sin(ang[:, 1])
    dy, dz = d[:, 1] * cx - d[:, 2] * sx, d[:, 1] * sx + d[:, 2] * cx
    rd = torch.stack([d[:, 0] * cy + dz * sy, dy, -d[:, 0] * sy + dz * cy], -1)
    t = torch.rand(n_rand, 3, generator=g)
    ro = torch.stack([(t[:, 0] - .5) * .6, (t[:, 1] - .5) * .6, (t[:, 2] - .5) * .1], -1)
    tgt = torch.rand(n_rgb, 3, generator=g)
    dep = .25 + .75 * torch.rand(n_dep, generator=g)
    dep = torch.where(torch.rand(n_dep, generator=g) < .15, torch.full_like(dep, 1. - 1e-7), dep)
    return ro.float(), rd.float(), tgt.float(), dep.float(), n_rgb, n_dep


def _ref_step_fn(n_sample, semK=0, seed=3407):
    """One step of the reference's CPU path (oracle port): render + loss + backward on n_sample rays."""
    import torch
    from oracle import nerf_oracle as O
    spec_c = O.MLPSpec(D=COARSE_D, semantic_num_classes=semK)
    spec_f = O.MLPSpec(D=FINE_D, semantic_num_classes=semK)
    pc = {k: v.requires_grad_(True) for k, v in O.init_params(spec_c, 3407 + COARSE_D).items()}
    pf = {k: v.requires_grad_(True) for k, v in O.init_params(spec_f, 3407 + FINE_D).items()}
    ro, rd, tgt, dep, n_rgb, n_dep = make_batch(n_sample, seed)
    rb = O.pack_rays(H, W, FOCAL, ro, rd)
    tsem = torch.randint(0, semK, (n_rgb,), generator=torch.Generator().manual_seed(seed)) if semK else None

    def step():
        rng = O.RenderRNG(torch.rand(n_sample, N_SAMPLES), torch.randn(n_sample, N_SAMPLES),
                          torch.rand(n_sample, N_IMPORTANCE), torch.randn(n_sample, N_SAMPLES + N_IMPORTANCE))
        out = O.render_rays(rb, pc, spec_c, pf, spec_f, N_SAMPLES, N_IMPORTANCE, rng, raw_noise_std=1.0,
                            semantic_loss=bool(semK))
        res = O.train_loss(out, n_rgb, tgt, dep, depth_lambda=DEPTH_LAMBDA, depth_importance=1.0,
                           target_semantic=tsem, semantic_lambda=SEMANTIC_LAMBDA if semK else 0.0)
        for p in list(pc.values()) + list(pf.values()):
            p.grad = None
        res["loss"].backward()
        return float(res["loss"].detach())
    return step, n_rgb, n_dep


def _time_steps(step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / max(steps, 1)


def run_reference(args):
    """CPU arm: the oracle port of the reference's render + loss + backward on the box's host cores -- the WHOLE
    N_rand-ray step when the host has the memory for its fp32 activations (~2.6 MB/ray) and K+W steps of it fit a few
    minutes, else a bounded ray sample (rays/s on the CPU does not depend on the batch size).  The timed loop runs with
    anomaly detection off; the as-shipped setting (`torch.autograd.set_detect_anomaly(True)`, run_nerf_helpers.py:6)
    and the semantic-head variant (fern_dsnerf.txt:55) are timed on 2 steps each and reported beside it."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is rank 0 alone on the box's host cores
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(3407)
    torch.autograd.set_detect_anomaly(False)
    total = args.steps + args.warmup
    n_rand = rays_per_gpu(args)
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    # ~1 300 rays/s on 16 cores: the whole step if K+W of them take < ~4 min, else a 64-ray multiple that does
    budget_rays = int(240.0 * 1300.0 / max(total, 1))
    n_sample = n_rand if (n_rand <= budget_rays and avail > n_rand * 6.0e6) else int(
        max(128, min(n_rand, 1024, budget_rays // 64 * 64)))
    step, n_rgb, n_dep = _ref_step_fn(n_sample)
    dt = _time_steps(step, args.steps, args.warmup)
    v = n_sample / dt
    cores = torch.get_num_threads()
    # as shipped: anomaly detection on (2 steps after 1 warm-up, same sample)
    torch.autograd.set_detect_anomaly(True)
    try:
        dt_anom = _time_steps(step, 2, 1)
    finally:
        torch.autograd.set_detect_anomaly(False)
    del step
    # fern_dsnerf.txt:55-56 as shipped: semantic head on (19 classes), anomaly off
    n_sem = min(n_sample, 1024)
    sem_step, _, _ = _ref_step_fn(n_sem, semK=19)
    dt_sem = _time_steps(sem_step, 2, 1)
    sample = "%d of the %d rays of the step (%d RGB + %d depth), full 64+64 samples, D=%d/%d nets" % (
        n_sample, n_rand, n_rgb, n_dep, COARSE_D, FINE_D)
    print(json.dumps({
        "impl": "reference", "metric": "training rays/sec (fwd+bwd)", "value": v, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": scaling_kind(args), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample,
                         "os_cpu_count": os.cpu_count(), "anomaly_detection": False,
                         "anomaly_on": {"value": n_sample / dt_anom, "unit": "rays/s", "ms_per_step": dt_anom * 1e3,
                                        "what": "torch.autograd.set_detect_anomaly(True) as run_nerf_helpers.py:6 ships it; "
                                                "2 steps after 1 warm-up on the same sample"}},
        "variants": {"semantic_head_19_classes": {
            "value": n_sem / dt_sem, "unit": "rays/s", "ms_per_step": dt_sem * 1e3,
            "what": "semantic_loss = True (fern_dsnerf.txt:55-56), %d rays, anomaly off, 2 steps after 1 warm-up" % n_sem}},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def rays_per_gpu(args):
    """--global-n-rand G (strong scaling, config E): G rays per step split evenly over the GPUs; else --n-rand each."""
    if args.global_n_rand:
        return max(2, args.global_n_rand // max(args.gpus, 1))
    return args.n_rand


def scaling_kind(args):
    return "strong" if args.global_n_rand else "weak"


def workload_config(args):
    """Identical in both arms (`--impl ours` / `--impl reference`) for the same command line."""
    world, n_rand = max(args.gpus, 1), rays_per_gpu(args)
    return {"workload": "fern_dsnerf.txt coarse+fine training step with depth loss, synthetic LLFF-shaped rays",
            "n_rand_per_gpu": n_rand, "global_rays_per_step": n_rand * world,
            "rgb_rays": n_rand - n_rand // 2, "depth_rays": n_rand // 2,
            "N_samples": N_SAMPLES, "N_importance": N_IMPORTANCE, "netdepth": COARSE_D, "netdepth_fine": FINE_D,
            "netwidth": 256, "use_viewdirs": True, "perturb": 1.0, "raw_noise_std": 1.0, "ndc": True,
            "depth_lambda": DEPTH_LAMBDA, "optimizer_step": "excluded (BASELINE.md §3)",
            "parallelism": "ray-sharded dp%d, NCCL all-reduce of MLP grads" % world,
            "semantic_head": ("off (headline workload: RGB + LiDAR-depth loss; the head-on step is in `variants`)" if not args.semantic else
                              "on: %d classes, cross-entropy of fine + coarse per-ray logits, lambda %g (fern_dsnerf.txt:55-56)" % (args.semantic, SEMANTIC_LAMBDA)),
            "l2": "per-step working set (activation stashes, ~1.5 MB/ray = 6 GB per 4096-ray step) is far larger than the 126 MB L2; no flush"}


def cpu_baseline(seconds_budget=25.0):
    """Oracle port of the reference on the host cores, config A (1024 rays), bounded to ~25 s."""
    import torch
    from oracle import nerf_oracle as O
    n = 1024
    spec_c, spec_f = O.MLPSpec(D=COARSE_D), O.MLPSpec(D=FINE_D)
    pc = {k: v.requires_grad_(True) for k, v in O.init_params(spec_c, 3407 + COARSE_D).items()}
    pf = {k: v.requires_grad_(True) for k, v in O.init_params(spec_f, 3407 + FINE_D).items()}
    ro, rd, tgt, dep, n_rgb, n_dep = make_batch(n, 3407)
    rb = O.pack_rays(H, W, FOCAL, ro, rd)
    times = []
    t_start = time.perf_counter()
    it = 0
    while True:
        rng = O.synth_rng(n, N_SAMPLES, N_IMPORTANCE, seed=it)
        t0 = time.perf_counter()
        out = O.render_rays(rb, pc, spec_c, pf, spec_f, N_SAMPLES, N_IMPORTANCE, rng, raw_noise_std=1.0)
        res = O.train_loss(out, n_rgb, tgt, dep, depth_lambda=DEPTH_LAMBDA, depth_importance=1.0)
        for p in list(pc.values()) + list(pf.values()):
            p.grad = None
        res["loss"].backward()
        times.append(time.perf_counter() - t0)
        it += 1
        if it >= 2 and (time.perf_counter() - t_start > seconds_budget or it >= 8):
            break
    ts = sorted(times[1:])
    med = ts[len(ts) // 2]
    # as shipped (run_nerf_helpers.py:6): anomaly detection on, 2 more steps
    torch.autograd.set_detect_anomaly(True)
    try:
        ta = []
        for it2 in range(2):
            rng = O.synth_rng(n, N_SAMPLES, N_IMPORTANCE, seed=100 + it2)
            t0 = time.perf_counter()
            out = O.render_rays(rb, pc, spec_c, pf, spec_f, N_SAMPLES, N_IMPORTANCE, rng, raw_noise_std=1.0)
            res = O.train_loss(out, n_rgb, tgt, dep, depth_lambda=DEPTH_LAMBDA, depth_importance=1.0)
            for p in list(pc.values()) + list(pf.values()):
                p.grad = None
            res["loss"].backward()
            ta.append(time.perf_counter() - t0)
    finally:
        torch.autograd.set_detect_anomaly(False)
    return {"anomaly_on": {"value": n / min(ta), "unit": "rays/s", "ms_per_step": min(ta) * 1e3,
                           "what": "torch.autograd.set_detect_anomaly(True) as run_nerf_helpers.py:6 ships it; best of 2"},
            "value": n / med, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "config A: 1024 rays (512 RGB + 512 depth), 64+64 samples, D=4/8; median of %d steps after 1 warm-up"
                      % len(ts), "os_cpu_count": os.cpu_count(), "ms_per_step": med * 1e3, "anomaly_detection": False}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import dlnerf_b200 as dn
    L = dn._lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    args.gpus = world
    args.n_rand = rays_per_gpu(args)          # --global-n-rand: the step's rays split evenly over the ranks
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(3407 + rank)

    torch.manual_seed(3407)              # identical random-init weights (nn.Linear default) on every rank
    semK = int(args.semantic)           # > 0: fern_dsnerf.txt:55-56 (semantic_loss = True, semantic_lambda = 0.01)
    net_c = dn.NeRF(D=COARSE_D, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True,
                    semantic_num_classes=semK or None).to(dev)
    net_f = dn.NeRF(D=FINE_D, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True,
                    semantic_num_classes=semK or None).to(dev)
    torch.manual_seed(3407 + rank)
    params = list(net_c.parameters()) + list(net_f.parameters())
    q = dn.FusedQuery(dn.get_embedder(10, 0)[0], dn.get_embedder(4, 0)[0], 65536, 10, 4, 0)
    kw = dict(network_query_fn=q, perturb=1.0, N_importance=N_IMPORTANCE, network_fine=net_f, N_samples=N_SAMPLES,
              network_fn=net_c, use_viewdirs=True, white_bkgd=False, raw_noise_std=1.0, ndc=True, near=0., far=1.,
              semantic_loss=bool(semK))

    ro, rd, tgt, dep, n_rgb, n_dep = make_batch(args.n_rand, 3407 + rank)
    host_rays = torch.stack([ro, rd], 0).pin_memory()
    host_tgt, host_dep = tgt.pin_memory(), dep.pin_memory()
    d_rays, d_tgt, d_dep = host_rays.to(dev), host_tgt.to(dev), host_dep.to(dev)
    # class index per RGB ray (run_nerf.py:1331-1332); device-resident in every route, 8 B/ray next to the 36 B/ray of rays
    d_sem = (torch.randint(0, semK, (n_rgb,), generator=torch.Generator().manual_seed(3407 + rank)).to(dev)
             if semK else None)
    sem_kw = dict(target_semantic=d_sem, semantic_lambda=SEMANTIC_LAMBDA) if semK else {}

    def allreduce_grads():
        dn.allreduce_gradients(params, world)

    def step(rays, t_rgb, t_dep):
        """Drop-in route: the reference's own call sequence (run_nerf.py:1416-1418, :1500-1536, :1759-1761, :1773)."""
        rgb, disp, acc, depth, extras = dn.render(H, W, FOCAL, chunk=1 << 30, rays=rays, retraw=True, **kw)
        for p in params:
            p.grad = None
        loss = dn.img2mse(rgb[:n_rgb], t_rgb) + DEPTH_LAMBDA * dn.img2mse(depth[n_rgb:], t_dep) \
            + dn.img2mse(extras["rgb0"][:n_rgb], t_rgb)
        if semK:                                                  # run_nerf.py:1541-1548
            ce = torch.nn.functional.cross_entropy
            loss = loss + SEMANTIC_LAMBDA * (ce(extras["sem_preds"][:n_rgb], d_sem) + ce(extras["sem_preds0"][:n_rgb], d_sem))
        loss.backward()
        allreduce_grads()
        return loss

    sched = {}                  # SM partition between the fine net's write-bound kernels and the coarse backward
    if args.coarse_sms:
        sched = dict(coarse_sms=args.coarse_sms, fine_sms=args.fine_sms or None)

    def fused_step(rays, t_rgb, t_dep, overlap=True):
        """Same step through dlnerf_b200.train_step: loss gradient fused into the compositing backward kernel."""
        out = dn.train_step(H, W, FOCAL, rays, t_rgb, t_dep, n_rgb, net_c, net_f, N_samples=N_SAMPLES,
                            N_importance=N_IMPORTANCE, perturb=1., raw_noise_std=1., depth_lambda=DEPTH_LAMBDA,
                            depth_importance=1., world_size=world, overlap_coarse_backward=overlap,
                            **(sched if overlap else {}), **sem_kw)
        return out["loss"]

    graphed = None
    if args.path == "graph":
        graphed = dn.GraphedTrainStep(H, W, FOCAL, args.n_rand, n_rgb, net_c, net_f, world_size=world,
                                      N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=1., raw_noise_std=1.,
                                      depth_lambda=DEPTH_LAMBDA, depth_importance=1., **sched,
                                      **({"semantic_lambda": SEMANTIC_LAMBDA} if semK else {}))

    def graph_step(rays, t_rgb, t_dep):
        """train_step replayed from a CUDA graph (one launch per step; weight re-pack included in the graph)."""
        return graphed(rays, t_rgb, t_dep, target_semantic=d_sem)["loss"]

    dev_step = {"dropin": step, "fused": fused_step, "graph": graph_step}[args.path]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput -------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        dev_step(d_rays, d_tgt, d_dep)
    L.TRACE = []
    n0 = L.LAUNCHES
    with ClockSampler(local) as clk:
        ms = timed(lambda: dev_step(d_rays, d_tgt, d_dep), args.steps)
    trace, L.TRACE = L.TRACE, None
    launches = (L.LAUNCHES - n0) // max(args.steps, 1)
    value = args.n_rand * world / (ms * 1e-3)
    kernel_times_from = "CUDA events around every launch of the timed region"
    ms_share = ms          # step time the per-kernel shares refer to
    if args.path in ("graph", "fused"):
        # a graph replay does not pass through the Python launch hooks (and the fused route overlaps two streams): take the per-kernel CUDA-event times (and
        # the launch count) from an eager pass of the very same step right after the timed region
        L.TRACE = []
        n0 = L.LAUNCHES
        timed(lambda: fused_step(d_rays, d_tgt, d_dep, overlap=False), args.steps)
        trace, L.TRACE = L.TRACE, None
        ms_share = None      # shares refer to the summed kernel time of this pass (its wall time carries the hooks)
        launches = (L.LAUNCHES - n0) // max(args.steps, 1)
        kernel_times_from = ("CUDA events around every launch of an eager, single-stream pass of the same %d steps (the "
                             "timed region replays them from a CUDA graph in which the coarse-net backward runs on a "
                             "second stream next to the fine-net backward, so there the kernels overlap)" % args.steps)

    # ---- per-kernel times from the events of the timed region --------------------------------------
    agg = {}
    for tag, a, b in trace:
        t = a.elapsed_time(b)
        s = agg.setdefault(tag, [0.0, 0])
        s[0] += t
        s[1] += 1
    kern = {k: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] / args.steps,
                "ms_per_step": v[0] / args.steps} for k, v in agg.items()}
    if ms_share is None:
        ms_share = sum(v["ms_per_step"] for v in kern.values())
    pk = peaks()
    pts = {COARSE_D: args.n_rand * N_SAMPLES, FINE_D: args.n_rand * (N_SAMPLES + N_IMPORTANCE)}
    flops = {}
    for D in (COARSE_D, FINE_D):
        flops["mlp_fwd D=%d" % D] = 2.0 * MACS_FWD[D] * pts[D]
        flops["mlp_dgrad D=%d" % D] = 2.0 * MACS_DGRAD[D] * pts[D]
        flops["mlp_wgrad D=%d" % D] = 2.0 * MACS_FWD[D] * pts[D]
    for k in kern:
        if k in flops:
            kern[k]["tflops"] = flops[k] / (kern[k]["ms_per_launch"] * 1e-3) / 1e12
            kern[k]["frac_of_sustained_peak"] = kern[k]["tflops"] / pk["tf_sust"]
            kern[k]["frac_of_burst_peak"] = kern[k]["tflops"] / pk["tf_burst"]
    mlp_keys = [k for k in kern if k in flops]
    top = max(mlp_keys, key=lambda k: kern[k]["ms_per_step"])
    mlp_ms = sum(kern[k]["ms_per_step"] for k in mlp_keys)
    mlp_tflops = sum(flops[k] for k in mlp_keys) / (mlp_ms * 1e-3) / 1e12
    # ---- HBM view of the MLP kernels: algorithmic bytes = the activation / dZ stash every point needs once (16 KB slab
    # per 128 points and 64 features) + the 1-bit ReLU masks + the point's in / out rows.  The chain
    # kernels WRITE their stash (a pure write stream), wgrad READS every slab it needs once.
    hbm_bytes = {}
    for D, net in ((COARSE_D, net_c), (FINE_D, net_f)):
        pl, P = net._plan, pts[D]
        mask_b = 32 * pl.mask_slots
        used = {(0 if it.b_from_bwd == 0 else 1, it.b_slot + i) for it in pl.wgrad for i in range(it.b_nslab)} | \
               {(1, it.a_slot + i) for it in pl.wgrad for i in range(it.a_nslab)}
        hbm_bytes["mlp_fwd D=%d" % D] = P * (128 * pl.fwd_slots + mask_b + 16 + 4)
        hbm_bytes["mlp_dgrad D=%d" % D] = P * (128 * pl.bwd_slots + mask_b + 16)
        hbm_bytes["mlp_wgrad D=%d" % D] = P * 128 * len(used)
    for k in mlp_keys:
        kern[k]["hbm_algorithmic_bytes"] = hbm_bytes[k]
        kern[k]["hbm_gbs"] = hbm_bytes[k] / (kern[k]["ms_per_launch"] * 1e-3) / 1e9
        kern[k]["frac_of_hbm_peak"] = kern[k]["hbm_gbs"] / pk["hbm"]
    # a pure WRITE stream does not reach the copy figure of MEASURED_PEAKS.json: measured here with a memset
    wbuf = torch.empty(1 << 31, dtype=torch.uint8, device=dev)
    wbuf.zero_()
    torch.cuda.synchronize()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record()
    for _ in range(3):
        wbuf.zero_()
    w1.record()
    torch.cuda.synchronize()
    write_peak = 3 * wbuf.numel() / (w0.elapsed_time(w1) * 1e-3) / 1e9
    del wbuf
    traffic, traffic_from = None, None   # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", name)))
            if tj.get("n_rand") == args.n_rand and tj["dram_bytes_per_launch"].get(top):
                traffic, traffic_from = tj["dram_bytes_per_launch"][top], "profiles/" + name
                break
        except (OSError, ValueError, KeyError):
            pass
    # The dominant kernel is reported against the roofline it sits closer to.  Since round 2 all six MLP kernels of a
    # training step are memory-bound: the chain kernels by their stash writes, wgrad by its stash reads.
    t_frac, h_frac = kern[top]["tflops"] / pk["tf_burst"], kern[top]["frac_of_hbm_peak"]
    by_hbm = h_frac >= t_frac
    roofline = {"kernel": top, "bound": "hbm" if by_hbm else "tensor",
                "achieved": kern[top]["hbm_gbs"] if by_hbm else kern[top]["tflops"],
                "peak": pk["hbm"] if by_hbm else pk["tf_burst"], "unit": "GB/s" if by_hbm else "TFLOP/s",
                "frac": h_frac if by_hbm else t_frac, "traffic": traffic, "traffic_from": traffic_from,
                "peak_source": "MEASURED_PEAKS.json %s (%s)" % ("hbm_gbs" if by_hbm else "bf16_tflops (burst)", pk["src"]),
                "tensor": {"achieved": kern[top]["tflops"], "unit": "TFLOP/s", "frac_of_burst": t_frac,
                           "frac_of_sustained": kern[top]["tflops"] / pk["tf_sust"]},
                "hbm": {"achieved": kern[top]["hbm_gbs"], "unit": "GB/s", "frac": h_frac,
                        "algorithmic_bytes_per_launch": hbm_bytes[top]},
                "write_only_peak": {"value": write_peak, "unit": "GB/s",
                                    "what": "cudaMemset of 2 GiB timed in this run: what a pure write stream reaches on this "
                                            "GPU (MEASURED_PEAKS.json's hbm_gbs is a copy, half reads); the chain kernels' "
                                            "stash is such a stream",
                                    "frac_of_it": {k: kern[k]["hbm_gbs"] / write_peak for k in mlp_keys if "wgrad" not in k}},
                "share_of_step": kern[top]["ms_per_step"] / ms_share,
                "all_mlp_kernels": {"tflops": mlp_tflops, "frac": mlp_tflops / pk["tf_burst"],
                                    "frac_of_sustained": mlp_tflops / pk["tf_sust"],
                                    "hbm_gbs": sum(hbm_bytes[k] for k in mlp_keys) / (mlp_ms * 1e-3) / 1e9,
                                    "ms_per_step": mlp_ms, "share_of_step": mlp_ms / ms_share},
                "share_basis_ms_per_step": ms_share,
                "flops_basis": "algorithmic unpadded MACs/point of the reference's layer structure (SURVEY §8d) x points "
                               "per launch; the kernels fold feature_linear into views_linears (no activation between "
                               "them) and so execute 65 536 MACs/point fewer per pass (89 % of the algorithmic count "
                               "for D=8, 79 % for D=4) for the same result",
                "bytes_basis": "stash slabs (128 B per point and 64 features) the plan writes (chain kernels) or reads once "
                               "(wgrad) + 32 B of ReLU masks per layer + the point's input / output rows"}

    # ---- end to end from pinned host memory ------------------------------------------------------------
    # The drop-in route keeps every chunk's activations alive for autograd (as the reference does), so above the
    # ray-chunk size it cannot hold the batch; there the end-to-end number goes through train_step (ray-chunked).
    big = args.n_rand > dn.default_ray_chunk(net_c, net_f, N_SAMPLES, N_IMPORTANCE)
    e2e_fn = fused_step if big else step
    e2e_api = ("dlnerf_b200.train_step(...) (ray-chunked; the drop-in autograd route cannot hold %d rays)" % args.n_rand
               if big else "dlnerf_b200.render(...) + img2mse + loss.backward() (drop-in path)")

    # The loss of every step is read back to the host.  Headline: a non-blocking 4-byte copy into pinned memory per step,
    # consumed two steps later (the host waits for step k-2 before it launches step k: at most two steps in flight, so
    # the GPU always has a queued step and the allocator never sees more than two steps of live buffers) -- the
    # reference's loop reads its loss every i_print = 100 iterations (run_nerf.py:1943-1959) and otherwise never waits
    # for the device.  `sync_every_step` is the same loop with `.item()` after every step (the host then starts each
    # step with an idle GPU).
    loss_ring = torch.zeros(64, dtype=torch.float32).pin_memory()
    ring_ev = [torch.cuda.Event() for _ in range(64)]
    ring_i = [0]
    seen = []

    def read_back(loss):
        i = ring_i[0]
        loss_ring[i & 63].copy_(loss.detach().reshape(()), non_blocking=True)
        ring_ev[i & 63].record()
        if i >= 2:
            ring_ev[(i - 2) & 63].synchronize()
            seen.append(float(loss_ring[(i - 2) & 63]))
        ring_i[0] = i + 1

    def e2e_step(fn=None, wait=False):
        r = host_rays.to(dev, non_blocking=True)
        t1 = host_tgt.to(dev, non_blocking=True)
        t2 = host_dep.to(dev, non_blocking=True)
        loss = (fn or e2e_fn)(r, t1, t2)
        if wait:
            return float(loss.item())
        read_back(loss)

    h2d = int(host_rays.numel() + host_tgt.numel() + host_dep.numel()) * 4
    for _ in range(3):
        e2e_step(wait=True)
    ms_sync = timed(lambda: e2e_step(wait=True), args.steps)
    for _ in range(max(args.warmup, 3)):          # the host runs several steps ahead here: let the allocator grow first
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    torch.cuda.synchronize()
    assert len(seen) >= args.steps and all(v == v and 0. < v < float("inf") for v in seen), \
        "read-back losses must be finite: %s" % seen[-8:]
    e2e = {"value": args.n_rand * world / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "api": e2e_api,
           "loss_read": "non-blocking D2H copy of the loss into pinned memory every step, read by the host two steps later "
                        "(at most two steps in flight; the reference reads its loss every i_print = 100 iterations, "
                        "run_nerf.py:1943-1959)",
           "sync_every_step": {"value": args.n_rand * world / (ms_sync * 1e-3), "unit": "rays/s", "ms_per_step": ms_sync,
                               "loss_read": "loss.item() after every step"}}
    if graphed is not None:
        # same host-buffer protocol through the CUDA-graph step (pinned host tensors are copied straight into the
        # graph's static inputs)
        def e2e_graph_step(wait=False):
            loss = graphed(host_rays, host_tgt, host_dep, target_semantic=d_sem)["loss"]
            if wait:
                return float(loss.item())
            read_back(loss)
        for _ in range(3):
            e2e_graph_step(wait=True)
        ms_gs = timed(lambda: e2e_graph_step(wait=True), args.steps)
        for _ in range(3):
            e2e_graph_step()
        ms_g = timed(e2e_graph_step, args.steps)
        torch.cuda.synchronize()
        e2e["graph_route"] = {"value": args.n_rand * world / (ms_g * 1e-3), "unit": "rays/s", "ms_per_step": ms_g,
                              "sync_every_step": {"value": args.n_rand * world / (ms_gs * 1e-3), "unit": "rays/s",
                                                  "ms_per_step": ms_gs},
                              "api": "dlnerf_b200.GraphedTrainStep(...)(host_rays, host_target_s, host_target_depth)"}

    # ---- the same step with the semantic head on, reported next to the headline (one GPU only) ---------------
    variants = None
    if not semK and world == 1 and not args.no_variants:
        K = 19                                         # KITTI-360 label set of fern_dsnerf.txt:55-56
        torch.manual_seed(3407)
        vc = dn.NeRF(D=COARSE_D, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True, semantic_num_classes=K).to(dev)
        vf = dn.NeRF(D=FINE_D, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True, semantic_num_classes=K).to(dev)
        v_sem = torch.randint(0, K, (n_rgb,), generator=torch.Generator().manual_seed(3407)).to(dev)
        vg = dn.GraphedTrainStep(H, W, FOCAL, args.n_rand, n_rgb, vc, vf, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE,
                                 perturb=1., raw_noise_std=1., depth_lambda=DEPTH_LAMBDA, depth_importance=1.,
                                 semantic_lambda=SEMANTIC_LAMBDA)
        v_step = lambda: vg(d_rays, d_tgt, d_dep, target_semantic=v_sem)      # noqa: E731
        for _ in range(3):
            v_step()
        ms_v = timed(v_step, args.steps)
        variants = {"semantic_head_19_classes": {
            "value": args.n_rand / (ms_v * 1e-3), "unit": "rays/s", "ms_per_step": ms_v,
            "what": "same step with semantic_loss = True as fern_dsnerf.txt:55-56 ships it: 19-class semantic_linear head on "
                    "both nets, cross-entropy of the fine and coarse per-ray logits (semantic_lambda 0.01), "
                    "GraphedTrainStep; `bench.py --semantic 19` gives the full line for it"}}
        del vg, vc, vf

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline()
    if rank == 0:
        print(json.dumps({
            "metric": "training rays/sec (fwd+bwd)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": scaling_kind(args), "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args),
            "value_route": {"graph": "dlnerf_b200.GraphedTrainStep (CUDA graph of train_step)", "fused": "dlnerf_b200.train_step", "dropin": "render()+loss.backward()"}[args.path],
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cb, "variants": variants, "kernels": kern,
            "kernel_times_from": kernel_times_from}))
    if world > 1:
        if graphed is not None:
            graphed.close()          # the graph holds NCCL kernels: release it before the process group
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--n-rand", type=int, default=4096, help="rays per step per GPU (config B: 4096)")
    ap.add_argument("--global-n-rand", type=int, default=0,
                    help="strong-scaling mode (config E, 16k-256k rays): rays per step of the WHOLE job, split evenly over "
                         "the GPUs; overrides --n-rand and reports \"scaling\": \"strong\"")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--coarse-sms", type=int, default=0, help="SMs of the coarse backward when it runs next to the fine "
                    "forward / dgrad (0: train_step's default schedule)")
    ap.add_argument("--fine-sms", type=int, default=0, help="SMs of the fine forward / dgrad in that schedule (0: all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the semantic-head variant of the step")
    ap.add_argument("--semantic", type=int, default=0,
                    help="classes of the semantic head (fern_dsnerf.txt:55 turns it on with the KITTI-360 label set, 19); "
                         "0 = the headline workload (RGB + depth loss)")
    ap.add_argument("--path", default="graph", choices=["graph", "fused", "dropin"],
                    help="route of the device-resident `value`: CUDA-graph replay of train_step, train_step (fused "
                         "loss), or render()+loss.backward(); `e2e` always uses the drop-in route")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write there behind Python's back (NCCL prints its version
    # banner to fd 1 whenever NCCL_DEBUG is set), so fd 1 is pointed at stderr for the whole run and the JSON line is
    # written to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()

"""HBM roofline of the non-MLP kernels (compositing, hierarchical sampling, stratified sampling) at
sizes where they are bandwidth- rather than launch-bound.  Algorithmic bytes per ray follow SURVEY.md §8(d)
(+4S where the kernel takes the N(0,1) / U(0,1) draws as an input tensor, which is how the product calls it).

    python tools/render_microbench.py [--rays 262144] [--reps 20] [--json out.json]

Inputs at 262 144 rays are 0.3-0.8 GB per tensor (>> 126 MB L2), so every launch streams from HBM."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlnerf_b200 as dn  # noqa: E402
from dlnerf_b200 import ops  # noqa: E402


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6539.9


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, nargs="+", default=[4096, 65536, 262144])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    peak = peak_gbs()
    rows = []
    for N in args.rays:
        rb = torch.zeros(N, 11, device=dev)
        rb[:, 0:3] = torch.randn(N, 3, device=dev) * 0.1
        rb[:, 3:6] = torch.randn(N, 3, device=dev)
        rb[:, 5] = 2.0
        rb[:, 7] = 1.0
        rays_d = rb[:, 3:6].contiguous()
        n_rgb = N // 2
        tgt = torch.rand(n_rgb, 3, device=dev)
        tdep = torch.rand(N - n_rgb, device=dev)
        sums = torch.zeros(2, device=dev)
        for S in (64, 128):
            raw = torch.randn(N, S, 4, device=dev)
            raw[..., 3] += 1.0
            t_rand = torch.rand(N, S, device=dev)
            noise = torch.randn(N, S, device=dev)
            z = ops.stratified_z(rb, S, t_rand)
            cases = [
                ("stratified_z", 8 * S + 8, lambda: ops.stratified_z(rb, S, t_rand)),
                ("composite_fwd", 28 * S + 36, lambda: ops.composite(raw, z, rays_d, noise, 1.0, False)),
                ("composite_bwd_fused_loss", 40 * S + 32,
                 lambda: ops.composite_bwd_fused_loss(raw, z, rays_d, noise, 1.0, False, tgt, tdep, None, n_rgb,
                                                      1e-3, 1e-3, 0, 1.0, sums)),
            ]
            if S == 64:
                w = ops.composite(raw, z, rays_d, noise, 1.0, False)[3]
                u = torch.rand(N, 64, device=dev)
                cases.append(("importance_resample", 248 + 256 + 256 + 256 + 512,
                              lambda: ops.importance_resample(z, w, 64, u)))
            for name, bpr, fn in cases:
                ms = timeit(fn, args.reps)
                gbs = bpr * N / ms / 1e6
                rows.append({"kernel": name, "rays": N, "S": S, "bytes_per_ray": bpr, "ms": round(ms, 4),
                             "GBps": round(gbs, 1), "frac_hbm": round(gbs / peak, 3)})
                print("%-26s N=%7d S=%3d  %8.4f ms  %7.1f GB/s  %5.1f%% of %.0f" % (name, N, S, ms, gbs,
                                                                                 100 * gbs / peak, peak), flush=True)
            del raw, t_rand, noise, z
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"peak_gbs": peak, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()

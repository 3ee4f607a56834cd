"""bench.py prints exactly ONE JSON line on stdout with the keys the driver reads.  The reference arm runs on the
CPU (here); the B200 arm is checked on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--n-rand", "512"], timeout=600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "training rays/sec (fwd+bwd)" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert cb["anomaly_detection"] is False and cb["anomaly_on"]["value"] > 0          # both settings
    assert d["variants"]["semantic_head_19_classes"]["value"] > 0                       # fern_dsnerf.txt:55 as shipped
    # the same command line gives the same `config` in both arms (checked key by key on the GPU box)
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    ns = argparse.Namespace(gpus=1, n_rand=512, global_n_rand=0, semantic=0)
    assert d["config"] == bench.workload_config(ns)
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_b200_arm_prints_one_contract_line():
    d = _run(["--steps", "3", "--warmup", "3", "--no-cpu-baseline"], timeout=900)
    assert (BASE_KEYS | {"clocks", "gpu_launches", "roofline", "kernels"}) <= set(d)
    assert d["metric"] == "training rays/sec (fwd+bwd)" and d["unit"] == "rays/s" and d["n_gpus"] == 1
    assert d["steps"] == 3 and d["warmup"] >= 3 and d["dtype"] == "bf16" and d["data"] == "synthetic"
    assert d["value"] > 1e5 and abs(d["value"] - 4096 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert d["gpu_launches"] > 0 and "workload" in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "rays/s" and e["h2d_bytes_per_step"] == (2 * 4096 * 3 + 2048 * 3 + 2048) * 4
    assert e["d2h_bytes_per_step"] == 4 and e["value"] <= 1.05 * d["value"]
    # the loss is read back every step either way; waiting for it after every step can only be slower
    assert 0 < e["sync_every_step"]["value"] <= 1.02 * e["value"] and "loss_read" in e
    r = d["roofline"]
    assert r["bound"] in ("tensor", "hbm") and r["unit"] in ("TFLOP/s", "GB/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.05 < r["frac"] < 1.0 and "traffic" in r
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and isinstance(c["reasons"], list)


def test_strong_scaling_mode_splits_the_global_batch():
    """`--global-n-rand G` (config E): G rays per step over the GPUs, reported as strong scaling, same config in both arms."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    ns = argparse.Namespace(gpus=8, n_rand=4096, global_n_rand=262144, semantic=0)
    assert bench.rays_per_gpu(ns) == 32768 and bench.scaling_kind(ns) == "strong"
    cfg = bench.workload_config(ns)
    assert cfg["n_rand_per_gpu"] == 32768 and cfg["global_rays_per_step"] == 262144 and "dp8" in cfg["parallelism"]
    ns = argparse.Namespace(gpus=2, n_rand=4096, global_n_rand=0, semantic=19)
    assert bench.rays_per_gpu(ns) == 4096 and bench.scaling_kind(ns) == "weak"
    assert bench.workload_config(ns)["global_rays_per_step"] == 8192 and "19 classes" in bench.workload_config(ns)["semantic_head"]

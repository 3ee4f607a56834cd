#!/bin/bash
# N-GPU evidence: the NCCL parity test (N >= 2), the weak-scaling bench line (4096 rays per GPU) and the strong-scaling
# config-E line (262 144 rays per step over the N GPUs)
N=${1:-2}
mkdir -p gpurun_out
[ -z "$SKIP_TEST" ] && timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -p no:cacheprovider > gpurun_out/multi_test_$N.log 2>&1; echo "multi test exit=$? $(tail -1 gpurun_out/multi_test_$N.log)"; grep "rel-L2\|max|diff|" gpurun_out/multi_test_$N.log
run() { local tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_${tag}_$N.json 2> gpurun_out/bench_${tag}_$N.err
  echo "$tag exit=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${tag}_$N.json")); print("  n_gpus %d scaling %s rays/step %d: %.0f rays/s %.3f ms/step e2e %.0f clocks %s %s" % (d["n_gpus"], d["scaling"], d["config"]["global_rays_per_step"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("  no json", e)
PY
}
run weak --steps 20 --warmup 5
run strongE --steps 5 --warmup 3 --global-n-rand 262144

#!/bin/bash
# GPU parity suite of the tree + the semantic-head step (graphed) for the A/B of a dgrad change
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit=$? $(tail -1 gpurun_out/pytest_gpu.log)" | tee -a gpurun_out/summary.txt
grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --semantic 19 > gpurun_out/bench_sem.json 2> gpurun_out/bench_sem.err; echo "bench_sem exit=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_sem.json"))
print("value %.0f rays/s  %.3f ms/step  e2e %.0f" % (d['value'], d['ms_per_step'], d['e2e']['value']))
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step'])[:8]:
    print("  %-28s %8.4f ms/step  x%.0f" % (k, v['ms_per_step'], v['launches_per_step']))
PY

#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("value %.0f rays/s  %.3f ms/step  e2e %.0f  launches %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
print("variants", d.get("variants"))
print("cpu", d.get("cpu_baseline"))
PY
timeout 900 python -m pytest tests/test_bench_contract.py tests/test_gpu_semantic.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3

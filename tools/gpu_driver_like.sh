#!/bin/bash
# what the driver runs at round end: the GPU suite in one process, smoke(), both bench arms with default flags
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit=$? $(tail -1 gpurun_out/pytest_gpu.log)" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench reference exit=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
for f in ("bench_ref.json", "bench.json"):
    try:
        d = json.load(open("gpurun_out/" + f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print("%s: value %.0f %s  %.3f ms/step  e2e %.0f  launches %s roofline %s" % (f, d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_launches'), {k: d['roofline'][k] for k in ('kernel','bound','frac')} if d.get('roofline') else None))
    print("   cpu_baseline", {k: v for k, v in (d.get('cpu_baseline') or {}).items() if k in ('value','cores','kind','sample')}, "same config keys:", sorted(d['config'].keys())[:4], "...")
PY

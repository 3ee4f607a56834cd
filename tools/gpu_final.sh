#!/bin/bash
# what the driver runs at round end (GPU parity suite, smoke, both bench arms) + the semantic-head bench lines
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit=$? $(tail -1 gpurun_out/pytest_gpu.log)" | tee -a gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench reference exit=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --semantic 19 > gpurun_out/bench_sem.json 2> gpurun_out/bench_sem.err; echo "bench_sem exit=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --semantic 19 --path dropin > gpurun_out/bench_sem_dropin.json 2> gpurun_out/bench_sem_dropin.err; echo "bench_sem_dropin exit=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
for f in ("bench_ref.json", "bench.json", "bench_sem.json", "bench_sem_dropin.json"):
    try:
        d = json.load(open("gpurun_out/" + f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print("%s: value %.0f %s  %.3f ms/step  e2e %.0f  launches %s clocks %s cpu %s" % (f, d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_launches'), d.get('clocks'), d.get('cpu_baseline')))
    for k, v in sorted(d.get('kernels', {}).items(), key=lambda kv: -kv[1]['ms_per_step'])[:12]:
        print("  %-28s %8.4f ms/step  x%.0f  %s" % (k, v['ms_per_step'], v['launches_per_step'], ("%.0f TF/s (%.1f%%)" % (v['tflops'], 100 * v['frac_of_sustained_peak'])) if 'tflops' in v else ''))
PY

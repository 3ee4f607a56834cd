#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { local name=$1; shift
  timeout 600 python -m pytest "$@" -q -s --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit=$? $(tail -1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt; }
run sem tests/test_gpu_semantic.py -m gpu
run optim tests/test_gpu_optim.py -m gpu
grep -E "FAILED|Error|error" gpurun_out/sem.log gpurun_out/optim.log | head -20
grep -E "every 50" gpurun_out/optim.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --semantic 19 --path dropin > gpurun_out/bench_sem_dropin.json 2> gpurun_out/bench_sem_dropin.err; echo "bench_sem_dropin exit=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --semantic 19 > gpurun_out/bench_sem.json 2> gpurun_out/bench_sem.err; echo "bench_sem exit=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
for f in ("bench_sem_dropin.json", "bench_sem.json"):
    try:
        d = json.load(open("gpurun_out/" + f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print("%s: value %.0f rays/s  %.3f ms/step  e2e %.0f  launches %s clocks %s" % (f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
    for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step'])[:26]:
        print("  %-28s %8.4f ms/step  x%.0f  %s" % (k, v['ms_per_step'], v['launches_per_step'], ("%.0f TF/s (%.1f%%)" % (v['tflops'], 100 * v['frac_of_sustained_peak'])) if 'tflops' in v else ''))
PY
bash tools/ncu_sem.sh

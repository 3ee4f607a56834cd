#!/bin/bash
# full trip: every -m gpu test file (one log each), then the default bench line
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { local name=$1; shift
  timeout 900 python -m pytest "$@" -q --tb=short -p no:cacheprovider -s > gpurun_out/$name.log 2>&1
  echo "$name exit=$? $(tail -1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt; }
for f in tests/test_gpu_*.py; do n=$(basename $f .py); run ${n#test_gpu_} $f -m gpu; done
if [ -z "$NO_BENCH" ]; then
timeout 600 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print("value %.0f rays/s  %.3f ms/step  e2e %.0f  launches %s clocks %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:8]:
    print("%-28s %8.3f ms/step  %s" % (k, v['ms_per_step'], ("%.0f TF/s (%.1f%%)"%(v['tflops'],100*v['frac_of_sustained_peak'])) if 'tflops' in v else ''))
PY
fi

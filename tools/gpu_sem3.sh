#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_semantic.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/sem.log 2>&1; echo "sem exit=$? $(tail -1 gpurun_out/sem.log)" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --semantic 19 > gpurun_out/bench_sem.json 2> gpurun_out/bench_sem.err; echo "bench_sem exit=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_sem.json"))
print("value %.0f rays/s  %.3f ms/step  e2e %.0f" % (d['value'], d['ms_per_step'], d['e2e']['value']))
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step'])[:14]:
    print("  %-28s %8.4f ms/step  x%.0f" % (k, v['ms_per_step'], v['launches_per_step']))
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:sem_head_fwd -c 4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --path fused --semantic 19 2>&1 | grep -E "sem_head_fwd|gpu__time|dram__bytes" | head -20

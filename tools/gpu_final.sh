#!/bin/bash
# what the driver runs at round end (GPU parity suite, smoke, both bench arms) + the semantic-head bench lines and the
# ncu capture of the bulk-staged sem_head_fwd kernel
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
    print("%s: value %.0f %s  %.3f ms/step  e2e %.0f  launches %s variants %s" % (f, d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_launches'), {k: round(v['value']) for k, v in (d.get('variants') or {}).items()}))
    for k, v in sorted(d.get('kernels', {}).items(), key=lambda kv: -kv[1]['ms_per_step'])[:10]:
        print("  %-28s %8.4f ms/step  x%.0f  %s" % (k, v['ms_per_step'], v['launches_per_step'], ("%.0f TF/s (%.1f%%)" % (v['tflops'], 100 * v['frac_of_sustained_peak'])) if 'tflops' in v else ''))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --path fused --semantic 19"
ncu --set full --clock-control none --import-source on -k regex:"sem_head_fwd" -s 4 -c 2 -o gpurun_out/prof_sem_bulk -f $CMD > gpurun_out/ncu_sem_bulk.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_sem_bulk.ncu-rep gpurun_out/sem_head_fwd_bulk_ncu.txt > /dev/null 2>&1; echo "ncu exit=$?"
grep -E "^###|gpu__time_duration|dram__bytes_read|gpu__dram_throughput" gpurun_out/sem_head_fwd_bulk_ncu.txt

#!/bin/bash
# config E on one GPU: rays per step 16 384 / 65 536 / 262 144 (ray-chunked above the stash budget)
mkdir -p gpurun_out; : > gpurun_out/summary.txt
for n in 16384 65536 262144; do
  steps=10; [ $n -ge 262144 ] && steps=4
  timeout 900 python bench.py --n-rand $n --steps $steps --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; echo "bench $n exit=$?" | tee -a gpurun_out/summary.txt
  tail -2 gpurun_out/bench_$n.err
done
python - <<'PY'
import json
for n in (16384, 65536, 262144):
    try:
        d = json.load(open("gpurun_out/bench_%d.json" % n))
        print(n, "%.0f rays/s" % d["value"], "%.2f ms/step" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["clocks"]["reasons"], "mlp frac %.3f" % d["roofline"]["all_mlp_kernels"]["frac"])
    except Exception as e:
        print(n, "no json", e)
PY

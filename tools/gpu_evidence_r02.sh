#!/bin/bash
# round-2 single-GPU evidence lines: config E sizes, semantic-head variant (graph + drop-in), one-tile kernel A/B
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("  %s: %.0f rays/s  %.3f ms/step  e2e %.0f  clocks %s %s" % (sys.argv[1].split('/')[-1], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
PY
}
for n in 16384 65536 262144; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants --n-rand $n > gpurun_out/r02_bench_nrand$n.json 2> gpurun_out/e.err && line gpurun_out/r02_bench_nrand$n.json
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --semantic 19 > gpurun_out/r02_bench_semantic19.json 2> gpurun_out/e.err && line gpurun_out/r02_bench_semantic19.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --semantic 19 --path dropin > gpurun_out/r02_bench_semantic19_dropin.json 2> gpurun_out/e.err && line gpurun_out/r02_bench_semantic19_dropin.json
DLN_CHAIN=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants > gpurun_out/r02_bench_onetile_kernel.json 2> gpurun_out/e.err && line gpurun_out/r02_bench_onetile_kernel.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants > gpurun_out/r02_bench_pair_kernel.json 2> gpurun_out/e.err && line gpurun_out/r02_bench_pair_kernel.json

#!/bin/bash
# round-2 profile evidence: (1) launch list of a bench step (gpu__time_duration per launch), (2) one --set full capture of
# the six MLP launches of a step (chain2 fwd/dgrad for both nets, wgrad for both nets)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants --path fused"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1; echo "launch list exit=$?"
ncu --set full --clock-control none --import-source on -k regex:"chain2_kernel|wgrad_kernel" -s 18 -c 6 -f -o gpurun_out/prof_r02_mlp $CMD > gpurun_out/ncu_mlp.log 2>&1; echo "mlp capture exit=$?"
python tools/ncu_summary.py gpurun_out/prof_r02_mlp.ncu-rep gpurun_out/r02_mlp_ncu_full.txt > /dev/null 2>&1
grep -E "^###|gpu__time_duration|dram__bytes|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram_throughput" gpurun_out/r02_mlp_ncu_full.txt

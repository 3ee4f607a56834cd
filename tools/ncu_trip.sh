#!/bin/bash
# ncu evidence: launch list of one bench run + full captures of the chain and wgrad kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants --path fused"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 12 -c 4 -o gpurun_out/prof_chain $CMD > gpurun_out/ncu_chain.log 2>&1
echo "chain capture exit=$?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 6 -c 2 -o gpurun_out/prof_wgrad $CMD > gpurun_out/ncu_wgrad.log 2>&1
echo "wgrad capture exit=$?"
ls -la gpurun_out/

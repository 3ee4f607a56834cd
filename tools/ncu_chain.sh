#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants --path fused"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 13 -c 2 -o gpurun_out/prof_chain2 $CMD > gpurun_out/ncu_chain.log 2>&1
echo "chain capture exit=$?"

#!/bin/bash
# one ncu --set full capture of the CTA-pair chain kernel (forward with stash by default; BWD=1 for the dgrad chain)
mkdir -p gpurun_out
CMD="python tools/trace_chain.py"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain2 -s 1 -c 1 -f -o gpurun_out/prof_c2${BWD:+_bwd} $CMD > gpurun_out/ncu_c2.log 2>&1
echo "chain2 capture exit=$?"; tail -3 gpurun_out/ncu_c2.log

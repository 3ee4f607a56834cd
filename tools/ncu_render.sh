#!/bin/bash
# ncu capture of the HBM-bound render kernels at 262 144 rays (run via gpurun; one GPU).  KREGEX / COUNT narrow it.
mkdir -p gpurun_out
KREGEX=${KREGEX:-composite|resample|stratified}
COUNT=${COUNT:-40}
SKIP=${SKIP:-3}
CMD="python tools/render_microbench.py --rays 262144 --reps 1"
$CMD > gpurun_out/render_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" --launch-skip $SKIP --launch-count $COUNT \
    -o gpurun_out/prof_render -f $CMD > gpurun_out/ncu_render.log 2>&1
echo "render capture exit=$?"

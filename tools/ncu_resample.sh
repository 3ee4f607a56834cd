#!/bin/bash
# ncu capture of importance_resample at 262 144 rays + text summaries (run via gpurun; one GPU)
KREGEX=resample COUNT=1 SKIP=2 bash tools/ncu_render.sh
ncu -i gpurun_out/prof_render.ncu-rep --page raw --csv > gpurun_out/resample_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_render.ncu-rep --page details > gpurun_out/resample_details.txt 2>/dev/null
ncu -i gpurun_out/prof_render.ncu-rep --page source --csv > gpurun_out/resample_source.csv 2>/dev/null
tail -1 gpurun_out/ncu_render.log

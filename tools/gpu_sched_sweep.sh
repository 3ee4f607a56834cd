#!/bin/bash
# sweep of the SM partition between the fine net's write-bound kernels and the early coarse backward
mkdir -p gpurun_out
for cfg in "0 0" "28 120" "36 112" "44 104" "52 96" "36 0" "44 0"; do set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants --coarse-sms $1 --fine-sms $2 > gpurun_out/sched_$1_$2.json 2> gpurun_out/sched.err || { echo "cfg $cfg failed"; tail -3 gpurun_out/sched.err; continue; }
  python -c "
import json; d=json.load(open('gpurun_out/sched_$1_$2.json')); print('coarse_sms $1 fine_sms $2: %.3f ms/step  %.0f rays/s  e2e graph %.0f' % (d['ms_per_step'], d['value'], d['e2e']['graph_route']['value']))"
done

#!/bin/bash
# ncu evidence for the semantic-head step (bench.py --semantic 19, eager fused route): launch list of one run +
# full captures of the head kernels and of the dgrad chain instantiation that adds the per-ray semantic row.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --path fused --semantic 19"
$CMD > gpurun_out/plain_sem.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 500 --csv --log-file gpurun_out/launches_sem.csv $CMD > gpurun_out/ncu_launches_sem.log 2>&1
echo "launch list exit=$?"
$CMD > gpurun_out/plain_sem2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sem_head|sem_ce|sem_unfold|sem_fold" -s 20 -c 12 -o gpurun_out/prof_sem -f $CMD > gpurun_out/ncu_sem.log 2>&1
echo "sem capture exit=$?"
python tools/ncu_summary.py gpurun_out/prof_sem.ncu-rep gpurun_out/sem_ncu_full.txt > /dev/null 2>&1
ls -la gpurun_out/ | grep -i sem

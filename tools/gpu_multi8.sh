#!/bin/bash
# 8-GPU trip: config E (32 768 rays per GPU = 262 144 rays per step) and the headline size, ray-sharded
mkdir -p gpurun_out; : > gpurun_out/summary.txt
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --n-rand 32768 --steps 10 --warmup 3 > gpurun_out/bench${N}_nrand32768.json 2> gpurun_out/bench${N}_nrand32768.err
echo "bench$N n_rand 32768 exit=$?" | tee -a gpurun_out/summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench${N}.json 2> gpurun_out/bench${N}.err
echo "bench$N exit=$?" | tee -a gpurun_out/summary.txt
grep -E "Error|error|Traceback" -A3 gpurun_out/bench${N}_nrand32768.err gpurun_out/bench${N}.err | tail -12
python - <<PY
import json
for f in ("bench${N}_nrand32768.json", "bench${N}.json"):
    try:
        d = json.load(open("gpurun_out/" + f)); print(f, d["n_gpus"], "%.0f rays/s" % d["value"], "%.3f ms" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["clocks"])
    except Exception as e:
        print(f, "no json", e)
PY

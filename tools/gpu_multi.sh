#!/bin/bash
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench$N.json 2> gpurun_out/bench$N.err
echo "exit=$?"; ls -la gpurun_out/bench$N.json; grep -E "Error|error|Traceback" -A3 gpurun_out/bench$N.err | tail -20
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench$N.json")); print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
except Exception as e: print("no json", e)
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-200

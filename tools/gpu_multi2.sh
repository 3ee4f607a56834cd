#!/bin/bash
# 2-GPU trip: semantic tests (new forward kernel), then torchrun bench lines with the head off and on
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_semantic.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/sem.log 2>&1; echo "sem exit=$? $(tail -1 gpurun_out/sem.log)" | tee -a gpurun_out/summary.txt
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench$N.json 2> gpurun_out/bench$N.err
echo "bench$N exit=$?" | tee -a gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --semantic 19 > gpurun_out/bench${N}_sem.json 2> gpurun_out/bench${N}_sem.err
echo "bench${N}_sem exit=$?" | tee -a gpurun_out/summary.txt
grep -E "Error|error|Traceback" -A3 gpurun_out/bench$N.err gpurun_out/bench${N}_sem.err | tail -20
python - <<'PY'
import json
for f in ("bench2.json", "bench2_sem.json"):
    try:
        d = json.load(open("gpurun_out/" + f)); print(f, d["n_gpus"], "%.0f rays/s" % d["value"], "%.3f ms" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["clocks"])
        for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step'])[6:10]:
            print("  %-28s %8.4f ms/step  x%.0f" % (k, v['ms_per_step'], v['launches_per_step']))
    except Exception as e:
        print(f, "no json", e)
PY

#!/bin/bash
# chain2 bring-up: MLP parity, microbench, timeline
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -m gpu -x -q --tb=short -p no:cacheprovider -s > gpurun_out/c2_mlp.log 2>&1; echo "mlp(chain2) exit=$? $(tail -1 gpurun_out/c2_mlp.log)"
[ -n "$UB1" ] && { DLN_CHAIN=1 SPLITS=11 timeout 200 python tools/mlp_microbench.py > gpurun_out/c2_ub1.log 2>&1; echo "ub chain1: $(cat gpurun_out/c2_ub1.log | tail -2)"; }
SPLITS=11 timeout 200 python tools/mlp_microbench.py > gpurun_out/c2_ub2.log 2>&1; echo "ub chain2: $(cat gpurun_out/c2_ub2.log | tail -2)"
timeout 100 python tools/trace_chain.py; [ -n "$TR_ALL" ] && { KEEP=0 timeout 100 python tools/trace_chain.py; BWD=1 timeout 100 python tools/trace_chain.py; }

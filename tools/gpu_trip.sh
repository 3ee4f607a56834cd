#!/bin/bash
# Run the GPU parity suite piecewise (one process per group so a faulting kernel cannot take the rest
# of the run with it); logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, pytest args...
  local name=$1; shift
  timeout 600 python -m pytest "$@" -q -s --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/summary.txt
  tail -3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run render tests/test_gpu_render_kernels.py -m gpu
run mlp_fwd tests/test_gpu_mlp.py -m gpu -k "forward"
run mlp_bwd_dz tests/test_gpu_mlp.py -m gpu -k "dz_per_layer"
run mlp_bwd tests/test_gpu_mlp.py -m gpu -k "backward_gradients or full_size"
run mlp_misc tests/test_gpu_mlp.py -m gpu -k "repacked or unsupported"
run e2e tests/test_gpu_render_e2e.py -m gpu
run fullsize tests/test_gpu_fullsize_properties.py -m gpu
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/smoke.log | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
cut -c1-400 gpurun_out/bench.json | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/bench.err

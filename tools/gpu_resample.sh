#!/bin/bash
# importance_resample (warp-per-ray and thread-per-ray kernels): parity tests, then the HBM microbench for both
set -x
python -m pytest tests/test_gpu_render_kernels.py tests/test_gpu_rng.py tests/test_gpu_fullsize_properties.py -x -q -m gpu > gpurun_out/resample_tests.log 2>&1
echo "tests exit=$?" >> gpurun_out/resample_tests.log
tail -5 gpurun_out/resample_tests.log
python tools/render_microbench.py --json gpurun_out/r02_render_ubench.json > gpurun_out/ubench_thread.log 2>&1
DLN_RESAMPLE=warp python tools/render_microbench.py --json gpurun_out/r02_render_ubench_warp.json > gpurun_out/ubench_warp.log 2>&1
DLN_RESAMPLE=thread python tools/render_microbench.py --rays 4096 16384 32768 --json gpurun_out/r02_render_ubench_thread_small.json 2>&1 | grep resample
DLN_RESAMPLE=warp python tools/render_microbench.py --rays 16384 32768 2>&1 | grep resample
grep -h resample gpurun_out/ubench_thread.log gpurun_out/ubench_warp.log

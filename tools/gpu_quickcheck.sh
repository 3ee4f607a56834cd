#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit=$? $(tail -1 gpurun_out/pytest_gpu.log)" | tee -a gpurun_out/summary.txt
grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head

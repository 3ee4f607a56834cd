#!/bin/bash
# A/B of importance_resample builds: variants/lib_rs_*.so (DLN_SO_PATH) through the HBM microbench at 65 k / 262 k rays
for f in variants/lib_rs_*.so; do
  echo "== $f"
  DLN_SO_PATH=$PWD/$f python tools/render_microbench.py --rays 65536 262144 --reps 30 2>&1 | grep resample
done

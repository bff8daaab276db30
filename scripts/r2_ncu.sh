#!/bin/bash
# ncu --set full of selected kernels inside a torch-free run (scripts/kbench.py).
# usage: KREGEX='moloch_(a|b)' SKIP=n COUNT=m bash scripts/r2_ncu.sh TAG
TAG=${1:-ncu}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on \
  --kernel-name "regex:${KREGEX}" --launch-skip ${SKIP:-0} --launch-count ${COUNT:-3} -f -o gpurun_out/${TAG} \
  env ${EXTRA_ENV} python scripts/kbench.py --steps 1 --warmup 1 > gpurun_out/${TAG}.log 2>&1
tail -3 gpurun_out/${TAG}.log
ls -la gpurun_out/${TAG}.ncu-rep

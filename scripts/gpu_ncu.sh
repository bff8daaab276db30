#!/bin/bash
# ncu --set full of selected kernels of one bench step.
# usage: KREGEX='moloch_(a|b)' SKIP=n COUNT=m bash scripts/gpu_ncu.sh TAG [pytest]
TAG=${1:-run}
mkdir -p gpurun_out
if [ "$2" = "pytest" ]; then
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
  tail -12 gpurun_out/${TAG}_pytest.log
fi
timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name "regex:${KREGEX}" --launch-skip ${SKIP:-0} --launch-count ${COUNT:-2} -f -o gpurun_out/${TAG}_full \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
tail -3 gpurun_out/${TAG}_full.log

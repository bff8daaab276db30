#!/bin/bash
# Last GPU call of round 2 (final tree, ~2 GPU-minutes left): ncu --set full of one launch of the two WAF kernels and
# of status_update (torch-free kbench), then the ncu launch list of the bench command.  Digest is made off the box.
# usage: gpurun --timeout 125 -- bash scripts/r2_last.sh TAG
TAG=${1:-r2z}
mkdir -p gpurun_out
MOLOCH_B200_GRAPH=0 timeout 50 ncu --set full --clock-control none --import-source on \
  --kernel-name 'regex:moloch_(waf_horizontal|waf_vertical2|status_update)' --launch-skip 7 --launch-count 3 \
  -f -o gpurun_out/${TAG}_full python scripts/kbench.py --steps 1 --warmup 1 > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out/${TAG}_full.ncu-rep
MOLOCH_B200_WSOLVE=12 timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity \
  > gpurun_out/${TAG}_launches.log 2>&1
tail -c 300 gpurun_out/${TAG}_launches.log

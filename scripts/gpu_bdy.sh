#!/bin/bash
# Boundary / mkslice / TKE / spectral-nudging kernels at full workload size:
# per-kernel device timing (library CUDA events) and one ncu --set full capture.
# usage: gpurun --timeout 150 -- bash scripts/gpu_bdy.sh TAG
TAG=${1:-bdy}
mkdir -p gpurun_out
timeout 70 python scripts/kbench.py --boundary --slice --spectral --tke --steps 4 --warmup 1 \
  > gpurun_out/${TAG}_kbench.json 2> gpurun_out/${TAG}_kbench.err
tail -c 2500 gpurun_out/${TAG}_kbench.json; tail -3 gpurun_out/${TAG}_kbench.err
timeout 80 ncu --set full --clock-control none \
  --kernel-name 'regex:moloch_(bdy|mkslice|spec|chem|zstagtoh|htozstag|tke)' --launch-count 30 -f \
  -o gpurun_out/${TAG}_full python scripts/kbench.py --boundary --slice --spectral --tke --steps 1 --warmup 0 \
  > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log
ls -la gpurun_out | tail -5

#!/bin/bash
# One-GPU round: parity tests, bench line, ncu launch list, ncu --set full of one
# launch of every hot kernel.  Usage: gpurun --timeout 1500 -- bash scripts/gpu_profile.sh TAG
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline \
  > gpurun_out/${TAG}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name 'regex:moloch_(waf_horizontal|waf_vertical2|wsolve|sound_pre|uvupdate|divdamp_filter|destagger|restagger)' \
  --launch-skip ${SKIP:-66} --launch-count ${COUNT:-8} -f -o gpurun_out/${TAG}_full \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
tail -3 gpurun_out/${TAG}_full.log
ls -la gpurun_out

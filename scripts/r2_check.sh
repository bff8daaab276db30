#!/bin/bash
# First GPU call of a new round: everything that was added after round 1's GPU budget ran out.
# usage: gpurun --timeout 300 -- bash scripts/r2_check.sh        (1 GPU)
#        gpurun --gpus 2 --timeout 300 -- 'python -m pytest tests/test_gpu_multi.py -q -k "boundary or 2x1 or 1x2"'
mkdir -p gpurun_out
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 120 python -m pytest tests -m gpu -q --timeout 100 -p no:cacheprovider --durations=5 > gpurun_out/r2_gpu_tests.log 2>&1
tail -12 gpurun_out/r2_gpu_tests.log
timeout 90 python scripts/kbench.py --boundary --slice --spectral --tke --massck --diag --steps 4 --warmup 1 \
  > gpurun_out/r2_kbench_all.json 2> gpurun_out/r2_kbench_all.err
tail -c 1500 gpurun_out/r2_kbench_all.json; tail -3 gpurun_out/r2_kbench_all.err

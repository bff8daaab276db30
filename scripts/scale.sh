#!/bin/bash
# multi-GPU validation + scaling sweep (run under gpurun --gpus 8)
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x -k "2x2 or 2x4" 2>&1 | tail -5
for n in 8 4 2 1; do
  for t in p2p nccl; do
    if [ $n -eq 1 ] && [ $t = nccl ]; then continue; fi
    if [ $n -eq 1 ]; then
      timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/scale_${n}_${t}.json 2> gpurun_out/scale_${n}_${t}.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --transport $t --no-e2e > gpurun_out/scale_${n}_${t}.json 2> gpurun_out/scale_${n}_${t}.err
    fi
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${n}_${t}.json").read().strip().splitlines()[-1])
    print("N=${n} ${t}", d["config"]["decomposition"], "%.3e" % d["value"], "%.3f ms" % d["ms_per_step"], d["gpu_launches"], d["finite"])
    ks = {k["kernel"]: (round(k["avg_ms"]*1e3,1), k["launches_per_step"]) for k in d["kernels"]}
    print("   ", ks)
except Exception as e:
    print("N=${n} ${t} FAILED", e)
PY
  done
done

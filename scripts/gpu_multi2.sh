#!/bin/bash
# 2-GPU box: all parity tests that fit (single GPU + 2-rank decomposed), then N=2 bench A/B.
T=${TAG:-m2}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -6 gpurun_out/${T}_pytest.log
n=0
for envs in "" "$@"; do
  out=gpurun_out/${T}_ab${n}.json
  env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port $((29600+n)) bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $out 2> gpurun_out/${T}_ab${n}.err
  python - "$out" "$envs" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("[%s] %s %.3f ms/step %.3e c-u/s launches %d finite=%s" % (sys.argv[2], d["config"]["decomposition"], d["ms_per_step"], d["value"], d["gpu_launches"], d["finite"]))
    print("   ", {k["kernel"]: (round(k["avg_ms"]*1e3,1), k["launches_per_step"]) for k in d["kernels"]})
except Exception as e:
    print("[%s] FAILED %s" % (sys.argv[2], e))
PY
  n=$((n+1))
done
tail -3 gpurun_out/${T}_ab0.err

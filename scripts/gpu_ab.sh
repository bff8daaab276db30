#!/bin/bash
# parity tests + A/B bench of kernel variants selected by environment variables.
# usage: bash scripts/gpu_ab.sh TAG "ENV1=.. ENV2=.." "ENV3=.." ...
TAG=$1; shift
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_pytest.log
n=0
for envs in "" "$@"; do
  out=gpurun_out/${TAG}_ab${n}.json
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out 2> gpurun_out/${TAG}_ab${n}.err
  python - "$out" "$envs" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("[%s] %.3f ms/step %.3e c-u/s finite=%s" % (sys.argv[2], d["ms_per_step"], d["value"], d["finite"]))
    print("   ", {k["kernel"]: round(k["avg_ms"]*1e3,1) for k in d["kernels"]})
except Exception as e:
    print("[%s] FAILED %s" % (sys.argv[2], e))
PY
  n=$((n+1))
done

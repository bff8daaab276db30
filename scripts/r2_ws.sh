#!/bin/bash
# wsolve variants one by one (torch-free per-kernel timer); usage: bash scripts/r2_ws.sh TAG v1 v2 ...
T=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  MOLOCH_B200_WSOLVE=$v timeout 200 python scripts/kbench.py --steps 6 --warmup 2 > gpurun_out/${T}_ws$v.json 2> gpurun_out/${T}_ws$v.err
  python - "$v" "gpurun_out/${T}_ws$v.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    ks = d["kernels"]
    print("wsolve %-3s %.3f ms/step  wsolve=%.1f us  finite=%s" % (sys.argv[1], d["ms_per_step"], ks["wsolve"]["avg_ms"] * 1e3, d["finite"]))
except Exception as exc:
    print(sys.argv[1], "FAILED", exc, open(sys.argv[2].replace(".json", ".err")).read()[-400:])
PY
done

#!/bin/bash
# A/B of library variants (regcm_b200/variants/*.so built by scripts/build_variants.py) with the torch-free
# per-kernel timer.  usage: bash scripts/r2_ab.sh TAG [kbench args]
T=${1:-ab}; shift
mkdir -p gpurun_out
for lib in regcm_b200/libmoloch_b200.so regcm_b200/variants/*.so; do
  [ -f "$lib" ] || continue
  n=$(basename $lib .so)
  MOLOCH_B200_LIB=$PWD/$lib timeout 200 python scripts/kbench.py --steps 6 --warmup 2 "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err
  python - "$n" "gpurun_out/${T}_$n.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    ks = d["kernels"]
    print("%-28s %.3f ms/step " % (sys.argv[1], d["ms_per_step"]) + " ".join("%s=%.1f" % (k, ks[k]["avg_ms"] * 1e3) for k in ("sound_pre", "uvupdate", "wsolve", "waf_horizontal", "waf_vertical", "destagger", "restagger", "status_update") if k in ks))
except Exception as exc:
    print(sys.argv[1], "FAILED", exc)
PY
done

#!/bin/bash
# One short A/B of library builds (regcm_b200/variants/*.so next to the default) on one GPU: per-kernel times on the
# full cordex25 grid and, on a 128x128 crop, the wrap-around sums of the prognostic fields' bit patterns after 3 steps
# (equal digests = the builds computed the same bits).  usage: gpurun --timeout 60 -- bash scripts/r2_ab_last.sh TAG
T=${1:-abz}
mkdir -p gpurun_out
run() {  # name lib
  MOLOCH_B200_LIB=$PWD/$2 timeout 15 python scripts/kbench.py --steps 4 --warmup 1 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
}
dig() {
  MOLOCH_B200_LIB=$PWD/$2 timeout 15 python scripts/kbench.py --crop 128 --steps 2 --warmup 1 --digest > gpurun_out/${T}_dig_$1.json 2> gpurun_out/${T}_dig_$1.err
}
run base regcm_b200/variants/base.so
run new regcm_b200/libmoloch_b200.so
dig base regcm_b200/variants/base.so
dig new regcm_b200/libmoloch_b200.so
run pf2 regcm_b200/variants/pf2.so
dig pf2 regcm_b200/variants/pf2.so
run hpf2 regcm_b200/variants/hpf2.so
run vpf2 regcm_b200/variants/vpf2.so
run base2 regcm_b200/variants/base.so
run new2 regcm_b200/libmoloch_b200.so
python - $T <<'PY'
import json, sys, glob
T = sys.argv[1]
for n in ("base", "new", "pf2", "hpf2", "vpf2", "base2", "new2"):
    try:
        d = json.loads(open(f"gpurun_out/{T}_{n}.json").read().strip().splitlines()[-1]); ks = d["kernels"]
        print("%-6s %.3f ms/step " % (n, d["ms_per_step"]) + " ".join("%s=%.1f" % (k, ks[k]["avg_ms"] * 1e3) for k in ("waf_horizontal", "waf_vertical", "wsolve", "sound_pre", "uvupdate", "status_update")))
    except Exception as exc:
        print(n, "FAILED", exc)
dg = {}
for n in ("base", "new", "pf2"):
    try:
        dg[n] = json.loads(open(f"gpurun_out/{T}_dig_{n}.json").read().strip().splitlines()[-1])["digest"]
    except Exception as exc:
        print("digest", n, "FAILED", exc)
for n in dg:
    print("digest", n, "== base:", dg[n] == dg.get("base"))
PY

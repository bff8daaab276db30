#!/bin/bash
# ncu --set full of the sound-loop kernels on the per-rank grid of the 8-GPU run (400 x 50 columns), one GPU.
T=${1:-r2small}
mkdir -p gpurun_out
for v in 12 9; do
MOLOCH_B200_GRAPH=0 MOLOCH_B200_WSOLVE=$v timeout 300 ncu --set full --clock-control none --import-source on \
  --kernel-name 'regex:moloch_(wsolve_tm|wsolve8|sound_div|uvupdate2)' --launch-skip 36 --launch-count 3 -f -o gpurun_out/${T}_ws$v \
  python scripts/kbench.py --jx 400 --iy 50 --steps 1 --warmup 1 > gpurun_out/${T}_ws$v.log 2>&1
python scripts/ncu_digest.py gpurun_out/${T}_ws$v.ncu-rep > gpurun_out/${T}_ws${v}_digest.txt 2>&1
done
for v in 12 9 5 2; do
MOLOCH_B200_WSOLVE=$v timeout 100 python scripts/kbench.py --jx 400 --iy 50 --steps 20 --warmup 3 > gpurun_out/${T}_k_ws$v.json 2>/dev/null
python - gpurun_out/${T}_k_ws$v.json $v <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("wsolve", sys.argv[2], "%.4f ms/step" % d["ms_per_step"], {k: round(v["avg_ms"] * 1e3, 1) for k, v in d["kernels"].items() if k in ("wsolve", "sound_pre", "uvupdate")})
PY
done
du -sh gpurun_out

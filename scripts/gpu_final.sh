#!/bin/bash
# last GPU call of the round: smoke, all GPU tests, boundary kernels timed at full size, ncu on a cropped domain
mkdir -p gpurun_out
timeout 15 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 30 python -m pytest tests -m gpu -q --timeout 60 -p no:cacheprovider 2>&1 | tail -3
timeout 35 python scripts/kbench.py --boundary --slice --spectral --tke --steps 3 --warmup 1 > gpurun_out/r1_bdy2_kbench.json 2> gpurun_out/r1_bdy2_kbench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r1_bdy2_kbench.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"],3), {k:v["avg_ms"] for k,v in d["kernels"].items() if k in ("bdy_relax","bdy_finish","mkslice","bdyval","spectral_nudge","tke","massck")})
except Exception as e: print("kbench failed", e)
PY
timeout 30 ncu --set full --clock-control none --kernel-name 'regex:moloch_(bdy_relax|bdy_finish|mkslice)' --launch-count 4 -f \
  -o gpurun_out/r1_bdy2_full python scripts/kbench.py --crop 192 --boundary --slice --steps 1 --warmup 0 > gpurun_out/r1_bdy2_full.log 2>&1
tail -2 gpurun_out/r1_bdy2_full.log

#!/bin/bash
# e2e leg: hand-off tests, then bench with several slab counts.
mkdir -p gpurun_out
T=${TAG:-r2e}
timeout 200 python -m pytest tests/test_gpu_zz_handoff.py -q -x -p no:cacheprovider 2>&1 | tail -3
for n in ${SLABS:-8 16 32}; do
  BENCH_HANDOFF_SLABS=$n timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/${T}_slabs$n.json 2> gpurun_out/${T}_slabs$n.err
  python - "$n" "gpurun_out/${T}_slabs$n.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    e = d["e2e"]
    print("slabs", sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "e2e ms/step %.2f" % e["ms_per_step"], e["handoff"], e["ms_per_step_by_handoff"], e.get("handoff_note"))
except Exception as exc:
    print("slabs", sys.argv[1], "no result:", exc)
PY
done

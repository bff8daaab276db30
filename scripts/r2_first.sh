#!/bin/bash
# Round 2, first GPU call (1 GPU): the whole -m gpu suite, the default bench line, the wsolve variants
# forced one by one, config 2 (isc24_small).   usage: gpurun --timeout 900 -- bash scripts/r2_first.sh
mkdir -p gpurun_out
T=${TAG:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
( time timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
( time timeout 400 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider --durations=8 ) > gpurun_out/${T}_gpu_tests.log 2>&1
tail -14 gpurun_out/${T}_gpu_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 600 gpurun_out/${T}_bench.err
for v in 6 7 5; do
  MOLOCH_B200_WSOLVE=$v timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_wsolve$v.json 2>/dev/null
done
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --workload isc24_small > gpurun_out/${T}_bench_isc24_small.json 2> gpurun_out/${T}_bench_isc24_small.err
python - <<'PY'
import json, glob, os
T = os.environ.get("TAG", "r2a")
for f in sorted(glob.glob(f"gpurun_out/{T}_bench*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        ks = {k["kernel"]: (round(k["avg_ms"] * 1e3, 1), k["launches_per_step"]) for k in d["kernels"]}
        print(os.path.basename(f), "ms/step %.3f" % d["ms_per_step"], "wsolve", d["config"]["wsolve_variant"],
              d["config"].get("variant_tuning"), "e2e", d.get("e2e") and (round(d["e2e"]["ms_per_step"], 2), d["e2e"]["ms_per_step_by_handoff"]))
        print("   ", ks)
    except Exception as exc:
        print(f, "no result:", exc)
PY

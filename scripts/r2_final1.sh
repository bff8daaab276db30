#!/bin/bash
# Final one-GPU check of a round: the whole GPU suite, smoke(), both bench arms.
T=${1:-r2final}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider ) > gpurun_out/${T}_gpu_tests.log 2>&1
tail -4 gpurun_out/${T}_gpu_tests.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
python - $T <<'PY'
import json, sys
T = sys.argv[1]
d = json.loads([l for l in open(f"gpurun_out/{T}_bench.json") if l.startswith("{")][-1])
print("ms/step %.3f value %.4e" % (d["ms_per_step"], d["value"]), "wsolve", d["config"]["wsolve_variant"], "parity", d["parity"]["bit_exact"],
      "e2e %.2f ms" % d["e2e"]["ms_per_step"], d["e2e"].get("link_gbs"), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"])
r = json.loads([l for l in open(f"gpurun_out/{T}_bench_ref.json") if l.startswith("{")][-1])
print("reference arm: %.4e" % r["value"], r["cpu_baseline"]["cores"], "cores;", r["cpu_baseline"]["sample"])
PY

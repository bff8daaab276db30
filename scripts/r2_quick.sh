#!/bin/bash
# One-GPU iteration: [GPU tests] + bench line + kernel table.  TAG names the outputs, TESTS=0 skips pytest,
# BENCH_ARGS adds bench.py flags, EXTRA_ENV="A=1 B=2" adds environment.
mkdir -p gpurun_out
T=${TAG:-q}
if [ "${TESTS:-1}" != "0" ]; then
  ( time timeout 500 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider ${TESTS_ARGS} ) > gpurun_out/${T}_gpu_tests.log 2>&1
  tail -6 gpurun_out/${T}_gpu_tests.log
fi
env ${EXTRA_ENV} timeout 400 python bench.py --steps ${STEPS:-20} --warmup 5 ${BENCH_ARGS} > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.err
python - <<'PY'
import json, os
T = os.environ.get("TAG", "q")
try:
    d = json.loads([l for l in open(f"gpurun_out/{T}_bench.json") if l.startswith("{")][-1])
    print("ms/step %.3f" % d["ms_per_step"], "value %.4e" % d["value"], "wsolve", d["config"]["wsolve_variant"], d["config"].get("variant_tuning"))
    print("parity", d.get("parity") and {k: d["parity"][k] for k in ("bit_exact", "mismatch", "inputs_match_golden")})
    print("e2e", d.get("e2e") and (round(d["e2e"]["ms_per_step"], 2), d["e2e"].get("ms_per_step_by_handoff")))
    print("cpu", d.get("cpu_baseline"))
    for k in d["kernels"]:
        print("   %-18s %8.1f us x %4.1f  share %.3f  %s GB/s(alg)" % (k["kernel"], k["avg_ms"] * 1e3, k["launches_per_step"], k["share"], k["gbs"] and round(k["gbs"])))
except Exception as exc:
    print("no result:", exc)
PY

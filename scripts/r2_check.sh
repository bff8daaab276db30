#!/bin/bash
# First GPU call of a new round: everything that was added after round 1's GPU budget ran out
# (massck reduction, tendency diagnostics, pipelined hand-off, multi-rank spectral nudging, fused
# advection / first-sub-step exchanges).  All of it has run bit-exact on the CPU build of the CUDA
# sources (tests/test_emu_full.py); this is its first contact with a B200.
# usage: gpurun --timeout 420 -- bash scripts/r2_check.sh                      (1 GPU)
#        gpurun --gpus 2 --timeout 300 -- 'python -m pytest tests/test_gpu_multi.py -q'
#        gpurun --gpus 8 --timeout 600 -- bash scripts/scale8.sh                 (strong scaling, 16 -> 6 halo launches)
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider --durations=5 > gpurun_out/r2_gpu_tests.log 2>&1
tail -12 gpurun_out/r2_gpu_tests.log
# e2e: pipelined hand-off (default) against the sequential one, and the slab count
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_pipelined.json 2> gpurun_out/r2_bench_pipelined.err
BENCH_HANDOFF=sequential timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_sequential.json 2>/dev/null
BENCH_HANDOFF_SLABS=4 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_slabs4.json 2>/dev/null
BENCH_HANDOFF_SLABS=16 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_slabs16.json 2>/dev/null
for f in pipelined sequential slabs4 slabs16; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/r2_bench_{sys.argv[1]}.json") if l.startswith("{")][-1])
    e = d["e2e"]
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "e2e ms/step", round(e["ms_per_step"], 2), e.get("handoff"), e.get("handoff_note"))
except Exception as exc:
    print(sys.argv[1], "no result:", exc)
PY
done
# wsolve A/B (bench.py autotunes 5 vs 6 when MOLOCH_B200_WSOLVE is unset; here each is forced): 6 = 7 warps/SM,
# 5 = the measured round-1 kernel (45 % of the HBM peak, 4 warps/SM), 2 = CTA-parallel coefficients
for v in 6 7 5 2; do
  MOLOCH_B200_WSOLVE=$v timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_wsolve$v.json 2>/dev/null
  python - "$v" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/r2_bench_wsolve{sys.argv[1]}.json") if l.startswith("{")][-1])
    k = [x for x in d["kernels"] if x["kernel"] == "wsolve"][0]
    print("wsolve variant", sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "wsolve us", round(k["avg_ms"] * 1e3, 1), "GB/s", round(k["gbs"]))
except Exception as exc:
    print("wsolve variant", sys.argv[1], "no result:", exc)
PY
done
timeout 90 python scripts/kbench.py --boundary --slice --spectral --tke --massck --diag --steps 4 --warmup 1 \
  > gpurun_out/r2_kbench_all.json 2> gpurun_out/r2_kbench_all.err
tail -c 1500 gpurun_out/r2_kbench_all.json; tail -3 gpurun_out/r2_kbench_all.err

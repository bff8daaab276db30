#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r2n2}
N=${N:-2}
timeout 300 python -m pytest tests/test_gpu_multi.py -q -x --timeout 200 -p no:cacheprovider 2>&1 | tail -2
for g in 1 0; do
MOLOCH_B200_GRAPH=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$g \
  bench.py --gpus $N --steps 20 --warmup 3 --no-e2e ${BENCH_ARGS} > gpurun_out/${T}_graph$g.json 2> gpurun_out/${T}_graph$g.err
tail -c 200 gpurun_out/${T}_graph$g.err
python - "$g" "gpurun_out/${T}_graph$g.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print("graph", sys.argv[1], "N", d["n_gpus"], "ms/step %.3f" % d["ms_per_step"], "%.4e" % d["value"], "fusion", d["config"].get("halo_fusion_level"), "wsolve", d["config"]["wsolve_variant"], "parity", d.get("parity") and d["parity"]["bit_exact"])
    print("   ", {k["kernel"]: (round(k["avg_ms"] * 1e3, 1), k["launches_per_step"]) for k in d["kernels"]})
except Exception as exc:
    print("graph", sys.argv[1], "no result:", exc)
PY
done

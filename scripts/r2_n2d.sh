#!/bin/bash
# 2 GPUs: decomposed parity tests + cordex25 N=2 bench; then one GPU alone: kbench (per-kernel times).
T=${1:-r2n2d}
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests/test_gpu_multi.py -q -rs --timeout 300 -p no:cacheprovider ) > gpurun_out/${T}_pytest_multi.log 2>&1
tail -3 gpurun_out/${T}_pytest_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/${T}_n2.json 2> gpurun_out/${T}_n2.err
python - gpurun_out/${T}_n2.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("N=2 %.3f ms/step" % d["ms_per_step"], "wsolve", d["config"]["wsolve_variant"], "fusion", d["config"].get("halo_fusion_level"), d["config"].get("halo_signal"), d["config"].get("halo_wz_fused"), "parity", d["parity"]["bit_exact"])
print("   ", {k["kernel"]: (round(k["avg_ms"] * 1e3, 1), k["launches_per_step"]) for k in d["kernels"]})
PY
CUDA_VISIBLE_DEVICES=0 timeout 200 python scripts/kbench.py --steps 6 --warmup 2 > gpurun_out/${T}_kbench.json 2>/dev/null
python - gpurun_out/${T}_kbench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N=1 kbench %.3f ms/step" % d["ms_per_step"], {k: round(v["avg_ms"] * 1e3, 1) for k, v in d["kernels"].items()})
PY

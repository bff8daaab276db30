#!/bin/bash
# 2 GPUs: decomposed parity tests (incl. the rows-only cases with producer-side signalling and the fused wz round),
# then cordex25 at N=2 with the signalling side / wz fusion switched.   usage: gpurun --gpus 2 -- bash scripts/r2_n2b.sh TAG
T=${1:-r2n2b}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 500 python -m pytest tests/test_gpu_multi.py -q -rs --timeout 300 -p no:cacheprovider ) > gpurun_out/${T}_pytest_multi.log 2>&1
tail -4 gpurun_out/${T}_pytest_multi.log
run() {  # name extra-env...
  local name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/${T}_${name}.json 2> gpurun_out/${T}_${name}.err
  python - gpurun_out/${T}_${name}.json $name <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[2], "%.3f ms/step" % d["ms_per_step"], "wsolve", d["config"]["wsolve_variant"], d["config"]["variant_tuning"]["wsolve"].get("ms_per_step"),
          "fusion", d["config"].get("halo_fusion_level"), d["config"].get("halo_signal"), d["config"].get("halo_wz_fused"), "parity", d["parity"]["bit_exact"], "launches", d["gpu_launches"])
    print("   ", {k["kernel"]: (round(k["avg_ms"] * 1e3, 1), k["launches_per_step"]) for k in d["kernels"]})
except Exception as exc:
    print(sys.argv[2], "FAILED", exc)
PY
}
run psig1_wz1 X=1
run psig0_wz1 MOLOCH_B200_PSIGNAL=0
run psig1_wz0 MOLOCH_B200_FUSE_WZ=0
tail -3 gpurun_out/${T}_psig1_wz1.err

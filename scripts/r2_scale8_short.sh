#!/bin/bash
# 8-GPU box, short: decomposed parity tests, cordex25 at N=8 (defaults and with the status rounds unfused), N=4 and cp3km N=8.
T=${1:-r2s8e}
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_multi.py -q -rs --timeout 300 -p no:cacheprovider ) > gpurun_out/${T}_pytest_multi.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest_multi.log
run() {  # name ngpu devices port workload steps extra-env
  local name=$1 n=$2 dev=$3 port=$4 wl=$5 steps=$6; shift 6
  env CUDA_VISIBLE_DEVICES=$dev "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
    --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps $steps --warmup 3 --no-e2e --workload $wl \
    > gpurun_out/${T}_${name}.json 2> gpurun_out/${T}_${name}.err
}
show() {
  python - $T "$@" <<'PY'
import json, sys
T = sys.argv[1]
for name in sys.argv[2:]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/{T}_{name}.json") if l.startswith("{")][-1])
        c = d["config"]
        print(name, c["workload"], c["decomposition"], "fusion", c.get("halo_fusion_level"), "%.3e c-u/s" % d["value"], "%.3f ms/step" % d["ms_per_step"],
              "launches", d["gpu_launches"], "wsolve", c["wsolve_variant"], {k: round(v, 3) for k, v in c["variant_tuning"]["wsolve"].get("ms_per_step", {}).items()},
              c["variant_tuning"].get("fuse_halo"), "parity", d.get("parity") and d["parity"]["bit_exact"])
        print("   ", {k["kernel"]: (round(k["avg_ms"] * 1e3, 1), k["launches_per_step"]) for k in d["kernels"]})
    except Exception as e:
        print(name, "FAILED", e)
PY
}
ALL=0,1,2,3,4,5,6,7
run n8 8 $ALL 29518 cordex25 20 X=1;                              show n8
run n8_nostatus 8 $ALL 29519 cordex25 20 MOLOCH_B200_FUSE_STATUS=0; show n8_nostatus
run cp3km_n8 8 $ALL 29521 cp3km 6 X=1 ;                           show cp3km_n8
run n4 4 0,1,2,3 29523 cordex25 20 X=1 &
run n2 2 4,5 29524 cordex25 20 X=1 &
wait
show n4 n2

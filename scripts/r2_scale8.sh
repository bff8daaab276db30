#!/bin/bash
# 8-GPU box: decomposed parity tests (log kept), cordex25 N=8, cp3km N=8 [, tracer40 N=8], optionally N=4/2/1 side by side.
mkdir -p gpurun_out
T=${TAG:-r2s8}
nvidia-smi -L | wc -l
( time timeout 400 python -m pytest tests/test_gpu_multi.py -q -rs --timeout 300 -p no:cacheprovider ) > gpurun_out/${T}_pytest_multi.log 2>&1
tail -5 gpurun_out/${T}_pytest_multi.log
run() {  # name ngpu devices port workload steps extra-env...
  local name=$1 n=$2 dev=$3 port=$4 wl=$5 steps=$6; shift 6
  if [ $n -eq 1 ]; then
    env CUDA_VISIBLE_DEVICES=$dev "$@" timeout 600 python bench.py --steps $steps --warmup 3 --no-e2e --no-cpu-baseline --workload $wl \
      > gpurun_out/${T}_${name}.json 2> gpurun_out/${T}_${name}.err
  else
    env CUDA_VISIBLE_DEVICES=$dev "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
      --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps $steps --warmup 3 --no-e2e --workload $wl \
      > gpurun_out/${T}_${name}.json 2> gpurun_out/${T}_${name}.err
  fi
}
show() {
  python - "$@" <<'PY'
import json,sys,os
T=os.environ.get("TAG","r2s8")
for name in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/{T}_{name}.json") if l.startswith("{")][-1])
        print(name, d["config"]["workload"], d["config"]["decomposition"], d["config"].get("halo_fusion_level"), "%.3e c-u/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], "launches", d["gpu_launches"], "wsolve", d["config"]["wsolve_variant"], "parity", d.get("parity") and d["parity"]["bit_exact"])
        print("   ", {k["kernel"]: (round(k["avg_ms"]*1e3,1), k["launches_per_step"]) for k in d["kernels"]})
    except Exception as e:
        print(name, "FAILED", e)
PY
}
run n8 8 0,1,2,3,4,5,6,7 29518 cordex25 20
show n8
run cp3km_n8 8 0,1,2,3,4,5,6,7 29519 cp3km 6
show cp3km_n8
if [ -n "$TRACER40" ]; then run tracer40_n8 8 0,1,2,3,4,5,6,7 29520 tracer40 4; show tracer40_n8; fi
if [ -n "$SIDE" ]; then
  run n4 4 0,1,2,3 29521 cordex25 20 &
  run n2 2 4,5 29522 cordex25 20 &
  run n1 1 6 0 cordex25 20 &
  wait
  show n4 n2 n1
fi
tail -2 gpurun_out/${T}_n8.err

#!/bin/bash
# 8-GPU box: decomposed parity tests (2x4), then the strong-scaling sweep.
# N=8 runs alone; N=4, N=2 and N=1 run side by side on disjoint GPUs.
mkdir -p gpurun_out
T=${TAG:-s8}
nvidia-smi -L | wc -l
( time timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -k "${TESTS:-2x2 or 2x4}" ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
run() {  # name ngpu devices port extra-env...
  local name=$1 n=$2 dev=$3 port=$4; shift 4
  if [ $n -eq 1 ]; then
    env CUDA_VISIBLE_DEVICES=$dev "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline \
      > gpurun_out/${T}_${name}.json 2> gpurun_out/${T}_${name}.err
  else
    env CUDA_VISIBLE_DEVICES=$dev "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
      --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 ${NOE2E---no-e2e} \
      > gpurun_out/${T}_${name}.json 2> gpurun_out/${T}_${name}.err
  fi
}
show() {
  python - "$@" <<'PY'
import json,sys
for name in sys.argv[1:]:
    try:
        d=json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
        print(name, d["config"]["decomposition"], d["config"]["halo_transport"], "%.3e c-u/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], "launches", d["gpu_launches"], "finite", d["finite"])
        print("   ", {k["kernel"]: (round(k["avg_ms"]*1e3,1), k["launches_per_step"]) for k in d["kernels"]})
    except Exception as e:
        print(name, "FAILED", e)
PY
}
NOE2E= run n8 8 0,1,2,3,4,5,6,7 29518   # the driver's own N=8 invocation (with the e2e leg)
run n8_2x4 8 0,1,2,3,4,5,6,7 29519 BENCH_PX=2 BENCH_PY=4
show ${T}_n8 ${T}_n8_2x4
# A/B of what changed without GPU access at the end of round 1: halo fusion level (2 = all, 1 = sound loop's
# sub-steps 2.. only: the configuration of profiles/r1_scale_*.json) and the wsolve variant (6 default, 5 measured)
run n8_fuse1 8 0,1,2,3,4,5,6,7 29523 MOLOCH_B200_FUSE_HALO=1
run n8_wsolve5 8 0,1,2,3,4,5,6,7 29524 MOLOCH_B200_WSOLVE=5
show ${T}_n8_fuse1 ${T}_n8_wsolve5
if [ -n "$BIG" ]; then   # BASELINE configs 4 and 5 (8 GPUs, 2x4)
  for w in cp3km tracer40; do
    env BENCH_PX=2 BENCH_PY=4 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
      --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --workload $w \
      > gpurun_out/${T}_$w.json 2> gpurun_out/${T}_$w.err
    show ${T}_$w
  done
fi
run n4 4 0,1,2,3 29521 &
run n2 2 4,5 29522 &
run n1 1 6 0 &
wait
show ${T}_n4 ${T}_n2 ${T}_n1
tail -3 gpurun_out/${T}_n8.err

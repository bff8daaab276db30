#!/bin/bash
# 2 GPUs: decomposed parity tests (log kept under profiles/), then bench --gpus 2 with the parity block.
mkdir -p gpurun_out
T=${TAG:-r2m2}
nvidia-smi -L | wc -l
( time timeout 600 python -m pytest tests/test_gpu_multi.py -q -rs --timeout 300 -p no:cacheprovider ) > gpurun_out/${T}_pytest_multi.log 2>&1
tail -8 gpurun_out/${T}_pytest_multi.log
for lv in 2 1; do
MOLOCH_B200_FUSE_HALO=$lv MOLOCH_B200_FUSE_HALO_FIXED=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$lv \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/${T}_bench_n2_fuse$lv.json 2> gpurun_out/${T}_bench_n2_fuse$lv.err
tail -c 300 gpurun_out/${T}_bench_n2_fuse$lv.err
done
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/%s_bench_n2*.json" % os.environ.get("TAG", "r2m2"))):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(os.path.basename(f), "ms/step %.3f" % d["ms_per_step"], d["config"].get("halo_fusion_level"), "parity", d.get("parity") and (d["parity"]["bit_exact"], d["parity"]["n_ranks"], d["parity"]["inputs_match_golden"]))
        print("   ", {k["kernel"]: (round(k["avg_ms"] * 1e3, 1), k["launches_per_step"]) for k in d["kernels"]})
    except Exception as exc:
        print(f, "no result:", exc)
PY

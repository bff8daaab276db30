#!/bin/bash
# Round-2 evidence on one GPU (final tree): bench line, ncu launch list of the same command, ncu --set full of one
# launch of every hot kernel (torch-free kbench; digests made on the box, reports kept small: gpurun_out <= 64 MiB),
# cp3km / isc24_small / fast-mode bench lines, the per-rank grid of the 8-GPU run on one GPU.
# usage: gpurun --timeout 1500 -- bash scripts/r2_profile.sh TAG
TAG=${1:-r2f}
mkdir -p gpurun_out
nvidia-smi -L
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.err
MOLOCH_B200_WSOLVE=12 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity \
  > gpurun_out/${TAG}_launches.log 2>&1
# last sound sub-step of the first sound call + the advection kernels of the second step: 8 launches
MOLOCH_B200_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on \
  --kernel-name 'regex:moloch_(waf_horizontal|waf_vertical2|wsolve_tm|sound_div|uvupdate2|destagger|restagger|status_update|curvature|tetavf_init)' \
  --launch-skip ${SKIP:-57} --launch-count ${COUNT:-8} -f -o gpurun_out/${TAG}_full \
  python scripts/kbench.py --steps 1 --warmup 1 > gpurun_out/${TAG}_full.log 2>&1
python scripts/ncu_digest.py gpurun_out/${TAG}_full.ncu-rep > gpurun_out/${TAG}_full_digest.txt 2>&1
MOLOCH_B200_GRAPH=0 timeout 300 ncu --set full --clock-control none \
  --kernel-name 'regex:moloch_(status_update|tvirt_temp|diag_prq)' --launch-skip 3 --launch-count 3 -f -o gpurun_out/${TAG}_full2 \
  python scripts/kbench.py --steps 1 --warmup 1 > gpurun_out/${TAG}_full2.log 2>&1
python scripts/ncu_digest.py gpurun_out/${TAG}_full2.ncu-rep > gpurun_out/${TAG}_full2_digest.txt 2>&1
timeout 400 python bench.py --workload cp3km --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_cp3km_n1.json 2> gpurun_out/${TAG}_cp3km_n1.err
timeout 300 python bench.py --workload isc24_small --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/${TAG}_isc24.json 2> gpurun_out/${TAG}_isc24.err
timeout 400 python bench.py --mode fast --no-cpu-baseline > gpurun_out/${TAG}_bench_fast.json 2> gpurun_out/${TAG}_bench_fast.err
# the per-rank grid of cordex25 on 8 GPUs (400 x 50 columns) on ONE GPU: what the small grid costs without any halo wait
timeout 200 python scripts/kbench.py --jx 400 --iy 50 --steps 20 --warmup 3 > gpurun_out/${TAG}_k400x50.json 2> gpurun_out/${TAG}_k400x50.err
MOLOCH_B200_GRAPH=0 timeout 200 python scripts/kbench.py --jx 400 --iy 50 --steps 20 --warmup 3 > gpurun_out/${TAG}_k400x50_nograph.json 2> gpurun_out/${TAG}_k400x50_nograph.err
timeout 200 python scripts/kbench.py --jx 400 --iy 100 --steps 20 --warmup 3 > gpurun_out/${TAG}_k400x100.json 2> gpurun_out/${TAG}_k400x100.err
du -sh gpurun_out; ls -la gpurun_out | tail -25

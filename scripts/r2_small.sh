#!/bin/bash
# one GPU, the per-rank grid of cordex25 on 8 GPUs (400 x 50 columns): wsolve variants and the other kernels
# without any halo traffic -- the small-grid floor of every kernel
for v in 5 8 9 2; do
  MOLOCH_B200_WSOLVE=$v timeout 100 python scripts/kbench.py --jx 400 --iy 52 --steps 10 --warmup 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['kernels']
print('wsolve $v: %.3f ms/step ' % d['ms_per_step'] + ' '.join('%s=%.1f' % (k, ks[k]['avg_ms']*1e3) for k in ('wsolve','sound_pre','uvupdate','waf_horizontal','waf_vertical','status_update','destagger','restagger','curvature','reset_tendencies') if k in ks))"
done

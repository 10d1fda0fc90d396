#!/bin/bash
# GPU run (8 GPUs): A/B of the per-slab block size on ONE box: bench at N = 8 with and without the 224-thread block, twice each.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for rep in 1 2; do
for alt in 1 0; do
  QNMFIT_K1_ALT_BLOCK=$alt timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$rep \
      bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_ab8_alt${alt}_$rep.log 2>&1
  python - <<PY
import json
for l in open('gpurun_out/r2_ab8_alt${alt}_$rep.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('alt', $alt, 'rep', $rep, 'block', d['config'].get('kernel_plan', d['config']).get('block') if isinstance(d['config'].get('kernel_plan', None), dict) else '?', 'ms %.4f'%d['ms_per_step'], 'e2e %.4f ms'%d['e2e']['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
PY
done
done

#!/bin/bash
# GPU run (8 GPUs): the bench at N = 8, 4 and 2 (strong scaling of the 256 x 256 grid).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_bench_${n}gpu.log 2>&1; echo "bench exit $?" >> gpurun_out/r2_bench_${n}gpu.log
  grep -E '^\{|exit' gpurun_out/r2_bench_${n}gpu.log | cut -c1-400
done

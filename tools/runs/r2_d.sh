#!/bin/bash
# GPU run D (2 GPUs): peer-exchange tests, then the bench at N = 2.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.log 2>&1
timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > gpurun_out/r2_tests_peer.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_peer.log
tail -25 gpurun_out/r2_tests_peer.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu.log 2>&1; echo "bench exit $?" >> gpurun_out/r2_bench_2gpu.log
grep -E '^\{|exit' gpurun_out/r2_bench_2gpu.log | cut -c1-2500

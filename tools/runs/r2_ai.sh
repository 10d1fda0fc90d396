#!/bin/bash
# GPU run AI (2 GPUs): peer tests and the 2-GPU bench line on the final build of the round.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > gpurun_out/r2_tests_peer_ai.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_peer_ai.log
tail -3 gpurun_out/r2_tests_peer_ai.log
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_2gpu_ai.log 2>&1; echo "bench exit $?" >> gpurun_out/r2_bench_2gpu_ai.log
tail -2 gpurun_out/r2_bench_2gpu_ai.log | cut -c1-330
grep -o '"parity": {[^}]*}' gpurun_out/r2_bench_2gpu_ai.log | cut -c1-200
grep -o '"e2e": {[^}]*}' gpurun_out/r2_bench_2gpu_ai.log | cut -c1-200

#!/bin/bash
# GPU run R: full GPU tier after the rank-prefilter change; cfg4 kernel times (K3 / K4 now always run the estimator at 40 columns).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_r.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_r.log
tail -5 gpurun_out/r2_tests_r.log
timeout 300 python tools/k3_time.py 10 > gpurun_out/r2_k4_time.log 2>&1; tail -3 gpurun_out/r2_k4_time.log | cut -c1-120
timeout 600 python tools/fuzz_parity.py 42 200 > gpurun_out/r2_fuzz42.log 2>&1; tail -3 gpurun_out/r2_fuzz42.log | cut -c1-300
timeout 600 python tools/midn_time.py 8 10 12 14 16 20 24 > gpurun_out/r2_midn_r.log 2>&1; grep -E "kernel (5|1) " gpurun_out/r2_midn_r.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_r.log 2>&1; tail -1 gpurun_out/r2_bench_r.log | cut -c1-330

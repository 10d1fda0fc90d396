#!/bin/bash
# GPU run I: full GPU test tier, slab times and the bench on the build with the per-slab block size.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_i.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_i.log
tail -3 gpurun_out/r2_tests_i.log
timeout 300 python tools/slab_time.py 20 > gpurun_out/r2_slab.log 2>&1; cut -c1-200 gpurun_out/r2_slab.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_i.log 2>&1; tail -1 gpurun_out/r2_bench_i.log | cut -c1-600

#!/bin/bash
# GPU run V: K1 with per-fit table rows: K1 tests, bench (kernel time of the headline grid).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not peer" > gpurun_out/r2_tests_v.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_v.log
tail -3 gpurun_out/r2_tests_v.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_bench_v.log 2>&1; tail -1 gpurun_out/r2_bench_v.log | cut -c1-330
MIDN_ONLY_AUTO=1 timeout 300 python tools/midn_time.py 8 > gpurun_out/r2_midn8.log 2>&1; tail -1 gpurun_out/r2_midn8.log
timeout 300 python tools/smalln_time.py > gpurun_out/r2_smalln.log 2>&1; tail -10 gpurun_out/r2_smalln.log | cut -c1-200

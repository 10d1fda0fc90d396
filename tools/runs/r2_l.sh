#!/bin/bash
# GPU run L: full GPU tier with K1p in AUTO, mid-N timing, fuzz (N up to 16).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_l.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_l.log
tail -12 gpurun_out/r2_tests_l.log
timeout 600 python tools/midn_time.py 8 9 10 11 12 13 14 15 16 24 > gpurun_out/r2_midn.log 2>&1; cat gpurun_out/r2_midn.log | tail -40
timeout 900 python tools/fuzz_parity.py 40 300 > gpurun_out/r2_fuzz40.log 2>&1; tail -4 gpurun_out/r2_fuzz40.log | cut -c1-600

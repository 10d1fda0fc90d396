#!/bin/bash
# GPU run S: K1p with the group-cooperative rank estimator: tests, mid-N timing, fuzz.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "struct_kernel_vs_oracle or pair_kernel or free_frequency or deficient or rank or flagged" > gpurun_out/r2_tests_s.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_s.log
tail -5 gpurun_out/r2_tests_s.log
timeout 600 python tools/midn_time.py 9 10 11 12 13 14 15 16 18 20 24 > gpurun_out/r2_midn.log 2>&1; grep -E "kernel 5 " gpurun_out/r2_midn.log


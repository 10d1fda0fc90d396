#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "struct_kernel_vs_oracle or pair_kernel" > gpurun_out/r2_tests_y.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_y.log
tail -3 gpurun_out/r2_tests_y.log
MIDN_ONLY_AUTO=1 timeout 600 python tools/midn_time.py 11 13 14 15 16 17 > gpurun_out/r2_midn_y.log 2>&1; cat gpurun_out/r2_midn_y.log

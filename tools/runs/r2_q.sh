#!/bin/bash
# GPU run Q: K1p up to 24 columns: its tests and the mid-N timing for 16 .. 24.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "struct_kernel_vs_oracle or pair_kernel" > gpurun_out/r2_tests_q.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_q.log
tail -5 gpurun_out/r2_tests_q.log
timeout 600 python tools/midn_time.py 16 17 18 20 22 24 > gpurun_out/r2_midn_hi.log 2>&1; cat gpurun_out/r2_midn_hi.log

#!/bin/bash
# GPU run J: K1p (columns split over lanes) — parity tests of the kernels and the mid-N timing.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "struct_kernel_vs_oracle" > gpurun_out/r2_tests_j.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_j.log
tail -15 gpurun_out/r2_tests_j.log
timeout 600 python tools/midn_time.py 8 9 10 11 12 13 14 15 16 > gpurun_out/r2_midn.log 2>&1; cat gpurun_out/r2_midn.log | tail -40

#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "multimode or struct or config4 or quadratic" > gpurun_out/r2_tests_b.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_b.log
tail -4 gpurun_out/r2_tests_b.log
timeout 300 python tools/k3_time.py 10 > gpurun_out/r2_k4_time.log 2>&1; tail -4 gpurun_out/r2_k4_time.log | cut -c1-200

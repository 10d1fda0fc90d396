#!/bin/bash
# GPU run C: the whole 1-GPU parity suite.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_full.log
tail -40 gpurun_out/r2_tests_full.log

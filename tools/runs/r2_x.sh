#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "struct or multimode or config4 or quadratic or dynamic" > gpurun_out/r2_tests_x.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_x.log
tail -3 gpurun_out/r2_tests_x.log
timeout 600 python tools/other_configs.py > gpurun_out/r2_other.log 2>&1; grep -E "gpu_ms" gpurun_out/r2_other.log

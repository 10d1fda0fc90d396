#!/bin/bash
# GPU run N (2+ GPUs): peer / device-group tests and the single-process device-group timing.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -x -q -k device_group > gpurun_out/r2_tests_n.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_n.log
tail -12 gpurun_out/r2_tests_n.log
timeout 300 python tools/devgroup_time.py > gpurun_out/r2_devgroup.log 2>&1; tail -6 gpurun_out/r2_devgroup.log
timeout 300 python tools/devgroup_profile.py > gpurun_out/r2_devgroup_profile.log 2>&1; tail -12 gpurun_out/r2_devgroup_profile.log

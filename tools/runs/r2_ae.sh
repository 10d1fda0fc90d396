#!/bin/bash
# GPU run AE: GPU tier, cfg5 time and host profile after the lock-step Nelder-Mead moved into libqnmfit.so.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_ae.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_ae.log
tail -3 gpurun_out/r2_tests_ae.log
timeout 200 python tools/cfg5_time.py > gpurun_out/r2_cfg5_ae.log 2>&1; tail -16 gpurun_out/r2_cfg5_ae.log
timeout 200 python tools/cfg5_profile.py > gpurun_out/r2_cfg5_profile_ae.log 2>&1; head -40 gpurun_out/r2_cfg5_profile_ae.log | cut -c1-150
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/other_configs.py > gpurun_out/r2_other_ae.log 2>&1; head -8 gpurun_out/r2_other_ae.log

#!/bin/bash
# GPU run T: full GPU tier, smoke, fuzz (two seeds), cfg4 / cfg5 / other configs timing on the final K1p build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_t.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_t.log
tail -4 gpurun_out/r2_tests_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python tools/fuzz_parity.py 44 300 > gpurun_out/r2_fuzz44.log 2>&1; tail -4 gpurun_out/r2_fuzz44.log | cut -c1-200
timeout 300 python tools/k3_time.py 10 > gpurun_out/r2_k4_time.log 2>&1; tail -3 gpurun_out/r2_k4_time.log | cut -c1-100
timeout 600 python tools/cfg5_time.py > gpurun_out/r2_cfg5.log 2>&1; tail -3 gpurun_out/r2_cfg5.log | cut -c1-300
timeout 600 python tools/other_configs.py > gpurun_out/r2_other.log 2>&1; tail -6 gpurun_out/r2_other.log | cut -c1-300

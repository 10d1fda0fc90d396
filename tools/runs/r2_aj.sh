#!/bin/bash
# GPU run AJ: the tests of the explicit-frequency / optimiser-driven paths and smoke() on the final tree.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "free_frequency or objective or omega_grid or epsilon or config5 or ringdown_fit_cases or dynamic" > gpurun_out/r2_tests_aj.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_aj.log
tail -3 gpurun_out/r2_tests_aj.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

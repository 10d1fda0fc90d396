#!/bin/bash
# GPU run P: full GPU tier, mid-N timing (all kernels), smoke, on the tuned K1p build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_p.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_p.log
tail -4 gpurun_out/r2_tests_p.log
timeout 600 python tools/midn_time.py 8 9 10 11 12 13 14 15 16 24 > gpurun_out/r2_midn.log 2>&1; cat gpurun_out/r2_midn.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

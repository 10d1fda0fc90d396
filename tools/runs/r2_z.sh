#!/bin/bash
# GPU run Z: full GPU tier + smoke + short bench on the last build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_z.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_z.log
tail -3 gpurun_out/r2_tests_z.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_z.log 2>&1; tail -1 gpurun_out/r2_bench_z.log | cut -c1-200
timeout 600 python tools/fuzz_parity.py 45 200 > gpurun_out/r2_fuzz45.log 2>&1; tail -3 gpurun_out/r2_fuzz45.log | cut -c1-160

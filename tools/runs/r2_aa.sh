#!/bin/bash
# GPU run AA: full GPU tier, cfg4 kernel times and the bench after the K3 load-order change.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_aa.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_aa.log
tail -3 gpurun_out/r2_tests_aa.log
timeout 300 python tools/k3_time.py 10 > gpurun_out/r2_k4_time.log 2>&1; tail -3 gpurun_out/r2_k4_time.log | cut -c1-100
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_aa.log 2>&1; tail -1 gpurun_out/r2_bench_aa.log | cut -c1-200

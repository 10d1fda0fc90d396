#!/bin/bash
# GPU run: what the driver runs at round end — the GPU test tier, smoke(), both bench arms.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_verify_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_verify_tests.log
tail -4 gpurun_out/r2_verify_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_verify_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2_verify_smoke.log
tail -2 gpurun_out/r2_verify_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_verify_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/r2_verify_ref.log
tail -2 gpurun_out/r2_verify_ref.log | cut -c1-700
timeout 600 python bench.py > gpurun_out/r2_verify_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/r2_verify_bench.log
tail -2 gpurun_out/r2_verify_bench.log | cut -c1-400

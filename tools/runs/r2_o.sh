#!/bin/bash
# GPU run O: K1p after the table re-layout / full-block path: its tests, mid-N timing, fuzz.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "struct_kernel_vs_oracle or pair_kernel or free_frequency or slab" > gpurun_out/r2_tests_o.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_o.log
tail -5 gpurun_out/r2_tests_o.log
timeout 600 python tools/midn_time.py 9 10 11 12 13 14 15 16 > gpurun_out/r2_midn.log 2>&1; grep -E "kernel 5" gpurun_out/r2_midn.log
timeout 900 python tools/fuzz_parity.py 41 200 > gpurun_out/r2_fuzz41.log 2>&1; tail -3 gpurun_out/r2_fuzz41.log | cut -c1-300

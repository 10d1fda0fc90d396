#!/bin/bash
# GPU run B: K4 parity (struct tests), timing, and one ncu --set full capture of fit_panel_kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
ls -la oracle/_ref/qnmfits > gpurun_out/r2_ref_ls.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "multimode or struct or config4 or quadratic" > gpurun_out/r2_tests_b.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_b.log
tail -25 gpurun_out/r2_tests_b.log
timeout 300 python tools/k3_time.py 10 > gpurun_out/r2_k4_time.log 2>&1; tail -5 gpurun_out/r2_k4_time.log
K3_FITS=296 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fit_panel -c 1 \
   -o gpurun_out/prof_k4_r02_v4 -f python tools/k3_time.py 1 > gpurun_out/r2_ncu_k4.log 2>&1
tail -3 gpurun_out/r2_ncu_k4.log

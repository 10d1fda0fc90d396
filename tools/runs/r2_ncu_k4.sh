#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
K3_FITS=148 QNMFIT_K4_ONE_PER_SM=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fit_panel -c 1 \
   -o gpurun_out/prof_k4_r02_solo -f python tools/k3_time.py 1 > gpurun_out/r2_ncu_k4.log 2>&1
tail -2 gpurun_out/r2_ncu_k4.log

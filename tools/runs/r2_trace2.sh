#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
QNMFIT_K4_ONE_PER_SM=1 QNMFIT_LIB=tools/_variants/libqnmfit_k4trace.so timeout 300 python tools/k4_trace.py 148 > gpurun_out/r2_k4_trace_1persm.log 2>&1
QNMFIT_LIB=tools/_variants/libqnmfit_k4trace.so timeout 300 python tools/k4_trace.py 296 > gpurun_out/r2_k4_trace.log 2>&1
for f in 148 296; do
K3_FITS=$f QNMFIT_K4_ONE_PER_SM=1 timeout 300 python tools/k3_time.py 5 2>&1 | grep "^k4" | cut -c1-80
K3_FITS=$f timeout 300 python tools/k3_time.py 5 2>&1 | grep "^k4" | cut -c1-80
done

#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
QNMFIT_LIB=tools/_variants/libqnmfit_k4trace.so timeout 300 python tools/k4_trace.py 296 > gpurun_out/r2_k4_trace.log 2>&1
head -5 gpurun_out/r2_k4_trace.log

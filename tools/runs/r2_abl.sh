#!/bin/bash
# Stage durations of K4's pipeline from clock-stamp builds (tools/build_variant.sh <name> -DK4_TRACE ...),
# one fit per SM and two.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for lib in tools/_variants/libqnmfit_*.so; do
QNMFIT_K4_ONE_PER_SM=1 QNMFIT_LIB=$lib timeout 300 python tools/k4_trace_summary.py 148 2>&1 | grep -vE "^\s*$"
QNMFIT_LIB=$lib timeout 300 python tools/k4_trace_summary.py 296 2>&1 | grep -vE "^\s*$"
done > gpurun_out/r2_k4_abl.log 2>&1
cat gpurun_out/r2_k4_abl.log

#!/bin/bash
# GPU run G: randomised differential test on the current build (seed 31 with K4 preferred over K3).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/fuzz_parity.py 30 250 > gpurun_out/r2_fuzz30.log 2>&1; tail -3 gpurun_out/r2_fuzz30.log
QNMFIT_AUTO_PANEL=1 timeout 600 python tools/fuzz_parity.py 31 250 > gpurun_out/r2_fuzz31.log 2>&1; tail -3 gpurun_out/r2_fuzz31.log

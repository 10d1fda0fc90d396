#!/bin/bash
# GPU run M: the same fuzz seed with K1p disabled in AUTO (K1 / K3 take N = 10 .. 16), for comparison.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cp gpurun_out/fuzz_40.json gpurun_out/fuzz_40_pair.json 2>/dev/null
QNMFIT_AUTO_PAIR=0 timeout 900 python tools/fuzz_parity.py 40 300 > gpurun_out/r2_fuzz40_nopair.log 2>&1; tail -4 gpurun_out/r2_fuzz40_nopair.log | cut -c1-300

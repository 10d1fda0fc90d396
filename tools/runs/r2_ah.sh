#!/bin/bash
# GPU run AH: fuzz seed 46, 1000 cases, final round-2 build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 200 python tools/fuzz_parity.py 46 1000 ) > gpurun_out/r2_fuzz46_ah.log 2>&1; tail -14 gpurun_out/r2_fuzz46_ah.log | cut -c1-400

#!/bin/bash
# GPU run AG: fuzz seed 45 again on the build with the one-call host paths (single fits, explicit-frequency launches).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 170 python tools/fuzz_parity.py 45 200 ) > gpurun_out/r2_fuzz45_ag.log 2>&1; tail -9 gpurun_out/r2_fuzz45_ag.log | cut -c1-120

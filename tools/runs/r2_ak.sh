#!/bin/bash
# GPU run AK: K1 with the two warps of a scheduler started out of phase (QNMFIT_K1_SKEW_NS), cfg3 grid kernel time.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 50 python tools/k1_variants.py run default skew300 skew1200 default > gpurun_out/r2_k1_skew_ak.log 2>&1; tail -5 gpurun_out/r2_k1_skew_ak.log | cut -c1-150

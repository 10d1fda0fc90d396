#!/bin/bash
# GPU run E: K4 timing only (quick A/B of kernel variants).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/k3_time.py 10 > gpurun_out/r2_k4_time.log 2>&1; tail -4 gpurun_out/r2_k4_time.log

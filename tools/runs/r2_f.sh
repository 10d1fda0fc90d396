#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/k34_sweep.py > gpurun_out/r2_k34_sweep.log 2>&1; tail -16 gpurun_out/r2_k34_sweep.log

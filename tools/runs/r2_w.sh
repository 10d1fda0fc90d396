#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/cfg4_check.py > gpurun_out/r2_cfg4_check.log 2>&1; cat gpurun_out/r2_cfg4_check.log | tail -20

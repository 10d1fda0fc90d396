#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/other_configs.py > gpurun_out/r2_other_configs.log 2>&1; tail -32 gpurun_out/r2_other_configs.log

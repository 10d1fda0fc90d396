#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/other_configs.py > gpurun_out/r2_other.log 2>&1; grep -E "gpu_ms|fits_per_s" gpurun_out/r2_other.log

#!/bin/bash
# GPU run AF: host profile of a single ringdown_fit and of a repeated t0 sweep.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python tools/cfg1_profile.py > gpurun_out/r2_cfg1_profile_af.log 2>&1; grep -v "^$" gpurun_out/r2_cfg1_profile_af.log | cut -c1-140 | head -70

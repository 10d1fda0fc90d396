#!/bin/bash
# GPU run U: the mid-N table (every kernel that takes each shape) and ncu --set full of K1p at N = 12, 16, 24 (after the plain run).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python tools/midn_time.py 8 9 10 11 12 13 14 15 16 17 18 19 20 21 22 23 24 28 > gpurun_out/r2_midn.log 2>&1; cat gpurun_out/r2_midn.log | cut -c1-130
for n in 12 16 24; do
MIDN_ONLY_AUTO=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fit_pair -c 1 \
   -o gpurun_out/prof_k1p_n${n}_r02 -f python tools/midn_time.py $n > gpurun_out/r2_ncu_k1p_$n.log 2>&1
tail -1 gpurun_out/r2_ncu_k1p_$n.log
done

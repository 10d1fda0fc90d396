#!/bin/bash
# GPU run: the new GPU tests first (plain), then ncu --set full of K1p (N = 12 and N = 16, 128 x 128 grid).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "free_frequency_objective or pair_kernel" > gpurun_out/r2_tests_p.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_p.log
tail -4 gpurun_out/r2_tests_p.log
MIDN_ONLY_AUTO=1 timeout 300 python tools/midn_time.py 12 16 > gpurun_out/r2_midn_plain.log 2>&1; tail -2 gpurun_out/r2_midn_plain.log
for n in 12 16; do
MIDN_ONLY_AUTO=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fit_pair -c 1 \
   -o gpurun_out/prof_k1p_n${n}_r02 -f python tools/midn_time.py $n > gpurun_out/r2_ncu_k1p_$n.log 2>&1
tail -1 gpurun_out/r2_ncu_k1p_$n.log
done

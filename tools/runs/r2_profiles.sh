#!/bin/bash
# GPU run: the evidence for profiles/ (round 2).  Every ncu pass is preceded by the same command run plain.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
set -x
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.log 2>&1; echo "bench exit $?" >> gpurun_out/r2_bench_final.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r2_bench_short.log 2>&1 \
 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r2_ncu_list.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fit_small -c 1 -o gpurun_out/prof_k1_r02 -f \
      python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r2_ncu_k1.log 2>&1
timeout 300 python tools/k3_time.py 10 > gpurun_out/r2_k4_time.log 2>&1
K3_FITS=296 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fit_panel -c 1 \
   -o gpurun_out/prof_k4_r02_final -f python tools/k3_time.py 1 > gpurun_out/r2_ncu_k4.log 2>&1
timeout 300 python tools/e2e_profile.py > gpurun_out/r2_e2e_profile.log 2>&1
timeout 300 python tools/midn_time.py 8 9 10 12 13 16 24 > gpurun_out/r2_midn.log 2>&1
timeout 300 python tools/cfg5_time.py > gpurun_out/r2_cfg5.log 2>&1
head -3 gpurun_out/r2_e2e_profile.log; cat gpurun_out/r2_midn.log; tail -3 gpurun_out/r2_cfg5.log
grep -E '^\{' gpurun_out/r2_bench_final.log | cut -c1-600

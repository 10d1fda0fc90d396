#!/bin/bash
# GPU run (final build): GPU test tier, smoke, both bench arms, launch list + ncu --set full of the headline kernel,
# cfg4 kernel times, e2e breakdown.  Every ncu pass comes after the same command run plain (exit 0).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_final.log
tail -3 gpurun_out/r2_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2_bench_final.log 2>&1; echo "bench exit $?" >> gpurun_out/r2_bench_final.log
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_ref_final.log 2>&1; echo "bench exit $?" >> gpurun_out/r2_bench_ref_final.log
grep -E '^\{|exit' gpurun_out/r2_bench_final.log | cut -c1-250; grep -E '^\{|exit' gpurun_out/r2_bench_ref_final.log | cut -c1-250
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r2_bench_short.log 2>&1 \
 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r2_ncu_list.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fit_small -c 1 -o gpurun_out/prof_k1_r02 -f \
      python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r2_ncu_k1.log 2>&1
timeout 300 python tools/k3_time.py 10 > gpurun_out/r2_k4_time.log 2>&1; tail -3 gpurun_out/r2_k4_time.log | cut -c1-100
timeout 300 python tools/e2e_profile.py > gpurun_out/r2_e2e_profile.log 2>&1; head -4 gpurun_out/r2_e2e_profile.log | cut -c1-200

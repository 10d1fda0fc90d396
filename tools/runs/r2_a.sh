#!/bin/bash
# GPU run A: parity tests, a short bench, FP64 peak probes (incl. DFMA+DMMA interleaved), K4 timing.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests.log
tail -25 gpurun_out/r2_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/r2_bench.log
tail -3 gpurun_out/r2_bench.log | cut -c1-3000
timeout 120 python - > gpurun_out/r2_peaks.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
from qnmfits_b200._engine import get_engine
e = get_engine(0)
for kind in (0, 1, 3, 4, 10, 11, 14):
    print(kind, e.ctx.fp64_peak(kind, 4096))
PY
cat gpurun_out/r2_peaks.log
timeout 300 python tools/k3_time.py 10 > gpurun_out/r2_k4_time.log 2>&1; tail -5 gpurun_out/r2_k4_time.log

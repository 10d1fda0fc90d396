#!/bin/bash
# GPU run K: K1p configurations (lanes per row slice, rows per block) over N = 9 .. 16.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for lib in tools/_variants/libqnmfit_p*.so; do
  echo "== $lib"
  MIDN_ONLY_AUTO=1 QNMFIT_K1P_MIN=9 QNMFIT_LIB=$lib timeout 300 python tools/midn_time.py 10 11 12 13 14 15 16 2>&1 | tail -7
done > gpurun_out/r2_k1p_variants.log 2>&1
cat gpurun_out/r2_k1p_variants.log

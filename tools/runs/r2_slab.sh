#!/bin/bash
# Slab times of the headline grid under the 1 / 2 / 4 / 8-GPU plans, product build and K1 block-size variants.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
echo "== product"; timeout 300 python tools/slab_time.py 20
for lib in tools/_variants/libqnmfit_t*.so; do echo "== $lib lpf 4"; QNMFIT_K1_LPF=4 QNMFIT_LIB=$lib timeout 300 python tools/slab_time.py 20; done
} > gpurun_out/r2_slab.log 2>&1
cat gpurun_out/r2_slab.log | cut -c1-260

#!/bin/bash
# Developer tool: build libqnmfit with extra -D flags into tools/_variants/libqnmfit_<name>.so
#   tools/build_variant.sh name -DFOO=1 ...   ; run with QNMFIT_LIB=tools/_variants/libqnmfit_name.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/_variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared \
  -Iinclude -Iqnmfits_b200/csrc -DQNMFIT_ONLY_N8 "$@" -o tools/_variants/libqnmfit_$name.so qnmfits_b200/csrc/qnmfit_api.cu

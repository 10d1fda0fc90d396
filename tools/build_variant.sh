#!/bin/bash
# Developer tool: build libqnmfit with extra -D flags into tools/_variants/libqnmfit_<name>.so
#   tools/build_variant.sh name -DFOO=1 ...   ; run with QNMFIT_LIB=tools/_variants/libqnmfit_name.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/_variants
python - "$name" "$@" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
name, flags = sys.argv[1], sys.argv[2:]
ge.build_library(force=True, extra_flags=flags, lib=os.path.join("tools", "_variants", f"libqnmfit_{name}.so"),
                 build_dir=os.path.join("build", "variant_" + name))
PY

#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python - > gpurun_out/r2_latency.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
from qnmfits_b200._engine import get_engine
e = get_engine(0)
for kind in (40, 41, 42, 21, 31):
    print(kind, round(e.ctx.fp64_peak(kind, 512), 2))
PY
cat gpurun_out/r2_latency.log

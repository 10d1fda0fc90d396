import sys, os
sys.path.insert(0, os.getcwd())
from qnmfits_b200._engine import get_engine
eng = get_engine(0)
names = {0: 'dfma reuse (8 chains)', 2: 'dfma 3 distinct reads', 3: 'dfma 2 reads + reuse', 1: 'dmma m8n8k4'}
for k, n in names.items():
    hi = eng.ctx.fp64_peak(k, 4096)
    lo = eng.ctx.fp64_peak(10 + k, 4096)
    print(f"{n:26s} 8 CTAs/SM: {hi:6.2f} TF   1 CTA/SM (2 warps/scheduler): {lo:6.2f} TF  ({100*lo/37.2:.1f} % of 37.2)")

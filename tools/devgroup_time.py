#!/usr/bin/env python
"""Developer tool (multi-GPU box): mismatch_M_chi_grid (cfg3, 256 x 256) from ONE process
driving 1..N GPUs (qnmfits_b200.use_devices), wall clock per call (repeated calls: the
prepared sweep of the group is found and only times + data travel)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qnmfits_b200 as qf  # noqa: E402
from qnmfits_b200 import workloads  # noqa: E402

workloads.use_synthetic_tables()
wl = workloads.config3(res=256)
args = (wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0)
kw = dict(T=wl.T, res=256)
out = {}
ref = None
for n in [k for k in (1, 2, 4, 8) if k <= torch.cuda.device_count()]:
    qf.use_devices(list(range(n)) if n > 1 else None)
    for _ in range(5):
        grid = qf.mismatch_M_chi_grid(*args, **kw)
    t = time.perf_counter()
    for _ in range(30):
        grid = qf.mismatch_M_chi_grid(*args, **kw)
    dt = (time.perf_counter() - t) / 30
    if ref is None:
        ref = grid
    out[n] = dict(ms=dt * 1e3, fits_per_s=65536 / dt, max_abs_diff_vs_1gpu=float(abs(grid - ref).max()))
    print(n, json.dumps(out[n]), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "devgroup_time.json"), "w"), indent=1)

#!/usr/bin/env python
"""Developer tool (multi-GPU box): where the wall clock of a repeated single-process
multi-GPU mismatch_M_chi_grid call goes (cfg3, 256 x 256): the stages of
_DeviceGroupSweep.rerun, timed by wrapping its calls."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qnmfits_b200 as qf  # noqa: E402
from qnmfits_b200 import workloads, _cabi  # noqa: E402
from qnmfits_b200 import qnmfits as api  # noqa: E402

workloads.use_synthetic_tables()
wl = workloads.config3(res=256)
args = (wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0)
kw = dict(T=wl.T, res=256)
n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
qf.use_devices(list(range(n)))
for _ in range(5):
    qf.mismatch_M_chi_grid(*args, **kw)

marks = {}


def timed(name, fn):
    def wrapper(*a, **k):
        t = time.perf_counter()
        out = fn(*a, **k)
        marks.setdefault(name, []).append(time.perf_counter() - t)
        return out
    return wrapper


_cabi.Context.run_host = timed("run_host (enqueue, per device)", _cabi.Context.run_host)
_cabi.Context.stream_sync = timed("stream_sync (per device)", _cabi.Context.stream_sync)
np_copyto = np.copyto
api.np.copyto = timed("np.copyto (staging, slabs)", np_copyto)
api._DeviceGroupSweep.rerun = timed("rerun (whole)", api._DeviceGroupSweep.rerun)
api._cached_sweep = timed("_cached_sweep", api._cached_sweep)
api._problem_key = timed("_problem_key", api._problem_key)
reps = 50
t = time.perf_counter()
for _ in range(reps):
    qf.mismatch_M_chi_grid(*args, **kw)
total = (time.perf_counter() - t) / reps
print(f"{n} devices: {total * 1e6:.0f} us per call")
for name, v in marks.items():
    v = np.array(v) * 1e6
    print(f"  {name:34s} {len(v) / reps:4.1f} x {np.median(v):7.1f} us (median)  sum per call {v.sum() / reps:7.1f} us")

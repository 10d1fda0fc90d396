#!/usr/bin/env python
"""Developer tool (GPU box): cfg4 through the public API — per-call wall clock of repeated calls, the
number of fits the device flags, and where a repeated call spends its time."""
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
wl = workloads.config4()
call = lambda: qf.mismatch_t0_array(wl.times, wl.data, wl.modes, wl.Mf, wl.chif, wl.t0_array, T_array=wl.T,  # noqa: E731
                                    spherical_modes=wl.spherical_modes)
for i in range(6):
    t = time.perf_counter()
    mm = call()
    torch.cuda.synchronize()
    print("call %d: %.3f ms" % (i, (time.perf_counter() - t) * 1e3), flush=True)
print("cache entries", len(api._sweep_cache))
marks = {}


def timed(name, fn):
    def wrapper(*a, **k):
        t = time.perf_counter()
        out = fn(*a, **k)
        marks.setdefault(name, []).append((time.perf_counter() - t) * 1e3)
        return out
    return wrapper


_cabi.Context.run_host = timed("run_host", _cabi.Context.run_host)
api._Sweep.rerun = timed("rerun", api._Sweep.rerun)
api._Sweep._finish_single = timed("_finish_single", api._Sweep._finish_single)
api._repair_rank_deficient = timed("_repair_rank_deficient", api._repair_rank_deficient)
api._problem_key = timed("_problem_key", api._problem_key)
api._cached_sweep = timed("_cached_sweep", api._cached_sweep)
api._data_rows = timed("_data_rows", api._data_rows)
api._prepare_t0_sweep = timed("_prepare_t0_sweep", api._prepare_t0_sweep)
for _ in range(5):
    call()
for k, v in marks.items():
    print("  %-24s %d x median %.3f ms" % (k, len(v), float(np.median(v))))

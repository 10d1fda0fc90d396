#!/usr/bin/env python
"""Developer tool: wall-clock timings of BASELINE.json configs 1, 2 and 4 through the
public API on one GPU, next to the oracle (CPU) on a bounded sample."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qnmfits_b200 as qf  # noqa: E402
from qnmfits_b200 import workloads, synthetic  # noqa: E402
from oracle import qnmfits_oracle as orc  # noqa: E402

workloads.use_synthetic_tables()
tables = orc.OracleTables(synthetic.modes_cache)
out = {}


def timeit(fn, n=5):
    """Median wall clock of n calls after two warm-up calls (the first prepares the sweep)."""
    fn()
    fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(n):
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t)
    return float(np.median(times)), r


wl = workloads.config1()
dt, fit = timeit(lambda: qf.ringdown_fit(wl.times, wl.data, wl.modes, wl.Mf, wl.chif, 0.0), 20)
t = time.perf_counter()
for _ in range(50):
    ref = orc.ringdown_fit(tables, wl.times, wl.data, wl.modes, wl.Mf, wl.chif, 0.0)
cpu = (time.perf_counter() - t) / 50
out["cfg1"] = dict(gpu_ms=dt * 1e3, cpu_ms=cpu * 1e3,
                   relC=float(np.max(np.abs(fit["C"] - ref["C"])) / np.max(np.abs(ref["C"]))),
                   dmm=float(abs(fit["mismatch"] - ref["mismatch"])), cond=float(ref["s"][0] / ref["s"][-1]))

wl = workloads.config2()
dt, mm = timeit(lambda: qf.mismatch_t0_array(wl.times, wl.data, wl.modes, wl.Mf, wl.chif, wl.t0_array), 10)
t = time.perf_counter()
want = orc.mismatch_t0_array(tables, wl.times, wl.data, wl.modes, wl.Mf, wl.chif, wl.t0_array[::10])
cpu = (time.perf_counter() - t) * 10
out["cfg2"] = dict(gpu_ms=dt * 1e3, fits=1000, gpu_fits_per_s=1000 / dt, cpu_s_scaled=cpu,
                   max_dmm=float(np.max(np.abs(np.array(mm)[::10] - want))))

wl = workloads.config4()
dt, mm = timeit(lambda: qf.mismatch_t0_array(wl.times, wl.data, wl.modes, wl.Mf, wl.chif, wl.t0_array,
                                             T_array=wl.T, spherical_modes=wl.spherical_modes), 9)
t = time.perf_counter()
idx = [0, 250, 499]
want = orc.mismatch_t0_array(tables, wl.times, wl.data, wl.modes, wl.Mf, wl.chif, wl.t0_array[idx],
                             T_array=wl.T, spherical_modes=wl.spherical_modes)
cpu = (time.perf_counter() - t) / 3
f1 = orc.multimode_ringdown_fit(tables, wl.times, wl.data, wl.modes, wl.Mf, wl.chif, 0.0, T=wl.T,
                                spherical_modes=wl.spherical_modes)
out["cfg4"] = dict(gpu_ms=dt * 1e3, fits=500, gpu_fits_per_s=500 / dt, cpu_s_per_fit=cpu,
                   max_dmm=float(np.max(np.abs(np.array(mm)[idx] - want))), rows=21 * 1000, modes=len(wl.modes))
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "other_configs.json"), "w"), indent=1)

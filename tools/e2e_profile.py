#!/usr/bin/env python
"""Developer tool (GPU box): where the host time of one mismatch_M_chi_grid call goes
(cProfile + wall clock of the stages) next to the kernel time."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qnmfits_b200 as qf  # noqa: E402
from qnmfits_b200 import workloads  # noqa: E402
from qnmfits_b200 import qnmfits as api  # noqa: E402

workloads.use_synthetic_tables()
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
wl = workloads.config3(res=res)
args = (wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0)
kw = dict(T=wl.T, res=res)
for _ in range(5):
    qf.mismatch_M_chi_grid(*args, **kw)
n = 50
t = time.perf_counter()
for _ in range(n):
    qf.mismatch_M_chi_grid(*args, **kw)
print("call: %.1f us" % ((time.perf_counter() - t) / n * 1e6))
tp = tl = tf = 0.0
for _ in range(n):
    t0 = time.perf_counter()
    sweep, shape = api._prepare_M_chi_grid(*args, **kw)
    t1 = time.perf_counter()
    sweep.launch()
    t2 = time.perf_counter()
    sweep.fetch()
    t3 = time.perf_counter()
    tp += t1 - t0; tl += t2 - t1; tf += t3 - t2
print("prepare %.1f us, launch %.1f us, fetch (incl. kernel wait) %.1f us" % (tp / n * 1e6, tl / n * 1e6, tf / n * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    qf.mismatch_M_chi_grid(*args, **kw)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)

# ---- the other sweeps: cfg2 (1000 start times) and cfg4 (multimode, 500 start times)
for name, wlx in (("cfg2", workloads.config2()), ("cfg4", workloads.config4())):
    a = (wlx.times, wlx.data, wlx.modes, wlx.Mf, wlx.chif, wlx.t0_array)
    for _ in range(3):
        qf.mismatch_t0_array(*a)
    t = time.perf_counter()
    for _ in range(20):
        qf.mismatch_t0_array(*a)
    print("%s call: %.1f us" % (name, (time.perf_counter() - t) / 20 * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        qf.mismatch_t0_array(*a)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(12)
wl1 = workloads.config1()
a = (wl1.times, wl1.data, wl1.modes, 0.95, 0.69, 0.0)
for _ in range(3):
    qf.ringdown_fit(*a)
t = time.perf_counter()
for _ in range(50):
    qf.ringdown_fit(*a)
print("cfg1 ringdown_fit call: %.1f us" % ((time.perf_counter() - t) / 50 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    qf.ringdown_fit(*a)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)

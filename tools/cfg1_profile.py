#!/usr/bin/env python
"""Developer tool (GPU box): where a single ``ringdown_fit`` call (config 1) and a repeated
``mismatch_t0_array`` call (config 2) spend their time on the host (cProfile; the C call
``run_host`` holds upload + kernel + download + synchronisation)."""
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

workloads.use_synthetic_tables()
wl = workloads.config1()
w2 = workloads.config2()
for name, fn, reps in (
        ("cfg1 ringdown_fit", lambda: qf.ringdown_fit(wl.times, wl.data, wl.modes, wl.Mf, wl.chif, 0.0), 300),
        ("cfg2 mismatch_t0_array", lambda: qf.mismatch_t0_array(w2.times, w2.data, w2.modes, w2.Mf, w2.chif, w2.t0_array), 100)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    print(f"{name}: {(time.perf_counter() - t) / reps * 1e6:.1f} us per call (plain)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(reps):
        fn()
    pr.disable()
    print(f"--- {name}, {reps} calls under cProfile")
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)

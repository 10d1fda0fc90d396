#!/usr/bin/env python
"""Developer tool (GPU box): cProfile of the batched free-frequency search (host overhead)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qnmfits_b200 as qf  # noqa: E402
from qnmfits_b200 import workloads  # noqa: E402

workloads.use_synthetic_tables()
wl = workloads.config5(n_waveforms=4096, n_fixed=2)
qf.free_frequency_fit_batch(wl.times, wl.data[:64], 0.0, modes=wl.modes, Mf=wl.Mf, chif=wl.chif)
pr = cProfile.Profile()
pr.enable()
qf.free_frequency_fit_batch(wl.times, wl.data, 0.0, modes=wl.modes, Mf=wl.Mf, chif=wl.chif)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

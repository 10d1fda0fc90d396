#!/usr/bin/env python
"""Developer tool: median durations of the stages of K4's pipeline (CTA 0) from the trace of a
-DK4_TRACE build (tools/k4_trace.py prints the full timeline)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from qnmfits_b200 import workloads, _cabi  # noqa: E402
from qnmfits_b200 import qnmfits as api  # noqa: E402

workloads.use_synthetic_tables()
n_fits = int(sys.argv[1]) if len(sys.argv) > 1 else 148
wl = workloads.config4(n_t0=n_fits)
sweep = api._prepare_t0_sweep(np.asarray(wl.times), wl.data, wl.modes, wl.Mf, wl.chif,
                              np.asarray(wl.t0_array, dtype=float), 'geq', wl.T * np.ones(n_fits), wl.spherical_modes, 0.0)
sweep.batch.kernel = _cabi.KERNEL_PANEL
lib = _cabi.load_library()
buf = (C.c_longlong * (4 * 2048))()
sweep.eng.fit(sweep.batch); torch.cuda.synchronize()
lib.qnmfit_debug_trace(buf, 2048)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sweep.eng.fit(sweep.batch); e1.record(); torch.cuda.synchronize()
n = lib.qnmfit_debug_trace(buf, 2048)
ev = np.array(buf[:4 * n], dtype=np.int64).reshape(n, 4)
ev = ev[np.argsort(ev[:, 0], kind='stable')]
last, dur = {}, {}
for t, w, tag, pi in ev:
    w, tag = int(w), int(tag)
    if w in last:
        dur.setdefault((last[w][1], tag), []).append(int(t - last[w][0]))
    last[w] = (t, tag)
names = {(0, 1): "V load", (8, 1): "to next panel", (1, 2): "chunk 0 (split)", (2, 3): "barrier A", (3, 4): "panel", (3, 5): "chunks 1..", (4, 8): "barrier B (panel warp)",
         (5, 8): "barrier B (update warps)", (6, 7): "last panel's chunks", (8, 0): "to next panel", (8, 6): "to last panel"}
print(os.environ.get("QNMFIT_LIB", "default"), "kernel ms %.3f" % e0.elapsed_time(e1))
for key, name in names.items():
    if key in dur:
        v = np.array(dur[key])
        print(f"  {name:28s} median {int(np.median(v)):6d}  min {v.min():6d}  max {v.max():6d}  n {len(v)}")

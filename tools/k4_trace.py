#!/usr/bin/env python
"""Developer tool (GPU box): per-warp timeline of K4's panel pipeline for CTA 0 (one cfg4 fit),
from the clock stamps of a -DK4_TRACE build:
    tools/build_variant.sh k4trace -DK4_TRACE
    QNMFIT_LIB=tools/_variants/libqnmfit_k4trace.so python tools/k4_trace.py"""
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
n_fits = int(sys.argv[1]) if len(sys.argv) > 1 else 296
wl = workloads.config4(n_t0=n_fits)
sweep = api._prepare_t0_sweep(np.asarray(wl.times), wl.data, wl.modes, wl.Mf, wl.chif,
                              np.asarray(wl.t0_array, dtype=float), 'geq', wl.T * np.ones(n_fits), wl.spherical_modes, 0.0)
sweep.batch.kernel = _cabi.KERNEL_PANEL
lib = _cabi.load_library()
buf = (C.c_longlong * (4 * 2048))()
sweep.eng.fit(sweep.batch); torch.cuda.synchronize()
lib.qnmfit_debug_trace(buf, 2048)                       # discard the warm-up launch
sweep.eng.fit(sweep.batch); torch.cuda.synchronize()
n = lib.qnmfit_debug_trace(buf, 2048)
ev = np.array(buf[:4 * n], dtype=np.int64).reshape(n, 4)
ev = ev[np.argsort(ev[:, 0], kind='stable')]
t0 = ev[0, 0]
names = {0: 'iter', 1: 'V loaded', 2: 'chunk0 done', 3: 'past barrier A', 4: 'panel done', 5: 'chunks done',
         6: 'last: start', 7: 'last: chunks done', 8: 'past barrier B',
         10: 'refl: start', 11: 'refl: tail synced', 12: 'refl: dots+scalars+T', 13: 'refl: updated'}
print(f"{n} events")
# durations per (warp, iteration) for the first 3 tiles
last = {}
rows = []
for t, w, tag, pi in ev[:900]:
    key = int(w)
    dt = t - last.get(key, t)
    last[key] = t
    rows.append((int(t - t0), int(w), int(pi), names[int(tag)], int(dt)))
for r in rows[:420]:
    print("%8d  warp %d  panel %d  %-18s +%d" % r)

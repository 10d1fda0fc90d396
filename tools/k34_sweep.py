#!/usr/bin/env python
"""Developer tool (GPU box): kernel time of K3 vs K4 (vs K2 beyond 64 columns) over a set of
(N modes, L series) shapes, 296 fits of 1000 rows each.  Output -> gpurun_out/k34_sweep.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from qnmfits_b200 import _cabi  # noqa: E402
from qnmfits_b200._engine import get_engine  # noqa: E402

eng = get_engine()
rng = np.random.default_rng(0)
K_tot, B = 1100, 296
times = np.arange(K_tot) * 0.1
out = {}
for N, L in [(13, 1), (16, 1), (24, 1), (32, 1), (48, 1), (63, 1), (8, 2), (16, 5), (24, 10), (32, 21), (40, 21), (43, 21),
             (44, 21), (64, 21)]:
    freq = np.linspace(-0.1 * N, 0.1 * N, N) + 0.013 * rng.standard_normal(N) - 1j * (0.02 + 0.06 * rng.random(N))
    coef = rng.standard_normal((L, N)) + 1j * rng.standard_normal((L, N))
    C = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    E = np.exp(-1j * np.outer(times, freq))
    data = np.stack([E @ (coef[i] * C) for i in range(L)]) + 1e-5 * rng.standard_normal((L, K_tot))
    rb = rng.integers(0, 90, B).astype(np.int32)
    d = dict(times_d=eng.to_device(times, np.float64), data_d=eng.to_device(data, np.complex128),
             omega_d=eng.to_device(freq.reshape(1, -1), np.complex128), omega_shared=True,
             n_fits=B, n_modes=N, n_series=L, row_begin_all=0, row_end_all=K_tot,
             row_begin_d=eng.to_device(rb, np.int32), row_end_d=eng.to_device(rb + 1000, np.int32),
             t0_d=eng.to_device(times[rb], np.float64), dt_nominal=0.1, uniform_weights=True)
    if L > 1:
        d.update(coef_d=eng.to_device(coef.reshape(1, L, N), np.complex128), n_coef=1,
                 coef_index_d=eng.to_device(np.zeros(B, np.int32), np.int32))
    row = {}
    ref = None
    for name, kid in (("k4", _cabi.KERNEL_PANEL), ("k3", _cabi.KERNEL_STRUCT), ("k2", _cabi.KERNEL_GENERAL)):
        if name == "k3" and N + L > 64:
            continue
        if name == "k2" and N + L <= 64:
            continue
        mm = eng.empty((B,), torch.float64)
        b = eng.make_batch(kernel=kid, mismatch_d=mm, **d)
        for _ in range(2):
            eng.fit(b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.fit(b)
        e1.record()
        torch.cuda.synchronize()
        row[name + "_ms"] = e0.elapsed_time(e1) / 5
        got = eng.to_host(mm)
        if ref is None:
            ref = got
        else:
            row[name + "_vs_k4"] = float(np.max(np.abs(got - ref)))
    out[f"{N}x{L}"] = row
    print(N, L, json.dumps(row), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "k34_sweep.json"), "w"), indent=1)

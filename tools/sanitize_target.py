"""Small run of every kernel family for compute-sanitizer (tools/sanitize.sh):
K1 (staged grid, per-fit windows, free-frequency series_index), K3 (multimode sweep),
K2 (dynamic multimode fit), single fits with model output, rank-deficient repair."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import qnmfits_b200 as qf                      # noqa: E402
from qnmfits_b200 import workloads             # noqa: E402

workloads.use_synthetic_tables()
wl = workloads.config3(res=6)
g = qf.mismatch_M_chi_grid(wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0, T=wl.T, res=6)
print("K1 grid", g.shape, float(g.min()))
wl2 = workloads.config2(n_t0=40)
print("K1 t0 sweep", float(np.min(qf.mismatch_t0_array(wl2.times, wl2.data, wl2.modes, wl2.Mf, wl2.chif, wl2.t0_array))))
fit = qf.ringdown_fit(wl.times, wl.data, wl.modes, 0.95, 0.69, 0.0)
print("K1 single", float(fit["mismatch"]))
fit = qf.ringdown_fit(wl.times, wl.data, [(2, 2, n, 1) for n in range(12)], 0.95, 0.69, 0.0)
print("K1 N=12", float(fit["mismatch"]))
import cases                                   # noqa: E402
wl4 = cases.cfg4_small()
print("K3 multimode sweep", float(np.min(qf.mismatch_t0_array(wl4.times, wl4.data, wl4.modes, wl4.Mf, wl4.chif,
                                                            wl4.t0_array))))
wq = workloads.config4(n_t0=2, quadratic=True)
print("K3/K4 cfg4 quadratic", qf.mismatch_t0_array(wq.times, wq.data, wq.modes, wq.Mf, wq.chif, wq.t0_array,
                                                   T_array=wq.T, spherical_modes=wq.spherical_modes,
                                                   coef_columns=wq.extra["coef_columns"]))
w5 = workloads.config5(n_waveforms=5, n_fixed=1)
print("K1 free frequency", qf.free_frequency_fit_batch(w5.times, w5.data, 0.0, modes=w5.modes, Mf=w5.Mf, chif=w5.chif)[:2])
K = len(wl.times)
Mf_t = 0.95 + 0.001 * np.tanh((wl.times - 10) / 20)
chi_t = 0.69 - 0.002 * np.tanh((wl.times - 10) / 20)
d = {(2, 2): wl.data, (3, 2): 0.1 * wl.data}
fit = qf.dynamic_multimode_ringdown_fit(wl.times, d, [(2, 2, 0, 1), (2, 2, 1, 1), (3, 2, 0, 1)], Mf_t, chi_t, 5.0, T=60)
print("K2 dynamic multimode", float(fit["mismatch"]))
dup = qf.mismatch_t0_array(wl2.times, wl2.data, [(2, 2, n, 1) for n in (0, 1, 9, 10)], wl2.Mf, wl2.chif, wl2.t0_array[:9])
print("repair path", float(np.min(dup)))
print("sanitize target done")

"""
Generate tests/golden/dynamic.npz: the UNMODIFIED reference's dynamic_ringdown_fit
(/root/reference/qnmfits/qnmfits.py:318-475), dynamic_multimode_ringdown_fit (:676-911) and the
dynamic branch of mismatch_t0_array (:1286-1299), through oracle/ref_loader.py, with mass and
spin drifting smoothly towards their final values.  Build container only.
    python tests/golden/make_golden_dynamic.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")

from oracle.ref_loader import load_reference  # noqa: E402
from qnmfits_b200 import workloads  # noqa: E402
import cases  # noqa: E402

ref = load_reference()
workloads.use_synthetic_tables()


def drift(times):
    """Mf(t), chif(t): relax exponentially to (0.95, 0.69)."""
    x = np.exp(-np.clip(times, 0, None) / 15.0)
    return 0.95 - 0.03 * x, 0.69 - 0.05 * x


out = {}
wl = workloads.config1()
Mf_t, chi_t = drift(wl.times)
fit = ref.dynamic_ringdown_fit(wl.times, wl.data, wl.modes[:5], Mf_t, chi_t, 2.0, T=70)
for k in ("C", "mismatch", "residual", "model", "frequencies"):
    out["single_" + k] = np.asarray(fit[k])
fit = ref.dynamic_ringdown_fit(wl.times, wl.data, wl.modes[:3], 0.95, chi_t, 3.37, t0_method='closest', T=50)
out["single_closest_C"] = fit["C"]
out["single_closest_mismatch"] = np.asarray(fit["mismatch"])
t0s = np.linspace(-5.0, 30.0, 9)
out["sweep_single"] = np.array(ref.mismatch_t0_array(wl.times, wl.data, wl.modes[:5], Mf_t, chi_t, t0s, T_array=60))

wl4 = cases.cfg4_small()
Mf4, chi4 = drift(wl4.times)
# the reference's reshaping of mu (qnmfits.py:846) only works when no coefficient is the
# scalar 0 that qnm.mu returns for m' != m (qnm.py:336-337): same-m modes and series only
DYN_SPH = [(2, 2), (3, 2), (4, 2)]
DYN_MODES = [(2, 2, 0, 1), (2, 2, 1, 1), (3, 2, 0, 1), (2, 2, 0, -1), (4, 2, 0, 1)]
fit = ref.dynamic_multimode_ringdown_fit(wl4.times, wl4.data, DYN_MODES, Mf4, chi4, 5.0, T=80,
                                         spherical_modes=DYN_SPH)
out["multi_C"] = fit["C"]
out["multi_mismatch"] = np.asarray(fit["mismatch"])
out["multi_residual"] = np.asarray(fit["residual"])
out["multi_frequencies_shape"] = np.array(fit["frequencies"].shape)
lm = DYN_SPH[1]
out["multi_model_1"] = fit["model"][lm]
out["multi_weighted_1"] = fit["weighted_C"][lm]
out["sweep_multi"] = np.array(ref.mismatch_t0_array(wl4.times, wl4.data, DYN_MODES, Mf4, chi4, wl4.t0_array,
                                                    T_array=70, spherical_modes=DYN_SPH))
out["t0s"] = t0s
for k, v in out.items():
    print(k, np.asarray(v).shape)
path = os.path.join(HERE, "dynamic.npz")
np.savez_compressed(path, **out)
print("dynamic:", os.path.getsize(path), "bytes")

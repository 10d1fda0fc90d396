"""
Generate tests/golden/cfg5.npz: the UNMODIFIED reference's free_frequency_fit
(/root/reference/qnmfits/qnmfits.py:1905-2043, through oracle/ref_loader.py) on the first
waveforms of workloads.config5, with 2 / 1 / 0 fixed modes.  Build container only.
    python tests/golden/make_golden_cfg5.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle.ref_loader import load_reference  # noqa: E402
from qnmfits_b200 import workloads  # noqa: E402

ref = load_reference()
workloads.use_synthetic_tables()

out = {}
for n_fixed, n_wf in ((2, 6), (1, 3), (0, 3)):
    wl = workloads.config5(n_waveforms=n_wf, n_fixed=n_fixed)
    best = [ref.free_frequency_fit(wl.times, wl.data[b], 0.0, modes=wl.modes, Mf=wl.Mf, chif=wl.chif)
            for b in range(n_wf)]
    out[f"fixed{n_fixed}_omega"] = np.array(best)
    out[f"fixed{n_fixed}_truth"] = wl.extra["omega_free"]
    print(n_fixed, np.abs(np.array(best) - wl.extra["omega_free"]))
# off-grid start, 'closest', shorter window
wl = workloads.config5(n_waveforms=2, n_fixed=1)
out["closest_omega"] = np.array([ref.free_frequency_fit(wl.times, wl.data[b], 3.37, modes=wl.modes, Mf=wl.Mf,
                                                        chif=wl.chif, t0_method='closest', T=60) for b in range(2)])
path = os.path.join(HERE, "cfg5.npz")
np.savez_compressed(path, **out)
print("cfg5:", os.path.getsize(path), "bytes")

"""
Generate tests/golden/next.npz: the UNMODIFIED reference's mismatch_omega_grid
(/root/reference/qnmfits/qnmfits.py:1679-1827) and calculate_epsilon (:1418-1594), through
oracle/ref_loader.py, on seeded synthetic inputs.  Build container only.
    python tests/golden/make_golden_next.py
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
ref.tqdm = lambda x, *a, **k: x          # silence the progress bar
workloads.use_synthetic_tables()

out = {}
wl = workloads.config1()
m2 = wl.modes[:2]
out["omega_grid_geq"] = ref.mismatch_omega_grid(wl.times, wl.data, m2, 0.95, 0.69, (0.2, 0.9), (-0.9, -0.1), 5.0,
                                                T=80, res=7)
out["omega_grid_nofixed"] = ref.mismatch_omega_grid(wl.times, wl.data, [], 0.95, 0.69, (0.3, 0.8), (-0.3, -0.05),
                                                    20.0, res=4)
out["omega_grid_closest"] = ref.mismatch_omega_grid(wl.times, wl.data, m2, 0.95, 0.69, (0.2, 0.9), (-0.9, -0.1),
                                                    3.37, t0_method='closest', T=60, res=5)
out["eps_single"] = np.array(ref.calculate_epsilon(wl.times, wl.data, wl.modes[:4], 0.95, 0.69, 10.0))
out["eps_single_x0_delta"] = np.array(ref.calculate_epsilon(wl.times, wl.data, wl.modes[:3], 0.95, 0.69, 15.0, T=70,
                                                            delta=[0.0, 0.01, 0.0], x0=[1.0, 0.6]))
wl4 = cases.cfg4_small()
out["eps_multimode"] = np.array(ref.calculate_epsilon(wl4.times, wl4.data, cases.MM_MODES, 0.95, 0.69, 5.0, T=80))
out["eps_multimode_x0"] = np.array(ref.calculate_epsilon(wl4.times, wl4.data, cases.MM_MODES, 0.95, 0.69, 5.0, T=80,
                                                         x0=[0.97, 0.65]))
for k, v in out.items():
    print(k, np.asarray(v).shape, np.asarray(v).ravel()[:3])
path = os.path.join(HERE, "next.npz")
np.savez_compressed(path, **out)
print("next:", os.path.getsize(path), "bytes")

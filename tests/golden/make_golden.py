"""
Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/qnmfits/{qnm,qnmfits}.py, imported through oracle/ref_loader.py)
on seeded synthetic inputs with the synthetic Kerr table provider.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The fixtures are committed; tests never regenerate them.
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


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path)} bytes")


def fit_arrays(prefix, fit):
    out = {prefix + k: np.asarray(fit[k]) for k in
           ("C", "mismatch", "residual", "frequencies", "model", "model_times")}
    if "rank" in fit:
        out[prefix + "rank"] = np.asarray(fit["rank"])
        out[prefix + "s"] = np.asarray(fit["s"])
    return out


# ---- provider values ---------------------------------------------------------
labels = [(2, 2, 0, 1), (2, 2, 7, 1), (2, 2, 3, -1), (3, 2, 1, 1), (4, 4, 2, 1), (2, 0, 1, 1),
          (2, -2, 0, 1), (3, -2, 1, -1), (2, 2, 9, 1), (2, 2, 10, 1), (2, 1, 11, 1),
          (2, 2, 0, 1, 2, 2, 0, 1), (2, 2, 0, 1, 3, 3, 0, 1, 2, 0, 1, -1)]
spins = np.array([0.0, 0.1, 0.5, 0.69, 0.7, 0.95, 0.99])
om = np.array([[ref.qnm.omega_list([lab], chi, 0.95)[0] for lab in labels] for chi in spins])
mu_idx = [(2, 2, 2, 2, 0, 1), (3, 2, 2, 2, 0, 1), (2, 2, 3, 2, 1, 1), (2, 2, 2, 2, 0, -1),
          (3, 2, 2, 2, 1, -1), (4, -2, 2, -2, 0, 1), (2, 2, 2, 1, 0, 1), (2, 0, 4, 0, 2, 1),
          (4, 4, 4, 4, 0, 1), (3, 3, 4, 3, 1, -1)]
mu = np.array([[complex(v) for v in ref.qnm.mu_list(mu_idx, chi)] for chi in spins])
save("provider", labels=np.array([str(l) for l in labels]), spins=spins, omega=om,
     mu_indices=np.array(mu_idx), mu=mu, Mf=np.array(0.95))

# ---- config 1: single fits and variants ---------------------------------------
wl = workloads.config1()
out = {"times": wl.times, "data": wl.data}
cases = {
    "base": dict(modes=wl.modes, Mf=0.95, chif=0.69, t0=0.0),
    "offgrid_geq": dict(modes=wl.modes, Mf=0.93, chif=0.66, t0=3.37, T=77.7),
    "closest": dict(modes=wl.modes, Mf=0.95, chif=0.69, t0=3.37, T=77.7, t0_method="closest"),
    "delta_float": dict(modes=wl.modes[:4], Mf=0.95, chif=0.69, t0=10.0, delta=0.01),
    "delta_list": dict(modes=wl.modes[:3], Mf=0.95, chif=0.69, t0=10.0, delta=[0.0, 0.01, -0.02]),
    "quadratic": dict(modes=[(2, 2, 0, 1), (2, 2, 1, 1), (2, 2, 0, 1, 2, 2, 0, 1)], Mf=0.95,
                      chif=0.69, t0=15.0),
    "mirror": dict(modes=[(2, 2, 0, 1), (2, 2, 0, -1), (2, 2, 1, 1), (2, 2, 1, -1)], Mf=0.95,
                   chif=0.69, t0=5.0),
    "one_mode": dict(modes=[(2, 2, 0, 1)], Mf=0.95, chif=0.69, t0=20.0),
    "twelve": dict(modes=[(2, 2, n, 1) for n in range(12)], Mf=0.95, chif=0.69, t0=0.0),
    "duplicate_label": dict(modes=[(2, 2, n, 1) for n in (0, 1, 9, 10)], Mf=0.95, chif=0.69,
                            t0=5.0),
}
for name, kw in cases.items():
    fit = ref.ringdown_fit(wl.times, wl.data, **kw)
    out.update(fit_arrays(name + "__", fit))
save("cfg1", **out)

# non-uniform time grid (direct-evaluation path)
rng = np.random.default_rng(7)
t_nu = np.sort(np.concatenate([np.arange(-100, 0) * 0.5,
                               np.cumsum(0.05 + 0.1 * rng.random(900))]))
omega_nu = np.array(ref.qnm.omega_list(wl.modes[:5], 0.69, 0.95))
C_nu = rng.normal(size=5) + 1j * rng.normal(size=5)
d_nu = ref.ringdown(t_nu, 0.0, C_nu, omega_nu) + 1e-7 * (rng.normal(size=len(t_nu))
                                                         + 1j * rng.normal(size=len(t_nu)))
fit = ref.ringdown_fit(t_nu, d_nu, wl.modes[:5], 0.95, 0.69, 1.0, T=60)
save("nonuniform", times=t_nu, data=d_nu, **fit_arrays("fit__", fit))

# ---- config 2: t0 sweep -------------------------------------------------------
wl2 = workloads.config2(n_t0=40)
mm = ref.mismatch_t0_array(wl2.times, wl2.data, wl2.modes, 0.95, 0.69, wl2.t0_array)
mm_closest = ref.mismatch_t0_array(wl2.times, wl2.data, wl2.modes[:4], 0.95, 0.69,
                                   wl2.t0_array[:10], t0_method="closest",
                                   T_array=np.linspace(50, 80, 10))
save("cfg2", t0_array=wl2.t0_array, mismatch=np.array(mm), mismatch_closest=np.array(mm_closest))

# ---- config 3: M-chi grid -----------------------------------------------------
wl3 = workloads.config3(res=12)
grid = ref.mismatch_M_chi_grid(wl3.times, wl3.data, wl3.modes, wl3.Mf_minmax, wl3.chif_minmax,
                               wl3.t0, T=wl3.T, res=12)
grid_q = ref.mismatch_M_chi_grid(wl3.times, wl3.data, cases["quadratic"]["modes"],
                                 wl3.Mf_minmax, wl3.chif_minmax, 15.0, T=60, res=5,
                                 delta=[0.0, 0.01, 0.0])
save("cfg3", grid=grid, grid_quadratic=grid_q)

# ---- config 4: multimode (small) ----------------------------------------------
sph = [(2, 2), (3, 2), (4, 2), (2, 0), (2, -2)]
mmodes = [(2, 2, 0, 1), (2, 2, 1, 1), (3, 2, 0, 1), (2, 2, 0, -1), (2, -2, 0, 1), (2, 0, 0, 1),
          (2, 0, 1, 1), (4, 2, 0, 1)]
wl4 = workloads.config4(n_t0=6, spherical=sph, modes=mmodes)
fit = ref.multimode_ringdown_fit(wl4.times, wl4.data, mmodes, 0.95, 0.69, 5.0, T=80)
out = {"C": fit["C"], "mismatch": np.asarray(fit["mismatch"]), "residual": fit["residual"],
       "frequencies": fit["frequencies"]}
for lm in sph:
    out[f"model_{lm[0]}_{lm[1]}"] = fit["model"][lm]
    out[f"weighted_C_{lm[0]}_{lm[1]}"] = fit["weighted_C"][lm]
    out[f"data_{lm[0]}_{lm[1]}"] = wl4.data[lm]
out["t0_sweep"] = np.array(ref.mismatch_t0_array(wl4.times, wl4.data, mmodes, 0.95, 0.69,
                                                 wl4.t0_array, T_array=70))
out["grid"] = ref.mismatch_M_chi_grid(wl4.times, wl4.data, mmodes, (0.9, 1.0), (0.6, 0.75), 5.0,
                                      T=80, res=4)
sub = [(2, 2), (3, 2)]
fit_sub = ref.multimode_ringdown_fit(wl4.times, wl4.data, mmodes[:4], 0.95, 0.69, 5.0, T=80,
                                     spherical_modes=sub)
out["sub_C"] = fit_sub["C"]
out["sub_mismatch"] = np.asarray(fit_sub["mismatch"])
out["t0_array"] = wl4.t0_array
save("cfg4", times=wl4.times, **out)

# ---- G1: injection / recovery (examples/correcting_measured_amplitude.ipynb) ---
t_g1 = np.linspace(0, 100, 500)
w = ref.qnm.omega_list([(2, 2, 0, 1)], 0.7, 1)
d_g1 = ref.ringdown(t_g1, 0.0, [1 - 1j], w)
f0 = ref.ringdown_fit(t_g1, d_g1, [(2, 2, 0, 1)], 1, 0.7, 0)
f10 = ref.ringdown_fit(t_g1, d_g1, [(2, 2, 0, 1)], 1, 0.7, 10)
save("g1", times=t_g1, data=d_g1, omega=np.array(w), C0=f0["C"], mm0=np.asarray(f0["mismatch"]),
     C10=f10["C"], mm10=np.asarray(f10["mismatch"]))

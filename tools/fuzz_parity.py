#!/usr/bin/env python
"""Developer tool (GPU box): randomised differential test of the public API against the oracle
(the reference's algorithm): random label sets (mirror, quadratic, duplicated), start times,
durations, t0_method, delta, uniform / jittered / non-uniform time grids; single fits, start-time
sweeps, Mf-chi grids, multimode.  Prints every case whose mismatch differs by more than 1e-10
(or whose amplitudes differ by more than the conditioning allows)."""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qnmfits_b200 as qf  # noqa: E402
from qnmfits_b200 import workloads, synthetic  # noqa: E402
from oracle import qnmfits_oracle as orc  # noqa: E402

workloads.use_synthetic_tables()
tables = orc.OracleTables(synthetic.modes_cache)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 120
rng = np.random.default_rng(seed)
warnings.simplefilter("ignore")


def random_modes():
    n = int(rng.integers(1, 17))
    pool = [(2, 2, k, 1) for k in range(12)] + [(2, 2, k, -1) for k in range(3)] + \
        [(3, 2, k, 1) for k in range(3)] + [(2, 2, 0, 1, 2, 2, 0, 1), (2, 2, 0, 1, 2, 2, 1, 1), (2, 2, 9, 1), (2, 2, 10, 1)]
    idx = rng.choice(len(pool), size=min(n, len(pool)), replace=False)
    return [pool[i] for i in idx]


def random_times():
    kind = rng.integers(0, 4)
    base = np.arange(-300, 1501) * 0.1
    if kind == 0:
        return base
    if kind == 1:
        return base + rng.normal(scale=2e-13, size=base.size)          # recurrence, non-uniform weights
    if kind == 2:
        return np.sort(base + rng.uniform(-0.03, 0.03, size=base.size))  # direct evaluation
    return np.arange(-150, 700) * 0.2


bad = []
for case in range(n_cases):
    times = random_times()
    modes = random_modes()
    Mf, chif = float(rng.uniform(0.85, 1.05)), float(rng.uniform(0.3, 0.9))
    w = np.array(tables.omega_list([(2, 2, k, 1) for k in range(6)], 0.69, 0.95))
    amp = rng.normal(size=6) + 1j * rng.normal(size=6)
    data = np.where(times >= 0, (amp[None, :] * np.exp(-1j * w[None, :] * times[:, None])).sum(axis=1), 0)
    data = data + 1e-7 * (rng.normal(size=data.size) + 1j * rng.normal(size=data.size))
    method = 'geq' if rng.random() < 0.7 else 'closest'
    T = float(rng.uniform(20, 110))
    delta = 0.0 if rng.random() < 0.6 else list(rng.uniform(-0.02, 0.02, size=len(modes)))
    kind = rng.integers(0, 8)
    try:
        if kind == 0:
            t0 = float(rng.uniform(-5, 40))
            got = qf.ringdown_fit(times, data, modes, Mf, chif, t0, method, T, delta)
            want = orc.ringdown_fit(tables, times, data, modes, Mf, chif, t0, method, T, delta)
            dmm = abs(got['mismatch'] - want['mismatch'])
            extra = dict(rank=(int(got['rank']), int(want['rank'])))
        elif kind == 1:
            t0s = np.sort(rng.uniform(-5, 60, size=17))
            Ts = T if rng.random() < 0.5 else rng.uniform(20, 90, size=17)
            got = np.array(qf.mismatch_t0_array(times, data, modes, Mf, chif, t0s, method, Ts, None, delta))
            want = np.array(orc.mismatch_t0_array(tables, times, data, modes, Mf, chif, t0s, method, Ts, None, delta))
            dmm = float(np.max(np.abs(got - want)))
            extra = {}
        elif kind == 2:
            t0 = float(rng.uniform(0, 30))
            got = qf.mismatch_M_chi_grid(times, data, modes, (0.9, 1.0), (0.55, 0.8), t0, method, T, 4, None, delta)
            want = orc.mismatch_M_chi_grid(tables, times, data, modes, (0.9, 1.0), (0.55, 0.8), t0, method, T, 4, None, delta)
            dmm = float(np.max(np.abs(got - want)))
            extra = {}
        elif kind == 4:                                   # dynamic single fit (time-dependent spectrum)
            lin = [m for m in modes if len(m) == 4][:8] or [(2, 2, 0, 1)]
            t0 = float(rng.uniform(0, 30))
            x = np.exp(-np.clip(times, 0, None) / 15.0)
            Mf_t, chi_t = Mf - 0.03 * x, chif - 0.05 * x
            got = qf.dynamic_ringdown_fit(times, data, lin, Mf_t, chi_t, t0, method, T)
            want = orc.dynamic_ringdown_fit(tables, times, data, lin, Mf_t, chi_t, t0, method, T)
            dmm = abs(got['mismatch'] - want['mismatch'])
            extra = {}
        elif kind == 5:                                   # frequency grid around a fixed-mode set
            lin = [m for m in modes if len(m) == 4][:5]
            t0 = float(rng.uniform(0, 30))
            got = qf.mismatch_omega_grid(times, data, lin, Mf, chif, (0.2, 0.9), (-0.7, -0.05), t0, method, T, 5)
            want = orc.mismatch_omega_grid(tables, times, data, lin, Mf, chif, (0.2, 0.9), (-0.7, -0.05), t0, method, T, 5)
            dmm = float(np.max(np.abs(got - want)))
            extra = {}
        elif kind == 6:                                   # dynamic multimode fit
            sph = [(2, 2), (3, 2), (4, 2)]
            lin = [m for m in modes if len(m) == 4][:6] or [(2, 2, 0, 1)]
            dd = {lm: data * (0.4 ** i) * np.exp(0.3j * i) for i, lm in enumerate(sph)}
            t0 = float(rng.uniform(0, 30))
            x = np.exp(-np.clip(times, 0, None) / 15.0)
            Mf_t, chi_t = Mf - 0.03 * x, chif - 0.05 * x
            got = qf.dynamic_multimode_ringdown_fit(times, dd, lin, Mf_t, chi_t, t0, method, T, sph)
            want = orc.dynamic_multimode_ringdown_fit(tables, times, dd, lin, Mf_t, chi_t, t0, method, T, sph)
            dmm = abs(got['mismatch'] - want['mismatch'])
            extra = {}
        elif kind == 7:                                   # multimode sweep with quadratic labels (coef_columns)
            sph = [(2, 2), (3, 2), (4, 4), (4, 2)]
            mm_modes = ([m for m in modes if len(m) == 4][:7] or [(2, 2, 0, 1)]) + list(workloads.QUADRATIC_LABELS[:2])
            dd = {lm: data * (0.4 ** i) * np.exp(0.3j * i) for i, lm in enumerate(sph)}
            t0s = np.sort(rng.uniform(0, 40, size=9))
            got = np.array(qf.mismatch_t0_array(times, dd, mm_modes, Mf, chif, t0s, method, T, sph,
                                                coef_columns=workloads.quadratic_columns(sph)))
            want = np.array(orc.mismatch_t0_array(tables, times, dd, mm_modes, Mf, chif, t0s, method, T, sph,
                                                  coef_override=workloads.coef_override(sph, mm_modes, chif, tables)))
            dmm = float(np.max(np.abs(got - want)))
            extra = {}
        else:
            sph = [(2, 2), (3, 2), (4, 2)]
            mm_modes = [m for m in modes if len(m) == 4] or [(2, 2, 0, 1)]
            dd = {lm: data * (0.4 ** i) * np.exp(0.3j * i) for i, lm in enumerate(sph)}
            t0s = np.sort(rng.uniform(0, 40, size=9))
            got = np.array(qf.mismatch_t0_array(times, dd, mm_modes, Mf, chif, t0s, method, T, sph))
            want = np.array(orc.mismatch_t0_array(tables, times, dd, mm_modes, Mf, chif, t0s, method, T, sph))
            dmm = float(np.max(np.abs(got - want)))
            extra = {}
    except Exception as exc:                              # both sides should fail alike
        try:
            if kind == 0:
                orc.ringdown_fit(tables, times, data, modes, Mf, chif, t0, method, T, delta)
            failed_alike = False
        except Exception:
            failed_alike = True
        if not failed_alike or kind != 0:
            bad.append(dict(case=case, kind=int(kind), error=str(exc)[:120], modes=modes))
            print("ERR", bad[-1], flush=True)
        continue
    if not dmm < 1e-10:
        bad.append(dict(case=case, kind=int(kind), dmm=float(dmm), n_modes=len(modes), modes=modes, method=method, T=T,
                        delta=delta != 0.0, **extra))
        print("DIFF", bad[-1], flush=True)
print(f"seed {seed}: {n_cases} cases, {len(bad)} discrepancies")
json.dump(bad, open(os.path.join(ROOT, "gpurun_out", f"fuzz_{seed}.json"), "w"), indent=1, default=str)

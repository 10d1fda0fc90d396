#!/usr/bin/env python
"""Developer tool (GPU box): accuracy of K1 against the oracle as a function of the
re-anchoring interval, on a 24x24 sub-grid of cfg3 (same window, same data), and the
kernel time of the full 256x256 grid for each interval."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qnmfits_b200 import workloads
    from qnmfits_b200 import qnmfits as api
    from oracle import qnmfits_oracle as orc
    workloads.use_synthetic_tables()
    tables = orc.OracleTables(workloads.synthetic_modes_cache())
    res = 256
    wl = workloads.config3(res=res)
    sweep, shape = api._prepare_M_chi_grid(wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0,
                                           T=wl.T, res=res)
    eng = sweep.eng
    N = len(wl.modes)
    C_d = eng.empty((res * res, N), torch.complex128)
    sweep.batch.C = C_d.data_ptr()
    Mfs, chis = orc.grid_axes(wl.Mf_minmax, wl.chif_minmax, res)
    ref_C = np.zeros((res * res, N), complex)
    ref_mm = np.zeros(res * res)
    pick = np.sort(np.random.default_rng(5).choice(res * res, 300, replace=False))
    for i in pick:
        r = orc.ringdown_fit(tables, wl.times, wl.data, wl.modes, Mfs[i // res], chis[i % res], wl.t0, T=wl.T)
        ref_C[i] = r['C']
        ref_mm[i] = r['mismatch']
    ref_C, ref_mm = ref_C[pick], ref_mm[pick]
    big = workloads.config3(res=256)
    sweep_big, _ = api._prepare_M_chi_grid(big.times, big.data, big.modes, big.Mf_minmax, big.chif_minmax, big.t0,
                                           T=big.T, res=256)
    out = {}
    print(sweep.eng.ctx.plan(sweep.batch).lanes_per_fit)
    for anchor in (16, 32, 64, 128, 256, 512):
        sweep.batch.anchor_rows = anchor
        sweep.launch_kernel()
        mm, _ = sweep.fetch()
        C = eng.to_host(C_d)[pick]
        mm = mm[pick]
        dC = np.max(np.abs(C - ref_C), axis=1) / np.max(np.abs(ref_C), axis=1)
        dmm = np.abs(mm - ref_mm)
        sweep_big.batch.anchor_rows = anchor
        for _ in range(3):
            sweep_big.launch_kernel()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.fit(sweep_big.batch)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out[anchor] = dict(max_dC=float(dC.max()), median_dC=float(np.median(dC)), max_dmm=float(dmm.max()), ms=ms)
        print(f"anchor {anchor:4d}: max rel dC {dC.max():.3e}  median {np.median(dC):.3e}  max |dmm| {dmm.max():.3e}"
              f"   256^2 kernel {ms:.4f} ms", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "anchor_accuracy.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

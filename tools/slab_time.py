#!/usr/bin/env python
"""Developer tool (GPU box): kernel-only time of ONE rank's slab of the headline grid (256 x 256
points, N = 8) under the plan of a 1 / 2 / 4 / 8-GPU job (first_fit offset, plan_fits = the whole
sweep), CUDA events, inputs resident, L2 flushed between launches.  Prints one JSON line per job
size with the launch geometry, and checks the slab's bits against the single launch.
Output -> gpurun_out/slab_time.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qnmfits_b200 import workloads, _cabi, _engine
    from qnmfits_b200 import qnmfits as api
    import qnmfits_b200 as qf
    workloads.use_synthetic_tables()
    res = 256
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    eng = _engine.get_engine()
    wl = workloads.config3(res=res)
    Mf = np.linspace(*wl.Mf_minmax, res)
    chi = np.linspace(*wl.chif_minmax, res)
    table, ptr = qf.qnm.constituent_table(wl.modes, chi)
    win = api._window_rows(wl.times, 0.0, 100, "geq")
    d = dict(times_d=eng.to_device(wl.times, np.float64),
             data_d=eng.to_device(wl.data.reshape(1, -1), np.complex128),
             omega_tilde_d=eng.to_device(table, np.complex128), mode_ptr_d=eng.to_device(ptr, np.int32),
             inv_Mf_d=eng.to_device(1.0 / Mf, np.float64), n_chi=res, n_mf=res,
             n_constituents=table.shape[1], n_modes=8, row_begin_all=win[0], row_end_all=win[1],
             t0_all=0.0, dt_nominal=0.1)
    n = res * res
    full = eng.empty((n,), torch.float64)
    eng.fit(eng.make_batch(n_fits=n, mismatch_d=full, **d))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for world in (1, 2, 4, 8):
        per = n // world
        lo = per * (world - 1)
        part = eng.empty((per,), torch.float64)
        b = eng.make_batch(n_fits=per, first_fit=lo, plan_fits=n, mismatch_d=part, **d)
        plan = eng.ctx.plan(b)
        for _ in range(3):
            eng.fit(b)
        torch.cuda.synchronize()
        ms = []
        for _ in range(steps):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); eng.fit(b); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms = float(np.median(ms))
        flops = _cabi.flops_per_fit(win[1] - win[0], 8, 1, True) * per
        out[world] = dict(ms=ms, fits=per, grid=plan.grid, block=plan.block, lanes_per_fit=plan.lanes_per_fit,
                          tflops=flops / ms * 1e-9, job_fits_per_s=n / ms * 1e3,
                          identical=bool(torch.equal(part, full[lo:lo + per])))
        print(world, json.dumps(out[world]), flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "slab_time.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

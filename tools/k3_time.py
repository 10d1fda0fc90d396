#!/usr/bin/env python
"""Developer tool (GPU box): kernel-only time of config 4 (500 start times, 21 spherical
modes x 40 QNMs; --quadratic: with the four quadratic labels) for K4, K3 and (--k2) K2,
CUDA events, inputs resident.  Output -> gpurun_out/k3_time.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qnmfits_b200 import workloads, _cabi
    from qnmfits_b200 import qnmfits as api
    workloads.use_synthetic_tables()
    n_t0 = int(os.environ.get("K3_FITS", "500"))
    wl = workloads.config4(n_t0=n_t0, quadratic="--quadratic" in sys.argv)
    T_array = wl.T * np.ones(len(wl.t0_array))
    sweep = api._prepare_t0_sweep(np.asarray(wl.times), wl.data, wl.modes, wl.Mf, wl.chif,
                                  np.asarray(wl.t0_array, dtype=float), 'geq', T_array, wl.spherical_modes, 0.0,
                                  wl.extra.get("coef_columns"))
    eng = sweep.eng
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    kernels = [("k4", _cabi.KERNEL_PANEL), ("k3", _cabi.KERNEL_STRUCT)] + \
        ([("k2", _cabi.KERNEL_GENERAL)] if "--k2" in sys.argv else [])
    results = {}
    out = {}
    for name, kid in kernels:
        sweep.batch.kernel = kid
        for _ in range(2):
            eng.fit(sweep.batch)
        torch.cuda.synchronize()
        results[name] = sweep.fetch()[0]
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            eng.fit(sweep.batch)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        plan = eng.ctx.plan(sweep.batch)
        N, L, K = len(wl.modes), len(wl.spherical_modes), 1000
        f_struct = 8 * K * N * N + 16 * K * N * L + (8 / 3) * L * N ** 3 + 6 * K * N
        f_dense = _cabi.flops_per_fit(K, N, L, False)
        out[name] = dict(ms=ms, fits=sweep.n_fits, fits_per_s=sweep.n_fits / ms * 1e3, regs=plan.regs_per_thread,
                         smem=plan.smem_bytes, lanes_per_column=plan.lanes_per_fit,
                         tflops_structured=f_struct * sweep.n_fits / ms * 1e-9,
                         tflops_dense_equivalent=f_dense * sweep.n_fits / ms * 1e-9)
        print(name, json.dumps(out[name]), flush=True)
    for name in results:
        out[name]["max_abs_diff_vs_" + kernels[-1][0]] = float(np.max(np.abs(results[name] - results[kernels[-1][0]])))
    print(json.dumps({k: v["max_abs_diff_vs_" + kernels[-1][0]] for k, v in out.items()}))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "k3_time.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Developer tool (GPU box): config 5 — 4096 free-frequency searches in lock step —
wall clock through the public API, next to the oracle (scipy Nelder-Mead + numpy lstsq,
the reference's algorithm) on a sample of the same waveforms."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import qnmfits_b200 as qf
    from qnmfits_b200 import workloads, synthetic
    from oracle import qnmfits_oracle as orc
    workloads.use_synthetic_tables()
    tables = orc.OracleTables(synthetic.modes_cache)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    wl = workloads.config5(n_waveforms=n, n_fixed=2)
    qf.free_frequency_fit_batch(wl.times, wl.data[:64], 0.0, modes=wl.modes, Mf=wl.Mf, chif=wl.chif)
    torch.cuda.synchronize()
    t = time.perf_counter()
    got, res = qf.free_frequency_fit_batch(wl.times, wl.data, 0.0, modes=wl.modes, Mf=wl.Mf, chif=wl.chif,
                                           return_result=True)
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t
    sample = np.arange(0, n, max(1, n // 12))[:12]
    t = time.perf_counter()
    ref = np.array([orc.free_frequency_fit(tables, wl.times, wl.data[b], 0.0, modes=wl.modes, Mf=wl.Mf, chif=wl.chif)
                    for b in sample])
    cpu_s = (time.perf_counter() - t) / len(sample)
    out = dict(waveforms=n, gpu_s=gpu_s, launches=int(res.launches), evaluations=int(res.nfev.sum()),
               nit_max=int(res.nit.max()), nfev_mean=float(res.nfev.mean()),
               cpu_s_per_waveform=cpu_s, cpu_s_scaled=cpu_s * n, speedup_vs_1_core=cpu_s * n / gpu_s,
               max_abs_diff_vs_oracle=float(np.max(np.abs(got[sample] - ref))),
               max_err_vs_truth=float(np.max(np.abs(got - wl.extra["omega_free"]))),
               fits_per_s=float(res.nfev.sum() / gpu_s))
    print(json.dumps(out, indent=1))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cfg5_time.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

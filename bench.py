#!/usr/bin/env python
"""
bench.py — BASELINE.json's headline metric: QNM least-squares fits per second on the
256 x 256 Mf-chif grid (8 overtones on h22, M = 1000 rows, synthetic injected-QNM
waveform), plus the fraction of the measured FP64 roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over the whole grid (65 536 fits), sharded by
flat grid index over the N ranks (strong scaling: the grid is fixed), followed at N > 1
by the NCCL all-gather of the mismatch slabs.

* value      device-resident throughput: tables, times and data already in HBM; per
             step one fit kernel (incl. the exchange); CUDA events on the launching stream;
             an L2 flush (256 MiB write) between steps, outside the events.
* e2e        the same grid through the public API ``mismatch_M_chi_grid`` with HOST
             numpy inputs: host tabulation, H2D, kernel, gather, D2H inside the region.
* roofline   FP64: algorithmic flops per fit F(M, N) x fits / kernel time,
             against the DFMA peak measured live on this GPU (MEASURED_PEAKS.json has
             no FP64 entry); HBM traffic is reported next to it.
* cpu_baseline / --impl reference   the UNMODIFIED reference's own mismatch_M_chi_grid
             (byte-compiled under oracle/_ref/ by oracle/make_ref.py so that it travels to
             the GPU box; oracle/qnmfits_oracle.py only if that is missing) timed on this
             box's host cores on a bounded, coarser sample of the same (Mf, chif) rectangle.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RES = 256
METRIC = "qnm_lstsq_fits_per_sec_256x256_Mf_chif_grid"
UNIT = "fits/s"
WORKLOAD = ("cfg3 mismatch_M_chi_grid: 256x256 Mf-chif grid, modes (2,2,0..7,+1) on synthetic "
            "h22, dt=0.1M, T=100M (M=1000 rows, N=8), synthetic Kerr tables")


# ------------------------------------------------------------------ CPU arms

_CPU_STATE = None
CPU_SAMPLE_FITS = 4096          # fits per step of the multi-process CPU arm (whole grid: 65 536)


def _cpu_init():
    """Per-process state of the CPU arm: workload + the implementation to time.

    ``kind`` "reference": the UNMODIFIED reference's own ``mismatch_M_chi_grid`` (its public
    API, stock code path), imported by oracle/ref_loader.py from /root/reference or — on the
    GPU box — from the byte-compiled copy oracle/make_ref.py built under oracle/_ref/.
    ``kind`` "port": oracle/qnmfits_oracle.py, the numpy restatement, when neither exists."""
    global _CPU_STATE
    if _CPU_STATE is None:
        os.environ.setdefault("TQDM_DISABLE", "1")   # the reference wraps its loop in a progress bar
        try:
            from threadpoolctl import threadpool_limits
            limiter = threadpool_limits(1)          # one BLAS thread per process (faster at 1000 x 8)
        except Exception:  # pragma: no cover
            limiter = None
        import warnings
        from oracle import ref_loader
        from qnmfits_b200 import synthetic, workloads
        workloads.use_synthetic_tables()
        wl = workloads.config3(res=RES)
        if ref_loader.reference_available():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref = ref_loader.load_reference(synthetic.modes_cache)
            kind, grid = "reference", ref.mismatch_M_chi_grid
        else:
            from oracle import qnmfits_oracle as orc
            tables = orc.OracleTables(synthetic.modes_cache)
            kind = "port"

            def grid(*a, **k):
                return orc.mismatch_M_chi_grid(tables, *a, **k)
        grid(wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0, T=wl.T, res=2)   # warm the spline caches
        _CPU_STATE = (grid, wl, kind, limiter)
    return _CPU_STATE


def _cpu_kind():
    return _cpu_init()[2]


def _cpu_chunk(job):
    """One call of the CPU implementation's mismatch_M_chi_grid on a sub-rectangle of the
    benchmark grid: job = (Mf_lo, Mf_hi, res) -> res x res fits over the full spin range."""
    grid, wl, _, _ = _cpu_init()
    lo, hi, res = job
    t = time.perf_counter()
    out = grid(wl.times, wl.data, wl.modes, (lo, hi), wl.chif_minmax, wl.t0, T=wl.T, res=res)
    return time.perf_counter() - t, float(np.sum(out))


def cpu_jobs(workers, fits=CPU_SAMPLE_FITS):
    """Split the benchmark's Mf range into ``workers`` bands; each band is one call of
    mismatch_M_chi_grid with res = r, r*r ~ fits / workers."""
    from qnmfits_b200 import workloads
    lo, hi = workloads.Workload.Mf_minmax
    r = max(4, int(round((fits / workers) ** 0.5)))
    edges = np.linspace(lo, hi, workers + 1)
    return [(float(edges[k]), float(edges[k + 1]), r) for k in range(workers)], workers * r * r


def cpu_baseline_single(res=128):
    """One process, one BLAS thread (the faster 'as shipped' setting): the CPU
    implementation's mismatch_M_chi_grid on the benchmark's (Mf, chif) rectangle at res x res."""
    from qnmfits_b200 import workloads
    dt, _ = _cpu_chunk(workloads.Workload.Mf_minmax + (res,))
    n = res * res
    what = ("the unmodified reference's mismatch_M_chi_grid (oracle/_ref or /root/reference)"
            if _cpu_kind() == "reference" else "oracle/qnmfits_oracle.py (numpy lstsq per fit)")
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": _cpu_kind(),
            "sample": f"{n} fits: the same (Mf, chif) rectangle at res = {res} instead of 256, "
                      f"{dt:.1f} s, {what}, 1 BLAS thread"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation on all host cores (process pool
    over bands of the grid, one mismatch_M_chi_grid call per band and step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    jobs, n = cpu_jobs(cores)
    ctx = mp.get_context("spawn")
    times = []
    with ctx.Pool(len(jobs), initializer=_cpu_init) as pool:
        kind = pool.apply(_cpu_kind)
        for step in range(args.warmup + args.steps):
            t = time.perf_counter()
            pool.map(_cpu_chunk, jobs, chunksize=1)
            dt = time.perf_counter() - t
            if step >= args.warmup:
                times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = n / (ms * 1e-3)
    what = ("the unmodified reference (oracle/ref_loader.py)" if kind == "reference"
            else "oracle/qnmfits_oracle.py, the numpy restatement (reference not available)")
    sample = (f"{n} fits per step: {len(jobs)} bands of the Mf range x the full chif range, each one "
              f"mismatch_M_chi_grid(res={jobs[0][2]}) call of {what}; {len(jobs)} worker processes, "
              "1 BLAS thread each")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": len(jobs), "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# -------------------------------------------------------------- clock sampler

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU via NVML while running."""

    def __init__(self, index, period=0.004):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
            "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
            "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
            "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap,
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------ GPU arm

def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    import qnmfits_b200 as qf
    from qnmfits_b200 import _cabi, workloads
    from qnmfits_b200 import qnmfits as api
    from qnmfits_b200._engine import get_engine

    workloads.use_synthetic_tables()
    wl = workloads.config3(res=RES)
    eng = get_engine(local)
    n_fits = RES * RES
    grid_args = (wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0)
    grid_kw = dict(T=wl.T, res=RES)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: upload once, launch K times -------------------------
    sweep, shape = api._prepare_M_chi_grid(*grid_args, **grid_kw)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=eng.device)
    stream = torch.cuda.current_stream()
    for _ in range(max(args.warmup, 3)):
        sweep.launch()
    barrier()
    plan = eng.ctx.plan(sweep.batch)
    exchange = ("none (one rank)" if world == 1 else
                "fused into the fit kernel: peer stores over NVLink + epoch flags (qnmfit_fit_batch_peers)"
                if sweep.window is not None else "NCCL all_gather_into_tensor after the kernel")
    launches0 = eng.ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        barrier()
        for e0, e1, e2 in ev:
            flush.fill_(1.0)                       # evict L2 between steps (not timed)
            e0.record(stream)
            sweep.launch_kernel()                  # this rank's slab
            e1.record(stream)
            sweep.gather()                         # NCCL all-gather of the slabs (N > 1)
            e2.record(stream)
        barrier()
    launches = eng.ctx.launch_count() - launches0
    step_ms = [e0.elapsed_time(e2) for e0, e1, e2 in ev]
    kern_ms = [e0.elapsed_time(e1) for e0, e1, e2 in ev]
    ms_per_step = max_over_ranks(sum(step_ms) / len(step_ms))
    kernel_ms = max_over_ranks(sum(kern_ms) / len(kern_ms))
    value = n_fits / (ms_per_step * 1e-3)
    grid_dev = sweep.fetch()[0].reshape(shape)

    # ---- end-to-end arm: the public API with host inputs ---------------------------
    for _ in range(3):
        grid_e2e = qf.mismatch_M_chi_grid(*grid_args, **grid_kw)
    e2e_steps = max(3, min(args.steps, 20))
    barrier()
    h2d0, d2h0 = eng.h2d_bytes, eng.d2h_bytes
    t = time.perf_counter()
    for _ in range(e2e_steps):
        grid_e2e = qf.mismatch_M_chi_grid(*grid_args, **grid_kw)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t) / e2e_steps)
    h2d = (eng.h2d_bytes - h2d0) // e2e_steps
    d2h = (eng.d2h_bytes - d2h0) // e2e_steps
    assert np.array_equal(grid_e2e, grid_dev), "e2e and device-resident grids differ"

    # ---- parity of the N-GPU grid (outside every timed region): sampled points against the
    # oracle (the CPU restatement of the reference, as the checker), and the whole grid against
    # the golden checksum of the 1-GPU grid when this is a multi-GPU run ------------------
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import qnmfits_oracle as orc
        from qnmfits_b200 import synthetic
        tables = orc.OracleTables(synthetic.modes_cache)
        idx = np.sort(np.random.default_rng(2024).choice(n_fits, args.parity_points, replace=False))
        # spread over the slabs of every rank: at least one point per slab
        per = -(-n_fits // world)
        idx = np.unique(np.concatenate([idx, np.arange(world) * per, np.minimum(np.arange(1, world + 1) * per, n_fits) - 1]))
        want = orc.mismatch_M_chi_grid(tables, wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax,
                                       wl.t0, T=wl.T, res=RES, flat_indices=idx)
        got = grid_e2e.reshape(-1)[idx]
        import hashlib
        parity = {"max_abs": float(np.max(np.abs(got - want))), "n": int(len(idx)), "tol": 1e-10,
                  "against": "oracle/qnmfits_oracle.py (numpy lstsq per fit) on sampled grid points, "
                             "at least one per rank's slab",
                  "grid_sha256": hashlib.sha256(np.ascontiguousarray(grid_e2e).tobytes()).hexdigest(),
                  "grid_sha256_note": "identical for every --gpus N: sharding does not change a bit"}
        assert parity["max_abs"] < 1e-10, f"N-GPU grid differs from the oracle by {parity['max_abs']}"

    # ---- roofline ------------------------------------------------------------------
    rows = sweep.rows_max
    flops_fit = _cabi.flops_per_fit(rows, len(wl.modes), 1, bool(plan.fast_mismatch))
    flops_fit_8d = _cabi.flops_per_fit(rows, len(wl.modes))
    fits_per_launch = sweep.hi - sweep.lo
    achieved = fits_per_launch * flops_fit / (kernel_ms * 1e-3) * 1e-12
    # FP64 roofline denominator: the best DFMA rate this GPU sustains in a dependent-free
    # loop (kind 3: two register reads + one reused operand per FMA; kind 0 and kind 2
    # are the same loop with other operand patterns), next to the DMMA tensor rate.
    peaks = {"dfma_reuse2": eng.ctx.fp64_peak(0, 4096), "dfma_2reads": eng.ctx.fp64_peak(3, 4096),
             "dfma_3reads": eng.ctx.fp64_peak(2, 4096)}
    peak_dfma = max(peaks.values())
    peak_dmma = eng.ctx.fp64_peak(1, 4096)
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_file)).get("hbm_gbs") if os.path.isfile(peaks_file) else 6650.0
    # algorithmic HBM bytes per launch: window (times + data) + tables + 8 B/fit mismatch
    alg_bytes = rows * 24 + RES * len(wl.modes) * 16 + RES * 8 + fits_per_launch * (8 + 4)
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel per launch: NOT measured in
    # this run (a number taken under a profiler is not a bench value) — read from the committed
    # summary of the round's `ncu --set full` capture of the same command, and labelled so
    traffic = traffic_source = None
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.isfile(prof):
        summary = json.load(open(prof))
        traffic = summary.get("dram_bytes_per_launch")
        traffic_source = "profiles/ncu_summary.json: " + str(summary.get("source", "ncu --set full capture"))
    roofline = {
        "bound": "fp64", "achieved": achieved, "peak": peak_dfma, "unit": "TFLOP/s",
        "frac": achieved / peak_dfma, "traffic": traffic, "traffic_source": traffic_source,
        "peak_source": "measured live: best of three dependent-free DFMA loops on all SMs "
                       "(qnmfit_fp64_peak); MEASURED_PEAKS.json has no FP64 entry; nominal "
                       "148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s",
        "peak_variants": peaks,
        "dmma_peak": peak_dmma, "flops_per_fit": flops_fit, "fits_per_launch": fits_per_launch,
        "flops_note": ("flops_per_fit counts what the launched algorithm needs (fast_mismatch: no "
                       "model pass); frac_survey_8d uses SURVEY.md 8d's F(M,N) incl. model + "
                       "trapezoid sums, which this path does not execute"),
        "frac_survey_8d": fits_per_launch * flops_fit_8d / (kernel_ms * 1e-3) * 1e-12 / peak_dfma,
        "kernel_ms": kernel_ms,
        "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                "achieved_gbs": alg_bytes / (kernel_ms * 1e-3) * 1e-9, "peak_gbs": hbm_peak,
                "frac": alg_bytes / (kernel_ms * 1e-3) * 1e-9 / hbm_peak},
    }

    if rank == 0:
        cpu = cpu_baseline_single(res=128 if args.cpu_sample == "large" else 64) \
            if world == 1 and not args.no_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "fits_per_step": n_fits, "rows": rows,
                       "modes": len(wl.modes), "sharding": f"flat grid index over {world} rank(s)",
                       "exchange": exchange,
                       "l2": "256 MiB fill between timed steps (outside the CUDA events)",
                       "kernel": {"id": plan.kernel, "lanes_per_fit": plan.lanes_per_fit,
                                  "grid": plan.grid, "block": plan.block,
                                  "smem_bytes": plan.smem_bytes, "regs": plan.regs_per_thread,
                                  "staged": bool(plan.staged),
                                  "fast_mismatch": bool(plan.fast_mismatch)}},
            "e2e": {"value": n_fits / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                    "api": "qnmfits_b200.mismatch_M_chi_grid(host numpy arrays)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "clocks": clocks.summary(),
        }
        if parity is not None:
            line["parity"] = parity
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-sample", default="large", choices=["small", "large"])
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the grid")
    ap.add_argument("--parity-points", type=int, default=512)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())

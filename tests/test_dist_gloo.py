"""N>1 path on CPU: two gloo ranks shard a grid by flat index (first_fit), compute their
slabs with the lane-emulation harness and all-gather; the result must be bit-identical
to the one-rank grid."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hostsim_driver as hs
    import qnmfits_b200 as qf
    from qnmfits_b200 import _dist, workloads
    from qnmfits_b200 import qnmfits as api
    workloads.use_synthetic_tables()
    res = 7                                    # 49 fits: not divisible by 2 -> padded slab
    wl = workloads.config3(res=res)
    Mf = np.linspace(*wl.Mf_minmax, res)
    chi = np.linspace(*wl.chif_minmax, res)
    table, ptr = qf.qnm.constituent_table(wl.modes, chi)
    win = api._window_rows(wl.times, 0.0, 100, "geq")
    n = res * res
    r, ws = _dist.world()
    assert (r, ws) == (rank, world)
    lo, hi, per = _dist.shard_bounds(n, r, ws)
    out = hs.run(wl.times, wl.data, n_fits=hi - lo, first_fit=lo, n_modes=8, window=win, t0=0.0,
                 lpf=4, table=table, mode_ptr=ptr, inv_Mf=1.0 / Mf, n_chi=res, n_mf=res)
    slab = torch.full((per,), float("nan"), dtype=torch.float64)
    slab[:hi - lo] = torch.from_numpy(out["mismatch"])
    full = _dist.all_gather_slabs(slab, n).numpy()
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_grid_is_bit_identical(tmp_path, qf):
    import torch.multiprocessing as mp
    import hostsim_driver as hs
    from qnmfits_b200 import workloads
    from qnmfits_b200 import qnmfits as api
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = 7
    wl = workloads.config3(res=res)
    Mf = np.linspace(*wl.Mf_minmax, res)
    chi = np.linspace(*wl.chif_minmax, res)
    table, ptr = qf.qnm.constituent_table(wl.modes, chi)
    win = api._window_rows(wl.times, 0.0, 100, "geq")
    one = hs.run(wl.times, wl.data, n_fits=res * res, n_modes=8, window=win, t0=0.0, lpf=4,
                 table=table, mode_ptr=ptr, inv_Mf=1.0 / Mf, n_chi=res, n_mf=res)["mismatch"]
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npy"))
        assert got.shape == (res * res,)
        assert np.array_equal(got, one)

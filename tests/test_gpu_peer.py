"""Two NCCL ranks on two GPUs: sweeps sharded by flat fit index with the exchange fused into
the fit kernels (peer stores + epoch flags, include/qnmfit.h ``qnmfit_fit_batch_peers``)
must give, on EVERY rank, the bit-identical result of the NCCL all-gather path AND of the
unsharded single-GPU sweep (the lanes-per-fit split, i.e. the summation tree of a fit, is
chosen from the whole sweep — qnmfit_batch.plan_fits — never from the slab).  Needs two
GPUs; skipped on a one-GPU box (tests/test_gpu_parity.py::test_slab_launches_are_bit_identical_
to_one_launch checks the same property there by launching the slabs one after another)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _sweeps(qf, wl3, wl2, wl4):
    """The sweeps under test (K1 shared window, K1 per-fit windows, K3 multimode, one
    single-fit sweep that leaves rank 1 without work), repeated to cycle epochs/slots."""
    out = {}
    for rep in range(3):
        out[f"grid{rep}"] = qf.mismatch_M_chi_grid(wl3.times, wl3.data, wl3.modes, wl3.Mf_minmax,
                                                   wl3.chif_minmax, wl3.t0, T=wl3.T, res=37 + rep)
    out["t0"] = np.array(qf.mismatch_t0_array(wl2.times, wl2.data, wl2.modes, wl2.Mf, wl2.chif,
                                              wl2.t0_array[:301]))
    out["one"] = np.array(qf.mismatch_t0_array(wl2.times, wl2.data, wl2.modes, wl2.Mf, wl2.chif,
                                               wl2.t0_array[:1]))
    out["multimode"] = np.array(qf.mismatch_t0_array(wl4.times, wl4.data, wl4.modes, wl4.Mf, wl4.chif,
                                                     wl4.t0_array))
    out["grid_again"] = qf.mismatch_M_chi_grid(wl3.times, wl3.data, wl3.modes, wl3.Mf_minmax,
                                               wl3.chif_minmax, wl3.t0, T=wl3.T, res=21)
    # numerically rank-deficient fits (duplicated label): every rank repairs its slab to numpy's
    # minimum-norm solution and the repaired slabs are exchanged once more
    out["deficient"] = np.array(qf.mismatch_t0_array(wl2.times, wl2.data, [(2, 2, n, 1) for n in (0, 1, 9, 10)],
                                                     wl2.Mf, wl2.chif, wl2.t0_array[:45]))
    # more results than the first peer window holds (2^20): the window is re-created collectively
    out["big"] = qf.mismatch_M_chi_grid(wl3.times, wl3.data, wl3.modes[:2], wl3.Mf_minmax,
                                        wl3.chif_minmax, wl3.t0, T=30, res=1030)
    out["after_big"] = qf.mismatch_M_chi_grid(wl3.times, wl3.data, wl3.modes, wl3.Mf_minmax,
                                              wl3.chif_minmax, wl3.t0, T=wl3.T, res=9)
    return out


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import cases
    import qnmfits_b200 as qf
    from qnmfits_b200 import _dist, workloads
    workloads.use_synthetic_tables()
    wl3, wl2, wl4 = workloads.config3(res=8), workloads.config2(), cases.cfg4_small()

    fused = _sweeps(qf, wl3, wl2, wl4)
    assert _dist._windows, "the fused exchange was not used"
    win = next(iter(_dist._windows.values()))
    epochs, capacity = win.epoch, win.capacity
    os.environ["QNMFITS_B200_PEER"] = "0"
    nccl = _sweeps(qf, wl3, wl2, wl4)
    os.environ["QNMFITS_B200_NO_SHARD"] = "1"
    single = _sweeps(qf, wl3, wl2, wl4)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), epochs=epochs, capacity=capacity,
             **{f"fused_{k}": v for k, v in fused.items()},
             **{f"nccl_{k}": v for k, v in nccl.items()},
             **{f"single_{k}": v for k, v in single.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_fused_exchange_two_gpus_bit_identical(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        assert int(got["capacity"]) == 1 << 21  # grown once, for the 1030 x 1030 grid
        assert int(got["epochs"]) == 2          # ... and the new window counts its own exchanges
        names = [k[len("single_"):] for k in got.files if k.startswith("single_")]
        assert len(names) == 10
        for name in names:
            one = got["single_" + name]
            assert np.all(np.isfinite(one))
            assert np.array_equal(got["fused_" + name], got["nccl_" + name]), (rank, name)
            assert np.array_equal(got["fused_" + name], one), (rank, name)   # sharding does not change a bit


def test_single_process_device_group_matches_one_device():
    """``use_devices``: one process drives two GPUs, the host concatenates the slabs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    import qnmfits_b200 as qf
    from qnmfits_b200 import workloads
    workloads.use_synthetic_tables()
    wl3, wl2, wl4 = workloads.config3(res=8), workloads.config2(), cases.cfg4_small()
    torch.cuda.set_device(0)
    one = _sweeps(qf, wl3, wl2, wl4)
    grid8_args = (wl3.modes, wl3.Mf_minmax, wl3.chif_minmax, wl3.t0)
    grid8 = qf.mismatch_M_chi_grid(wl3.times, wl3.data, *grid8_args, T=wl3.T, res=8)
    try:
        assert qf.use_devices("all")[:2] == [0, 1]
        qf.use_devices([1, 0])
        two = _sweeps(qf, wl3, wl2, wl4)
        assert torch.cuda.current_device() == 0
        # repeated calls find the prepared sweeps of the group and go through its one-call-per-
        # device path (staging shared by the devices, no wait between the devices' enqueues)
        launches = [qf.qnmfits.get_engine(dev).ctx.launch_count() for dev in (0, 1)]
        two_again = _sweeps(qf, wl3, wl2, wl4)
        other = wl3.data * (1.0 + 0.1j) + 1e-3 * np.exp(0.3j * wl3.times)
        grid8_two = qf.mismatch_M_chi_grid(wl3.times, wl3.data, *grid8_args, T=wl3.T, res=8)
        grid_other = qf.mismatch_M_chi_grid(wl3.times, other, *grid8_args, T=wl3.T, res=8)     # prepared: rerun
        assert all(qf.qnmfits.get_engine(dev).ctx.launch_count() > n for dev, n in zip((0, 1), launches))
    finally:
        qf.use_devices(None)
    again = _sweeps(qf, wl3, wl2, wl4)
    grid_other_one = qf.mismatch_M_chi_grid(wl3.times, other, *grid8_args, T=wl3.T, res=8)
    assert np.array_equal(grid8_two, grid8) and np.array_equal(grid_other, grid_other_one)
    assert not np.array_equal(grid_other_one, grid8)
    for name, ref in one.items():
        assert two[name].shape == ref.shape and np.all(np.isfinite(two[name]))
        assert np.array_equal(two[name], ref), name
        assert np.array_equal(two_again[name], ref), name
        assert np.array_equal(again[name], ref), name
    with pytest.raises(ValueError):
        qf.use_devices([0, 0])


def _late_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_WORLD_SIZE=str(world),
                      QNMFITS_B200_PEER_TIMEOUT_S="2")
    import time
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import qnmfits_b200 as qf
    from qnmfits_b200 import workloads
    workloads.use_synthetic_tables()
    wl = workloads.config3(res=8)
    args = (wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0)
    first = qf.mismatch_M_chi_grid(*args, T=wl.T, res=12)
    outcome = "ok"
    if rank == 1:
        time.sleep(6.0)                       # rank 0 gives up on this rank's slab after 2 s
    t = time.time()
    try:
        second = qf.mismatch_M_chi_grid(*args, T=wl.T, res=12)
        assert np.array_equal(second, first)
    except RuntimeError as exc:
        outcome = f"raised after {time.time() - t:.1f} s: {exc}"
    dist.barrier()                            # the late rank catches up
    third = qf.mismatch_M_chi_grid(*args, T=wl.T, res=12)     # both ranks again in step
    assert np.array_equal(third, first)
    with open(os.path.join(out_dir, f"late{rank}.txt"), "w") as f:
        f.write(outcome)
    dist.barrier()
    dist.destroy_process_group()


def test_missing_peer_times_out_instead_of_hanging(tmp_path):
    """A rank that is late by more than QNMFITS_B200_PEER_TIMEOUT_S makes the waiting rank
    raise (its kernel leaves a NaN marker) — never a hang; the next sweep is in step again."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_late_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    late0 = open(os.path.join(str(tmp_path), "late0.txt")).read()
    late1 = open(os.path.join(str(tmp_path), "late1.txt")).read()
    assert late0.startswith("raised after") and "did not deliver" in late0, late0
    assert late1 == "ok", late1

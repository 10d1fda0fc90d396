"""The C-ABI library loads on a CPU-only box and exports every symbol the header declares."""
import ctypes as C
import os
import re

from qnmfits_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "qnmfit.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qnmfit_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    names = _declared()
    assert set(names) == set(_cabi.EXPORTS)
    lib = _cabi.load_library()
    for n in names:
        assert hasattr(lib, n), n


def test_abi_version_and_struct_size():
    lib = _cabi.load_library()
    assert lib.qnmfit_abi_version() == _cabi.ABI_VERSION
    hs = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libqnmfit_hostsim.so"))
    assert hs.hostsim_sizeof_batch() == C.sizeof(_cabi.Batch)
    assert hs.hostsim_sizeof_peers() == C.sizeof(_cabi.Peers)


def test_create_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        return
    lib = _cabi.load_library()
    h = C.c_void_p()
    rc = lib.qnmfit_create(0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.qnmfit_last_error(None)


def test_flops_entry_point_matches_python():
    lib = _cabi.load_library()
    for rows, n, l in ((1000, 8, 1), (1000, 40, 21), (37, 3, 1)):
        for fast in (0, 1):
            assert abs(lib.qnmfit_flops_per_fit(rows, n, l, fast)
                       - _cabi.flops_per_fit(rows, n, l, bool(fast))) < 1e-6

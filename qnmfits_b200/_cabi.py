"""
ctypes binding of ``libqnmfit.so`` (C ABI declared in ``include/qnmfit.h``).

There is deliberately no fallback: if the shared library is missing or has no
usable sm_100 device the functions here raise, they never compute on the CPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqnmfit.so")

ABI_VERSION = 7
MAX_MODES_SMALL = 12
MIN_MODES_PAIR, MAX_MODES_PAIR = 9, 24
MAX_MODES = 64
MAX_PEERS = 8

KERNEL_AUTO, KERNEL_SMALL, KERNEL_GENERAL, KERNEL_STRUCT, KERNEL_PANEL, KERNEL_PAIR = 0, 1, 2, 3, 4, 5

ST_RANK_DEFICIENT, ST_NONFINITE, ST_UNDERDETERMINED = 1, 2, 4

_dp = C.c_void_p  # device (or, for the test harness, host) pointer


class Batch(C.Structure):
    """Mirror of ``struct qnmfit_batch`` — field order and types must match the header."""
    _fields_ = [
        ("struct_size", C.c_int32), ("kernel", C.c_int32),
        ("n_fits", C.c_int32), ("n_modes", C.c_int32),
        ("n_series", C.c_int32), ("n_times", C.c_int32),
        ("series_stride", C.c_int64), ("first_fit", C.c_int64),
        ("times", _dp), ("data", _dp),
        ("row_begin", _dp), ("row_end", _dp), ("t0", _dp),
        ("row_begin_all", C.c_int32), ("row_end_all", C.c_int32),
        ("t0_all", C.c_double),
        ("omega", _dp), ("omega_tilde", _dp), ("mode_ptr", _dp),
        ("inv_Mf", _dp), ("delta_factor", _dp),
        ("chi_index", _dp), ("mf_index", _dp),
        ("n_chi", C.c_int32), ("n_mf", C.c_int32),
        ("n_constituents", C.c_int32), ("omega_shared", C.c_int32),
        ("coef", _dp), ("coef_index", _dp),
        ("n_coef", C.c_int32), ("anchor_rows", C.c_int32),
        ("dt_nominal", C.c_double),
        ("C", _dp),
        ("mismatch", _dp), ("residual", _dp), ("R", _dp), ("status", _dp),
        ("model", _dp), ("model_stride", C.c_int64),
        ("uniform_weights", C.c_int32), ("plan_fits", C.c_int32),
        ("flagged_count", _dp),
        ("series_index", _dp),
        ("omega_rows", _dp), ("coef_rows", _dp),
        ("flag_list", _dp), ("flag_capacity", C.c_int32), ("reserved2", C.c_int32),
        ("fit_index", _dp),
    ]

    def __init__(self, **kw):
        super().__init__(**kw)
        self.struct_size = C.sizeof(Batch)


class Peers(C.Structure):
    """Mirror of ``struct qnmfit_peers`` (multi-GPU exchange fused into the fit kernels)."""
    _fields_ = [
        ("struct_size", C.c_int32), ("n_peers", C.c_int32),
        ("rank", C.c_int32), ("reserved", C.c_int32),
        ("epoch", C.c_int64), ("timeout_ns", C.c_int64),
        ("mismatch", _dp * MAX_PEERS), ("flagged", _dp * MAX_PEERS), ("flags", _dp * MAX_PEERS),
    ]

    def __init__(self, **kw):
        super().__init__(**kw)
        self.struct_size = C.sizeof(Peers)


class Copy(C.Structure):
    """Mirror of ``struct qnmfit_copy`` (one host -> device upload of ``qnmfit_run_host``)."""
    _fields_ = [("dst_dev", _dp), ("src_host", _dp), ("bytes", C.c_size_t)]


RUN_COALESCE, RUN_ZERO_COUNTER, RUN_RESULT_PINNED, RUN_NO_SYNC, RUN_UPLOADS_PINNED = 1, 2, 4, 8, 16


class Plan(C.Structure):
    _fields_ = [
        ("kernel", C.c_int32), ("lanes_per_fit", C.c_int32),
        ("grid", C.c_int32), ("block", C.c_int32),
        ("smem_bytes", C.c_int32), ("regs_per_thread", C.c_int32),
        ("staged", C.c_int32), ("fast_mismatch", C.c_int32),
    ]


#: every symbol include/qnmfit.h declares (checked by tests/test_cabi_symbols.py)
EXPORTS = (
    "qnmfit_create", "qnmfit_destroy", "qnmfit_last_error", "qnmfit_fit_batch",
    "qnmfit_eval_batch", "qnmfit_launch_count", "qnmfit_plan_batch",
    "qnmfit_fp64_peak", "qnmfit_flops_per_fit", "qnmfit_abi_version",
    "qnmfit_peer_alloc", "qnmfit_peer_open", "qnmfit_peer_close", "qnmfit_peer_free",
    "qnmfit_fit_batch_peers",
    "qnmfit_h2d", "qnmfit_h2d_wait", "qnmfit_d2h", "qnmfit_zero", "qnmfit_stream_sync",
    "qnmfit_run_host",
    "qnmfit_nm_create", "qnmfit_nm_step", "qnmfit_nm_result", "qnmfit_nm_destroy",
)

NM_MAX_VARS = 64                     # QNMFIT_NM_MAX_VARS
NM_ORDER_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int64, C.c_int, C.POINTER(C.c_int64), C.c_void_p)

_lib = None


class QnmfitError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libqnmfit error {code}: {message}")
        self.code = code


def load_library(path=None):
    """dlopen the CUDA library and declare prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("QNMFIT_LIB") or LIB_PATH   # QNMFIT_LIB: developer builds (tools/)
    if not os.path.isfile(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). qnmfits_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    lib.qnmfit_abi_version.restype = C.c_int
    lib.qnmfit_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.qnmfit_create.restype = C.c_int
    lib.qnmfit_destroy.argtypes = [C.c_void_p]
    lib.qnmfit_destroy.restype = C.c_int
    lib.qnmfit_last_error.argtypes = [C.c_void_p]
    lib.qnmfit_last_error.restype = C.c_char_p
    for name in ("qnmfit_fit_batch", "qnmfit_eval_batch"):
        fn = getattr(lib, name)
        fn.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_void_p]
        fn.restype = C.c_int
    lib.qnmfit_fit_batch_peers.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(Peers), C.c_void_p]
    lib.qnmfit_fit_batch_peers.restype = C.c_int
    lib.qnmfit_peer_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
    lib.qnmfit_peer_alloc.restype = C.c_int
    lib.qnmfit_peer_open.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.qnmfit_peer_open.restype = C.c_int
    for name in ("qnmfit_peer_close", "qnmfit_peer_free"):
        fn = getattr(lib, name)
        fn.argtypes = [C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
    lib.qnmfit_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.qnmfit_h2d.restype = C.c_int
    lib.qnmfit_h2d_wait.argtypes = [C.c_void_p]
    lib.qnmfit_h2d_wait.restype = C.c_int
    lib.qnmfit_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    lib.qnmfit_d2h.restype = C.c_int
    lib.qnmfit_zero.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.qnmfit_zero.restype = C.c_int
    lib.qnmfit_stream_sync.argtypes = [C.c_void_p, C.c_void_p]
    lib.qnmfit_stream_sync.restype = C.c_int
    lib.qnmfit_run_host.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(Peers), C.POINTER(Copy), C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    lib.qnmfit_run_host.restype = C.c_int
    lib.qnmfit_launch_count.argtypes = [C.c_void_p]
    lib.qnmfit_launch_count.restype = C.c_int64
    lib.qnmfit_plan_batch.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(Plan)]
    lib.qnmfit_plan_batch.restype = C.c_int
    lib.qnmfit_fp64_peak.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    lib.qnmfit_fp64_peak.restype = C.c_int
    lib.qnmfit_flops_per_fit.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    lib.qnmfit_flops_per_fit.restype = C.c_double
    lib.qnmfit_nm_create.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                     C.c_double, C.c_double, C.c_double, NM_ORDER_FN, C.c_void_p,
                                     C.POINTER(C.c_void_p)]
    lib.qnmfit_nm_create.restype = C.c_int
    lib.qnmfit_nm_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qnmfit_nm_step.restype = C.c_int64
    lib.qnmfit_nm_result.argtypes = [C.c_void_p] * 7
    lib.qnmfit_nm_result.restype = C.c_int
    lib.qnmfit_nm_destroy.argtypes = [C.c_void_p]
    lib.qnmfit_nm_destroy.restype = C.c_int
    if lib.qnmfit_abi_version() != ABI_VERSION:
        raise ImportError(
            f"{path}: ABI version {lib.qnmfit_abi_version()} != binding {ABI_VERSION}; rebuild")
    if path == LIB_PATH:
        _lib = lib
    return lib


class Context:
    """Owns one ``qnmfit_ctx`` (one CUDA device).  Not thread-safe, like the reference."""

    def __init__(self, device=0):
        self.lib = load_library()
        handle = C.c_void_p()
        rc = self.lib.qnmfit_create(int(device), C.byref(handle))
        if rc != 0:
            raise QnmfitError(rc, self.lib.qnmfit_last_error(None).decode())
        self.handle = handle
        self.device = int(device)

    def _check(self, rc):
        if rc != 0:
            raise QnmfitError(rc, self.lib.qnmfit_last_error(self.handle).decode())

    def fit_batch(self, batch, stream=0):
        self._check(self.lib.qnmfit_fit_batch(self.handle, C.byref(batch), C.c_void_p(stream)))

    def eval_batch(self, batch, stream=0):
        self._check(self.lib.qnmfit_eval_batch(self.handle, C.byref(batch), C.c_void_p(stream)))

    def fit_batch_peers(self, batch, peers, stream=0):
        self._check(self.lib.qnmfit_fit_batch_peers(self.handle, C.byref(batch), C.byref(peers),
                                                    C.c_void_p(stream)))

    def peer_alloc(self, nbytes):
        """(device pointer, 64-byte handle) of a zero-filled allocation other ranks can map."""
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._check(self.lib.qnmfit_peer_alloc(self.handle, int(nbytes), C.byref(ptr), handle))
        return int(ptr.value), handle.raw

    def peer_open(self, handle):
        ptr = C.c_void_p()
        self._check(self.lib.qnmfit_peer_open(self.handle, bytes(handle), C.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr):
        self._check(self.lib.qnmfit_peer_close(self.handle, C.c_void_p(ptr)))

    def peer_free(self, ptr):
        self._check(self.lib.qnmfit_peer_free(self.handle, C.c_void_p(ptr)))

    def h2d(self, dst, src, nbytes, stream=0):
        self._check(self.lib.qnmfit_h2d(self.handle, dst, src, nbytes, stream))

    def h2d_wait(self):
        self._check(self.lib.qnmfit_h2d_wait(self.handle))

    def d2h(self, dst, src, nbytes, stream=0, sync=True):
        self._check(self.lib.qnmfit_d2h(self.handle, dst, src, nbytes, stream, 1 if sync else 0))

    def zero(self, dst, nbytes, stream=0):
        self._check(self.lib.qnmfit_zero(self.handle, dst, nbytes, stream))

    def stream_sync(self, stream=0):
        self._check(self.lib.qnmfit_stream_sync(self.handle, stream))

    def run_host(self, batch, peers, uploads, n_uploads, result_dev, result_host, result_bytes, flags, stream=0):
        """``qnmfit_run_host``: uploads, launch, download, synchronise in one C call.
        ``uploads`` is a ctypes array of ``Copy`` (or None), ``peers`` a ``Peers`` or None."""
        self._check(self.lib.qnmfit_run_host(
            self.handle, C.byref(batch), None if peers is None else C.byref(peers), uploads, n_uploads,
            result_dev, result_host, result_bytes, flags, stream))

    def plan(self, batch):
        plan = Plan()
        self._check(self.lib.qnmfit_plan_batch(self.handle, C.byref(batch), C.byref(plan)))
        return plan

    def launch_count(self):
        return int(self.lib.qnmfit_launch_count(self.handle))

    def fp64_peak(self, kind=0, iters=4096):
        out = C.c_double()
        self._check(self.lib.qnmfit_fp64_peak(self.handle, int(kind), int(iters), C.byref(out)))
        return out.value

    def close(self):
        if getattr(self, "handle", None):
            self.lib.qnmfit_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def flops_per_fit(rows, n_modes, n_series=1, fast_mismatch=False):
    """Algorithmic FP64 flops credited to one fit (same formulas as the C library)."""
    M, N = float(rows) * n_series, float(n_modes)
    if fast_mismatch:
        return 8 * M * N * N + 22 * M * N + 4 * M - (8.0 / 3.0) * N ** 3 - 4 * N * N + 28 * N
    return 8 * M * N * N + 30 * M * N + 20 * M - (8.0 / 3.0) * N ** 3 - 4 * N * N

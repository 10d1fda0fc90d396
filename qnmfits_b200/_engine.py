"""
Device engine: owns the ``libqnmfit`` context of this process's GPU, moves host
arrays into torch CUDA tensors (used purely as device buffers) and fills the
``qnmfit_batch`` descriptor.  One process drives one GPU; with
``torch.distributed`` initialised (NCCL), sweeps are sharded by flat fit index
across ranks and the mismatch slabs are all-gathered (``_dist.py``).

No CPU fallback: constructing the engine without CUDA raises.
"""
import numpy as np

from . import _cabi

_engines = {}


def get_engine(device=None):
    import torch
    if device is None and _engines:
        eng = _engines.get(torch.cuda.current_device())
        if eng is not None:
            return eng
    if not torch.cuda.is_available():
        raise RuntimeError(
            "qnmfits_b200 needs an NVIDIA B200 (sm_100a) GPU: torch.cuda.is_available() "
            "is False and there is no CPU fallback.")
    if device is None:
        device = torch.cuda.current_device()
    device = int(device)
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]


def grid_plan(times, wmax, steps=None):
    """(dt_nominal, uniform_weights) of a window's time samples, from one pass over its steps.

    ``dt_nominal``: the nominal sample spacing if ``times`` is uniform enough for the
    recurrence path.  The kernels advance a row by ``z *= exp(-i w dt) * (1 - i w d_eps)``
    where ``d_eps`` is the deviation of the actual step from ``dt``; the dropped second
    order term is ``(|w| d_eps)^2 / 2`` per row, kept below 1e-19 here.  Grids like
    ``np.arange(n) * 0.1`` (deviations ~1e-14) qualify; genuinely non-uniform grids give
    0.0, which selects direct exp/sincos evaluation of every element.

    ``uniform_weights``: True when every step is within 1e-11 (relative) of ``dt``: the
    trapezoid weights are then uniform to 1e-11 and K1/K3 may take the mismatch from
    by-products of the factorisation instead of a weighted second pass (changes it by
    < 1e-11).  ``steps`` = ``np.diff(times)`` if the caller has it already.
    """
    return grid_decision(*step_stats(times, steps), wmax)


def step_stats(times, steps=None):
    """(dt, dev) of a window: mean step and largest deviation of a step from it, or
    (0.0, inf) when the window is too short or not increasing."""
    times = np.asarray(times, dtype=float)
    if times.size < 3:
        return 0.0, np.inf
    if steps is None:
        steps = np.diff(times)
    dt = float((times[-1] - times[0]) / (times.size - 1))
    if not np.isfinite(dt) or dt <= 0.0:
        return 0.0, np.inf
    return dt, float(np.max(np.abs(steps - dt)))


def grid_decision(dt, dev, wmax):
    """``grid_plan``'s two criteria from the window's step statistics."""
    if not dt > 0.0 or not dev * max(float(wmax), 1.0) <= 4e-10:
        return 0.0, False
    return dt, bool(dev <= 1e-11 * dt)


def nominal_step(times, wmax, steps=None):
    """``grid_plan(...)[0]``."""
    return grid_plan(times, wmax, steps)[0]


def uniform_weights(times, dt, steps=None):
    """``grid_plan``'s second criterion for a given ``dt``."""
    times = np.asarray(times, dtype=float)
    if not dt > 0.0 or times.size < 3:
        return False
    if steps is None:
        steps = np.diff(times)
    return bool(np.max(np.abs(steps - dt)) <= 1e-11 * dt)


def pack_layout(sizes, direct_bytes, zero_head=0):
    """Byte offsets of a packed upload (``Engine.upload_packed``): pure arithmetic, unit-tested
    on the CPU.  ``sizes``: byte counts (None = absent array).  Arrays of ``direct_bytes`` or
    more go first (each copied from its own memory); the others form ONE staged group behind
    them, 256-byte aligned each, closed by ``zero_head`` zero bytes; the result region starts
    right after the staged group.  Returns (offsets, stage_begin, stage_end, out_offset) with
    ``out_offset == stage_end`` and, when ``zero_head`` > 0, the zero bytes at
    ``[out_offset - zero_head, out_offset)``."""
    if zero_head % 16:
        raise ValueError("zero_head must be a multiple of 16")
    offsets, total = [None] * len(sizes), 0
    big = [i for i, n in enumerate(sizes) if n is not None and n >= direct_bytes]
    for i in big:
        total = (total + 255) // 256 * 256
        offsets[i] = total
        total += sizes[i]
    stage_begin = total = (total + 255) // 256 * 256
    for i, n in enumerate(sizes):
        if n is None or i in big:
            continue
        total = (total + 255) // 256 * 256
        offsets[i] = total
        total += n
    total = (total + 15) // 16 * 16 + zero_head
    return offsets, stage_begin, total, total


class Engine:
    def __init__(self, device):
        import torch
        self.torch = torch
        self.device_index = device
        self.device = torch.device("cuda", device)
        self.ctx = _cabi.Context(device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._pinned_bufs = {}
        self._np_dtype = {torch.float64: np.float64, torch.complex128: np.complex128,
                          torch.int32: np.int32, torch.int64: np.int64, torch.uint8: np.uint8}

    # ------------------------------------------------------------- transfers

    def to_device(self, array, dtype):
        torch = self.torch
        a = np.ascontiguousarray(array, dtype=dtype)
        self.h2d_bytes += a.nbytes
        if a.size == 0:
            return torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, device=self.device)
        if not a.flags.writeable:                   # memoised tables are read-only; torch wants writable
            a = a.copy()
        return torch.from_numpy(a).to(self.device)

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=dtype, device=self.device)

    def to_host(self, tensor):
        out = tensor.cpu().numpy()
        self.d2h_bytes += out.nbytes
        return out

    def stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    # ---------------------------------------- low-overhead packed transfers
    #
    # A sweep through the public API moves every input in ONE host->device copy and
    # every result in ONE device->host copy, issued through the thin C wrappers
    # (qnmfit_h2d / qnmfit_d2h): a torch copy_ costs ~10 us of dispatch each.

    DIRECT_BYTES = 1 << 22   # arrays this large are uploaded straight from their own memory

    def _pinned(self, name, nbytes):
        """A cached pinned host buffer of at least nbytes: (uint8 numpy view, address)."""
        buf = self._pinned_bufs.get(name)
        if buf is None or buf[0].numel() < nbytes:
            t = self.torch.empty(max(int(nbytes) * 5 // 4, 1 << 16), dtype=self.torch.uint8,
                                 pin_memory=True)
            buf = self._pinned_bufs[name] = (t, t.numpy(), t.data_ptr())
        return buf[1], buf[2]

    def upload_packed(self, arrays, out_bytes=0, stream=None, zero_head=0):
        """Copy several host arrays to the device with ONE cudaMemcpyAsync.

        ``arrays`` is a list of C-contiguous numpy arrays (or None).  They are packed,
        256-byte aligned, into a pinned staging buffer and copied into one freshly
        allocated device buffer, which also gets ``out_bytes`` of (uninitialised) room for
        results behind everything else, 16-byte aligned.  ``zero_head`` (a multiple of 16)
        bytes of zeros travel with the staged copy and sit IMMEDIATELY in front of the
        result region, whatever the sizes of the arrays (``pack_layout``): a sweep keeps its
        counter of flagged fits there, so that [counter | results] comes back in one copy.
        Returns (device_buffer, [device pointer or None, ...], pointer of the result region).
        The staging buffer is reused by the next call, which first waits until the
        previous copy has left it (``qnmfit_h2d_wait``).  Arrays of ``DIRECT_BYTES`` or
        more are not staged: they are copied from their own (pageable) memory into the
        front of the device buffer.
        """
        sizes = [None if a is None else a.nbytes for a in arrays]
        offsets, stage_begin, stage_end, out_off = pack_layout(sizes, self.DIRECT_BYTES, zero_head)
        stream = self.stream() if stream is None else stream
        dev = self.torch.empty(max(out_off + int(out_bytes), 16), dtype=self.torch.uint8, device=self.device)
        base = dev.data_ptr()
        if stage_end > stage_begin:
            self.ctx.h2d_wait()
            stage_np, stage_ptr = self._pinned("upload", stage_end - stage_begin)
            for a, off, size in zip(arrays, offsets, sizes):
                if a is not None and size and size < self.DIRECT_BYTES:
                    stage_np[off - stage_begin:off - stage_begin + size] = a.reshape(-1).view(np.uint8)
            if zero_head:
                stage_np[stage_end - stage_begin - zero_head:stage_end - stage_begin] = 0
            self.ctx.h2d(base + stage_begin, stage_ptr, stage_end - stage_begin, stream)
        for a, off, size in zip(arrays, offsets, sizes):   # pageable source: returns once the data has left it
            if a is not None and size is not None and size >= self.DIRECT_BYTES:
                self.ctx.h2d(base + off, a.ctypes.data, size, stream)
        self.h2d_bytes += sum(sz for sz in sizes if sz is not None)
        return dev, [None if off is None else base + off for off in offsets], base + out_off

    def download_raw(self, ptr, nbytes, dtype=np.float64, stream=None):
        """``nbytes`` at device address ``ptr`` -> new numpy array of ``dtype``: one async
        copy + one stream synchronisation.  Small results go through a cached pinned
        buffer and are copied out; large ones land in a pinned block of their own (torch's
        caching host allocator recycles it when the array dies), which saves the host copy."""
        stream = self.stream() if stream is None else stream
        self.d2h_bytes += nbytes
        if nbytes >= 1 << 16:
            block = self.torch.empty(nbytes, dtype=self.torch.uint8, pin_memory=True)
            self.ctx.d2h(block.data_ptr(), ptr, nbytes, stream, sync=True)
            return block.numpy().view(dtype)
        stage_np, stage_ptr = self._pinned("download", max(nbytes, 8))
        self.ctx.d2h(stage_ptr, ptr, nbytes, stream, sync=True)
        return stage_np[:nbytes].view(dtype).copy()

    def download(self, tensor):
        """Device tensor (contiguous) -> new numpy array, as ``download_raw``."""
        nbytes = tensor.numel() * tensor.element_size()
        return self.download_raw(tensor.data_ptr(), nbytes, self._np_dtype[tensor.dtype])

    # --------------------------------------------------------------- batches

    def make_batch(self, *, times_d, data_d, n_fits, n_modes, n_series=1, first_fit=0,
                   row_begin_all, row_end_all, t0_all=0.0,
                   row_begin_d=None, row_end_d=None, t0_d=None,
                   omega_d=None, omega_shared=False,
                   omega_tilde_d=None, mode_ptr_d=None, inv_Mf_d=None, delta_factor_d=None,
                   chi_index_d=None, mf_index_d=None, n_chi=0, n_mf=0, n_constituents=0,
                   coef_d=None, coef_index_d=None, n_coef=0,
                   dt_nominal=0.0, anchor_rows=0, kernel=_cabi.KERNEL_AUTO,
                   C_d=None, mismatch_d=None, residual_d=None, R_d=None, status_d=None,
                   model_d=None, model_stride=0, uniform_weights=False,
                   n_times=None, series_stride=None, flagged_d=None, series_index_d=None,
                   omega_rows_d=None, coef_rows_d=None, plan_fits=0,
                   flag_list_d=None, flag_capacity=0, fit_index_d=None):
        def p(t):                              # device pointer: int, None, or a torch tensor
            return t if t is None or type(t) is int else t.data_ptr()
        if n_times is None:
            n_times = int(times_d.numel())
        if series_stride is None:
            series_stride = int(data_d.shape[-1])
        b = _cabi.Batch(
            kernel=kernel, n_fits=int(n_fits), n_modes=int(n_modes), n_series=int(n_series),
            n_times=int(n_times), series_stride=int(series_stride), first_fit=int(first_fit),
            times=p(times_d), data=p(data_d),
            row_begin=p(row_begin_d), row_end=p(row_end_d), t0=p(t0_d),
            row_begin_all=int(row_begin_all), row_end_all=int(row_end_all), t0_all=float(t0_all),
            omega=p(omega_d), omega_shared=1 if omega_shared else 0,
            omega_tilde=p(omega_tilde_d), mode_ptr=p(mode_ptr_d), inv_Mf=p(inv_Mf_d),
            delta_factor=p(delta_factor_d), chi_index=p(chi_index_d), mf_index=p(mf_index_d),
            n_chi=int(n_chi), n_mf=int(n_mf), n_constituents=int(n_constituents),
            coef=p(coef_d), coef_index=p(coef_index_d), n_coef=int(n_coef),
            dt_nominal=float(dt_nominal), anchor_rows=int(anchor_rows),
            C=p(C_d), mismatch=p(mismatch_d), residual=p(residual_d), R=p(R_d),
            status=p(status_d), model=p(model_d), model_stride=int(model_stride),
            uniform_weights=1 if uniform_weights else 0, flagged_count=p(flagged_d),
            series_index=p(series_index_d), omega_rows=p(omega_rows_d), coef_rows=p(coef_rows_d),
            plan_fits=int(plan_fits), flag_list=p(flag_list_d), flag_capacity=int(flag_capacity),
            fit_index=p(fit_index_d))
        return b

    def fit(self, batch):
        self.ctx.fit_batch(batch, self.stream())

    def fit_peers(self, batch, peers):
        """Fit + exchange fused in the kernel (``_dist.PeerWindow``)."""
        self.ctx.fit_batch_peers(batch, peers, self.stream())

    def evaluate(self, batch):
        self.ctx.eval_batch(batch, self.stream())

    def synchronize(self):
        self.ctx.stream_sync(self.stream())

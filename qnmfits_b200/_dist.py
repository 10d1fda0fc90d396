"""
Index sharding of sweeps over the ranks of a ``torch.distributed`` job.

The reference treats every grid point / start time as an independent loop iteration
(reference qnmfits/qnmfits.py:1271-1281, :1391-1410), so the flat fit index is split
into one contiguous slab per rank with no data-path exchange; the only collective is
the exchange of the per-fit mismatches (fused into the fit kernels over peer-mapped memory,
or an NCCL all-gather; gloo in the CPU tests of this logic).  A fit's arithmetic — including
the number of lanes that share it and the order in which their partial factors are combined
— is a function of the fit's own shape (rows, columns) only, never of the slab, rank or CTA
that runs it (qnmfit_api.cu, make_plan), so the gathered result is bit-identical to the
single-GPU one (tests/test_gpu_peer.py, test_gpu_parity.py::test_slab_launches...).
"""
import os


def world():
    """(rank, world_size) of the active process group, (0, 1) without one."""
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return 0, 1
    if dist.is_available() and dist.is_initialized() and \
            os.environ.get("QNMFITS_B200_NO_SHARD", "0") != "1":
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


_local_devices = None


def use_devices(devices=None):
    """Single-process multi-GPU: make the sweep functions (``mismatch_t0_array``,
    ``mismatch_M_chi_grid``, ...) split their fits over these CUDA devices, driven from
    this one process.  ``devices``: a list of device indices, ``"all"`` for every visible
    GPU, or ``None`` to go back to the current device only.  Ignored while a
    torch.distributed job is active (one process per GPU then)."""
    global _local_devices
    if devices is None:
        _local_devices = None
        return None
    import torch
    if isinstance(devices, str):
        if devices != "all":
            raise ValueError('devices must be a list of device indices, "all" or None')
        devices = list(range(torch.cuda.device_count()))
    devices = [int(d) for d in devices]
    if not devices or len(set(devices)) != len(devices) or \
            any(d < 0 or d >= torch.cuda.device_count() for d in devices):
        raise ValueError(f"bad device list {devices} ({torch.cuda.device_count()} visible)")
    _local_devices = devices
    return devices


def local_devices():
    return _local_devices


def shard_bounds(n_items, rank, world_size):
    """Contiguous slab [lo, hi) of rank and the common padded slab length."""
    per = -(-int(n_items) // int(world_size)) if n_items > 0 else 0
    lo = min(rank * per, n_items)
    hi = min(lo + per, n_items)
    return lo, hi, per


def all_gather_slabs(slab, n_items):
    """Concatenate equally sized (padded) 1-D slabs from all ranks; trim to n_items."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size()
    full = torch.empty(ws * slab.numel(), dtype=slab.dtype, device=slab.device)
    dist.all_gather_into_tensor(full, slab.contiguous())
    return full[:n_items]


# --------------------------------------------------------------------------
# Exchange fused into the fit kernels (include/qnmfit.h, qnmfit_fit_batch_peers)

class _DeviceMemory:
    """Wraps a raw device pointer so that torch can view it (``__cuda_array_interface__``)."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = dict(shape=(int(n_words),), typestr="<f8",
                                             data=(int(ptr), False), version=2)


class PeerWindow:
    """This rank's result window, mapped by every other rank of the node.

    Layout in 8-byte words: ``[0:8]`` epoch flags (u64, slot r written by rank r), then two
    result slots (epoch parity) of ``8 + capacity`` doubles each: the per-rank counts of
    flagged fits followed by the mismatch array of the whole sweep.  Two slots, because a
    rank may start the next sweep while a slower peer still copies the previous result to
    its host; it cannot get further ahead than that, since a sweep's kernel returns only
    after every peer has published the same epoch.
    """

    WORDS_HEAD = 8

    def __init__(self, eng, capacity):
        import torch
        import torch.distributed as dist
        from . import _cabi
        self.eng = eng
        self.rank, self.ws = dist.get_rank(), dist.get_world_size()
        self.capacity = int(capacity)
        self.slot_words = _cabi.MAX_PEERS + self.capacity
        self.n_words = self.WORDS_HEAD + 2 * self.slot_words
        # every rank takes part in the handle exchange even if its own allocation failed
        self.local_ptr, handle, error = None, None, None
        try:
            self.local_ptr, handle = eng.ctx.peer_alloc(8 * self.n_words)
        except Exception as exc:
            error = exc
        handles = [None] * self.ws
        dist.all_gather_object(handles, handle)
        if error is not None or any(h is None for h in handles):
            raise RuntimeError(f"peer window allocation failed on some rank ({error})")
        self.ptrs = []
        self._opened = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.local_ptr)
            else:
                ptr = eng.ctx.peer_open(h)
                self._opened.append(ptr)
                self.ptrs.append(ptr)
        self.view = torch.as_tensor(_DeviceMemory(self.local_ptr, self.n_words), device=eng.device)
        self.epoch = 0
        self.closed = False
        self._peers = {}
        self.timeout_ns = int(float(os.environ.get("QNMFITS_B200_PEER_TIMEOUT_S", "600")) * 1e9)

    def _slot_base(self, slot):
        return self.WORDS_HEAD + slot * self.slot_words

    def next_launch(self):
        """Descriptor of the next exchange: (qnmfit_peers, local mismatch pointer, slot)."""
        from . import _cabi
        self.epoch += 1
        slot = self.epoch & 1
        pe = self._peers.get(slot)
        if pe is None:                       # built once per slot; only the epoch changes
            base = self._slot_base(slot)
            pe = _cabi.Peers(n_peers=self.ws, rank=self.rank, timeout_ns=self.timeout_ns)
            for r, ptr in enumerate(self.ptrs):
                pe.flags[r] = ptr
                pe.flagged[r] = ptr + 8 * base
                pe.mismatch[r] = ptr + 8 * (base + _cabi.MAX_PEERS)
            self._peers[slot] = pe
        pe.epoch = self.epoch
        return pe, self.local_ptr + 8 * (self._slot_base(slot) + _cabi.MAX_PEERS), slot

    def result_ptr(self, slot):
        """Device address of slot ``slot``: [per-rank flagged counts (8) | mismatch ...]."""
        return self.local_ptr + 8 * self._slot_base(slot)

    def result(self, slot, n_items):
        """Device view of slot ``slot``: [per-rank flagged counts (8) | mismatch (n_items)]."""
        from . import _cabi
        base = self._slot_base(slot)
        return self.view[base:base + _cabi.MAX_PEERS + n_items]

    def close(self):
        import torch.distributed as dist
        self.closed = True               # prepared sweeps that hold this window must not reuse it
        self.eng.synchronize()
        dist.barrier()                       # nobody writes into a window that is going away
        for ptr in self._opened:
            self.eng.ctx.peer_close(ptr)
        self._opened = []
        dist.barrier()
        self.view = None
        self.eng.ctx.peer_free(self.local_ptr)


_windows = {}
_peer_disabled = False


def peer_window(eng, n_items):
    """The engine's PeerWindow with room for ``n_items`` results, or None when the fused
    exchange does not apply (one rank, not NCCL, more than 8 ranks, several nodes,
    QNMFITS_B200_PEER=0, or peer mapping failed on some rank — then every rank uses the
    NCCL all-gather).  Collective: all ranks call it with the same ``n_items``."""
    global _peer_disabled
    import torch
    import torch.distributed as dist
    from . import _cabi
    rank, ws = world()
    if ws == 1 or _peer_disabled or os.environ.get("QNMFITS_B200_PEER", "1") == "0":
        return None
    if dist.get_backend() != "nccl" or ws > _cabi.MAX_PEERS \
            or int(os.environ.get("LOCAL_WORLD_SIZE", ws)) != ws:
        return None
    win = _windows.get(eng.device_index)
    if win is not None and (win.rank, win.ws) != (rank, ws):
        del _windows[eng.device_index]       # another process group than the one it was built for
        win = None
    if win is not None and win.capacity >= n_items:
        return win
    if win is not None:
        win.close()
        del _windows[eng.device_index]
    capacity = max(1 << 20, 1 << int(n_items - 1).bit_length())
    ok = torch.ones(1, dtype=torch.int32, device=eng.device)
    try:
        win = PeerWindow(eng, capacity)
    except Exception as exc:                 # e.g. no peer access between two GPUs
        import warnings
        warnings.warn(f"qnmfits_b200: peer mapping failed ({exc}); using the NCCL all-gather")
        win = None
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        _peer_disabled = True
        return None
    _windows[eng.device_index] = win
    return win

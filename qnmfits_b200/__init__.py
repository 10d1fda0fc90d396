"""
qnmfits_b200 — B200-native (sm_100a) implementation of the batched complex
least-squares ringdown-fitting path of eliotfinch/qnmfits, behind that package's own
Python API.  ``import qnmfits_b200 as qnmfits`` is the intended drop-in for
``ringdown_fit``, ``multimode_ringdown_fit``, ``mismatch_t0_array``,
``mismatch_M_chi_grid``, ``mismatch`` / ``multimode_mismatch`` (+ ``calculate_mismatch``)
and the ``qnm`` provider instance.

Like the reference (qnmfits/__init__.py:5-7) the star import below re-exports the
module-level provider *instance* ``qnm`` which shadows the class of the same name.
"""
from .qnm import qnm, set_table_provider  # noqa: F401  (class; shadowed below)
from .qnmfits import *  # noqa: F401,F403  (functions + the `qnm` instance)
from .qnmfits import clear_sweep_cache  # noqa: F401
from ._dist import use_devices  # noqa: F401  (single-process multi-GPU sweeps)

__version__ = "0.1.0"

"""
Synthetic Kerr-like QNM table provider (benchmark / test input generator).

The reference tabulates frequencies and mixing coefficients from the third-party
``qnm`` PyPI package through exactly one call, ``qnm.modes_cache(s, l, m, n)``
(reference ``qnmfits/qnm.py:134``), and touches only three attributes of the
returned object: ``.a`` (spin grid), ``.omega`` (complex frequencies on that grid)
and ``.C`` (mixing coefficients, one column per spherical ell') — reference
``qnmfits/qnm.py:137-141``.  That package needs a network download and is absent
from the build and GPU boxes, so benchmarks and tests use this deterministic,
closed-form stand-in with the same interface.  It is NOT physics: it is a smooth
overtone ladder with Kerr-like magnitudes (Im w ~ -(0.089 + 0.185 n), Re w rising
with m*a) so that design matrices have realistic conditioning (cond ~ 1e5 for eight
overtones, T = 100 M).  The same object feeds the reference (via the oracle's stub
loader) and this package, which is all parity needs.
"""
import numpy as np

#: largest spherical ell' carried by the synthetic mixing table
L_MAX = 12
#: number of spin samples of every synthetic sequence
N_SPIN = 200


class SyntheticSequence:
    """Mimics the attributes of ``qnm``'s ``KerrSpinSeq`` that the reference reads."""

    def __init__(self, s, l, m, n):
        self.s, self.l, self.m, self.n = s, l, m, n
        # Non-uniform spin grid, denser towards extremality like the real tables.
        x = np.linspace(0.0, 1.0, N_SPIN)
        a = 0.99 * (1.0 - (1.0 - x) ** 1.6)
        self.a = a

        # Frequencies: a smooth overtone ladder.
        re0 = 0.3737 + 0.2257 * (l - 2) - 0.0265 * n / (1.0 + 0.35 * (l - 2))
        re = re0 + m * a * (0.0629 + 0.0581 * a * a) / (1.0 + 0.11 * n) \
            + 0.0113 * a * a * (l - abs(m))
        im = -(0.0890 + 0.1852 * n + 0.0021 * (l - 2)) \
            * (1.0 - 0.1013 * a * a - 0.031 * m * a / (l + 1.0))
        self.omega = re + 1j * im

        # Mixing coefficients: column index = ell' - max(|m|, |s|)
        # (reference qnmfits/qnm.py:345-348).
        l_min = max(abs(m), abs(s))
        n_col = L_MAX - l_min + 1
        C = np.zeros((N_SPIN, n_col), dtype=complex)
        eps = 0.081 * a * (1.0 + 0.1j * (n + 1)) * (1.0 + 0.05 * m)
        for col in range(n_col):
            lp = l_min + col
            d = abs(lp - l)
            if d == 0:
                C[:, col] = 1.0 - 0.5 * np.abs(eps) ** 2
            else:
                sgn = 1.0 if lp > l else -1.0
                C[:, col] = sgn * eps ** d / (1.0 + 0.3 * (d - 1))
        self.C = C


_cache = {}


def modes_cache(s, l, m, n):
    """Drop-in for ``qnm.modes_cache(s, l, m, n)`` (reference ``qnmfits/qnm.py:134``)."""
    if l < max(abs(m), abs(s)):
        raise KeyError(f"no sequence for s={s}, l={l}, m={m}")
    key = (int(s), int(l), int(m), int(n))
    if key not in _cache:
        _cache[key] = SyntheticSequence(*key)
    return _cache[key]

"""
Synthetic injected-QNM workloads of BASELINE.json's configs.

Input generators only (seeded, deterministic): used by bench.py, the tests and
``__graft_entry__.smoke()`` so that every arm — reference, oracle, CUDA — is fed the
same arrays.  ``times = np.arange(-500, 1501) * 0.1`` (dt = 0.1 M, K_tot = 2001); truth
(Mf, chif) = (0.95, 0.69); amplitudes from ``default_rng(0)``; complex white noise of
sigma 1e-6 from ``default_rng(1)``; the signal is zero before t = 0
(reference ``ringdown``, qnmfits/qnmfits.py:15-70).
"""
from dataclasses import dataclass, field

import numpy as np

from . import synthetic
from .qnm import set_table_provider
from .qnmfits import qnm, ringdown

MF_TRUE, CHIF_TRUE = 0.95, 0.69


def use_synthetic_tables():
    """Serve ``qnm.modes_cache`` from the closed-form synthetic provider."""
    set_table_provider(synthetic.modes_cache)


def use_kerr_tables():
    """Serve ``qnm.modes_cache`` from the built-in Leaver solver (``qnmfits_b200.kerr``)."""
    from . import kerr
    set_table_provider(kerr.modes_cache)


def synthetic_modes_cache():
    return synthetic.modes_cache


@dataclass
class Workload:
    name: str
    times: np.ndarray
    data: object                     # ndarray (single series) or dict {(l, m): ndarray}
    modes: list
    t0: float = 0.0
    T: float = 100.0
    Mf: float = MF_TRUE
    chif: float = CHIF_TRUE
    Mf_minmax: tuple = (0.85, 1.05)
    chif_minmax: tuple = (0.59, 0.79)
    res: int = 256
    t0_array: np.ndarray = None
    spherical_modes: list = None
    extra: dict = field(default_factory=dict)


def default_times():
    return np.arange(-500, 1501) * 0.1


def overtone_modes(n_overtones=8, ell=2, m=2):
    return [(ell, m, n, 1) for n in range(n_overtones)]


def _noise(rng, n, sigma):
    return sigma * (rng.normal(size=n) + 1j * rng.normal(size=n))


def injected_series(times, modes, Mf=MF_TRUE, chif=CHIF_TRUE, sigma=1e-6):
    """h(t) = sum_j C_j exp(-i w_j t) for t >= 0, plus noise (single series)."""
    omega = np.array(qnm.omega_list(modes, chif, Mf))
    rng = np.random.default_rng(0)
    C = rng.normal(size=len(modes)) + 1j * rng.normal(size=len(modes))
    h = ringdown(times, 0.0, C, omega)
    return h + _noise(np.random.default_rng(1), len(times), sigma), C


def config1(n_overtones=8):
    """ringdown_fit of h22 with (2,2,n,+1), n < n_overtones, t0 = 0, T = 100."""
    times = default_times()
    modes = overtone_modes(n_overtones)
    data, C = injected_series(times, modes)
    return Workload("cfg1_ringdown_fit", times, data, modes, extra={"C_true": C})


def config2(n_t0=1000, n_overtones=8):
    """mismatch_t0_array: start times linspace(-10, 60, n_t0)."""
    wl = config1(n_overtones)
    wl.name = "cfg2_mismatch_t0_array"
    wl.t0_array = np.linspace(-10.0, 60.0, n_t0)
    return wl


def config3(res=256, n_overtones=8):
    """mismatch_M_chi_grid: res x res grid around the truth, 8 overtones on h22."""
    wl = config1(n_overtones)
    wl.name = "cfg3_mismatch_M_chi_grid"
    wl.res = res
    return wl


def multimode_labels():
    """21 spherical modes (ell = 2..4) and 40 QNMs: regular and mirror, three m's."""
    spherical = [(ell, m) for ell in (2, 3, 4) for m in range(-ell, ell + 1)]
    modes = []
    for m in (2, -2):
        for ell in (2, 3, 4):
            ell_eff = max(ell, abs(m))
            modes += [(ell_eff, m, n, 1) for n in range(4)]
        for ell in (2, 3):
            modes += [(ell, m, n, -1) for n in range(2)]
    for ell, n_max in ((2, 4), (3, 2), (4, 2)):
        modes += [(ell, 0, n, 1) for n in range(n_max)]
    return spherical, modes


QUADRATIC_LABELS = [
    (2, 2, 0, 1, 2, 2, 0, 1),       # m = 4: sourced in (4, 4)
    (2, 2, 0, 1, 2, 2, 1, 1),
    (2, -2, 0, 1, 2, -2, 0, 1),     # m = -4
    (2, 2, 0, 1, 2, -2, 0, 1),      # m = 0: the memory-like product
]


def multimode_labels_quadratic():
    """BASELINE.json config 4 as stated: 21 spherical modes (ell <= 4) and ~40 QNMs — regular,
    mirror AND quadratic.  36 linear labels (``multimode_labels`` without the (3,0,n) and
    (4,0,n) overtones) + the four ``QUADRATIC_LABELS``."""
    spherical, modes = multimode_labels()
    modes = [mode for mode in modes if not (mode[1] == 0 and mode[0] > 2)]
    return spherical, modes + list(QUADRATIC_LABELS)


def quadratic_alpha(label, spherical, chif):
    """Synthetic per-series coefficients of a quadratic QNM (the role of ``Qmu_B`` in the
    reference, spatial_mapping_functions.py:202-210, whose dependencies — spherical.Wigner3j,
    spin-weight-0 Kerr sequences — are absent here): non-zero only in the spherical modes with
    m = m1 + m2, smooth in the spin, decaying with ell away from ell1 + ell2.  Deterministic
    input generator, not physics.  Returns complex (L,) for a scalar spin, (n, L) for an array."""
    l1, m1, n1, _, l2, m2, n2, _ = label
    chif = np.asarray(chif, dtype=float)
    out = np.zeros(chif.shape + (len(spherical),), dtype=complex)
    for i, (ell, m) in enumerate(spherical):
        if m != m1 + m2 or ell < abs(m):
            continue
        out[..., i] = (0.21 + 0.05j * (1 + n1 + n2)) * (1.0 - 0.2 * chif + 0.1j * chif ** 2) \
            / (1.0 + 0.6 * abs(l1 + l2 - ell))
    return out


def quadratic_columns(spherical, labels=None):
    """``coef_columns`` argument of the public API for the synthetic quadratic coefficients."""
    labels = QUADRATIC_LABELS if labels is None else labels
    return {label: (lambda chif, label=label: quadratic_alpha(label, spherical, chif)) for label in labels}


def coef_override(spherical, modes, chif, tables=None):
    """The oracle's ``coef_override`` (L, N) that corresponds to ``quadratic_columns``: mu for
    linear labels, the synthetic alpha for quadratic ones."""
    provider = qnm if tables is None else tables
    out = np.zeros((len(spherical), len(modes)), dtype=complex)
    for j, mode in enumerate(modes):
        if len(mode) == 4:
            out[:, j] = [complex(v) for v in provider.mu_list([lm + mode for lm in spherical], chif)]
        else:
            out[:, j] = quadratic_alpha(mode, spherical, chif)
    return out


def config4(n_t0=500, spherical=None, modes=None, sigma=1e-6, quadratic=False):
    """multimode sweep: 21 series x 40 QNMs with spheroidal mixing and mirror modes.

    ``quadratic=False``: linear labels only — the shape the reference's
    multimode_ringdown_fit can express (qnm.py:390 raises for nonlinear labels), pinned to the
    golden fixtures.  ``quadratic=True``: config 4 as BASELINE.json states it, with quadratic
    QNMs entering through caller-supplied coefficient columns (``extra['coef_columns']`` for the
    public API, ``coef_override`` for the oracle).
    """
    if spherical is None or modes is None:
        spherical, modes = multimode_labels_quadratic() if quadratic else multimode_labels()
    times = default_times()
    omega = np.array(qnm.omega_list(modes, CHIF_TRUE, MF_TRUE))
    rng = np.random.default_rng(0)
    C = rng.normal(size=len(modes)) + 1j * rng.normal(size=len(modes))
    noise_rng = np.random.default_rng(1)
    coef = coef_override(spherical, modes, CHIF_TRUE)
    data = {}
    for i, lm in enumerate(spherical):
        data[lm] = ringdown(times, 0.0, coef[i] * C, omega) + _noise(noise_rng, len(times), sigma)
    extra = {"C_true": C}
    if any(len(mode) != 4 for mode in modes):
        extra["coef_columns"] = quadratic_columns(spherical, [m for m in modes if len(m) != 4])
    return Workload("cfg4_multimode_t0_sweep", times, data, modes,
                    t0_array=np.linspace(0.0, 50.0, n_t0), spherical_modes=spherical,
                    extra=extra)


def config5(n_waveforms=4096, n_fixed=2, sigma=1e-6):
    """Free-frequency fitting: n_waveforms series, each = n_fixed fixed Kerr modes
    ((2,2,0,+1), (2,2,1,+1)) + one damped sinusoid whose frequency is drawn uniformly in
    [0.3, 1.7] x [-0.9, -0.05] (inside the optimiser's box), default_rng(2), noise sigma.
    t0 = 0, T = 100."""
    times = default_times()
    modes = overtone_modes(n_fixed)
    fixed = np.array(qnm.omega_list(modes, CHIF_TRUE, MF_TRUE)) if n_fixed else np.zeros(0, complex)
    rng = np.random.default_rng(2)
    w_free = rng.uniform(0.3, 1.7, n_waveforms) + 1j * rng.uniform(-0.9, -0.05, n_waveforms)
    amps = rng.normal(size=(n_waveforms, n_fixed + 1)) + 1j * rng.normal(size=(n_waveforms, n_fixed + 1))
    noise_rng = np.random.default_rng(3)
    data = np.empty((n_waveforms, len(times)), dtype=complex)
    for b in range(n_waveforms):
        om = np.concatenate([fixed, w_free[b:b + 1]])
        data[b] = ringdown(times, 0.0, amps[b], om) + _noise(noise_rng, len(times), sigma)
    return Workload("cfg5_free_frequency", times, data, modes, extra={"omega_free": w_free, "amps": amps})

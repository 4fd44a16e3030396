"""Host-side transmittance helpers with the reference's interface (ref crt1d/solvers/common.py:11-95).

Used only by the plugin path, where `K_b_fn` is an arbitrary Python callable that has to be integrated
on the host with the same scipy QUADPACK calls as the reference, so that the band-independent scalars
handed to the CUDA kernels are bit-identical to the reference's.  The batched path evaluates the same
integrals on the device (`crt1d_tau_d`, `crt1d_leaf_integrals`) for parametric leaf-angle families.
"""
import math

import numpy as np
import scipy.integrate as integrate


def tau_b_fn(K_b_fn, psi, lai):
    """Direct-beam transmittance through LAI `lai` at zenith angle `psi`  (ref common.py:11-27)."""
    return np.exp(-K_b_fn(psi) * lai)


def _tau_df_quad(K_b_fn, lai_val):
    def integrand(psi):
        return tau_b_fn(K_b_fn, psi, lai_val) * np.sin(psi) * np.cos(psi)

    return 2 * integrate.quad(integrand, 0, np.pi / 2, epsrel=1e-9)[0]  # ref common.py:36-37


def _tau_df_9sky(K_b_fn, lai_val):
    total = 0
    for sza in range(5, 90, 10):  # ref common.py:46-51
        psi = math.radians(sza)
        total += tau_b_fn(K_b_fn, psi, lai_val) * math.sin(psi) * math.cos(psi)
    return total * 2 * math.radians(10)


def tau_df_fn(K_b_fn, lai, *, method="quad"):
    """Hemispherical transmittance of diffuse light, scalar or array `lai`  (ref common.py:56-87)."""
    try:
        f = {"quad": _tau_df_quad, "9sky": _tau_df_9sky}[method]
    except KeyError:
        raise ValueError("invalid `method`. Valid options are 'quad' and '9sky'.") from None
    if np.isscalar(lai):
        return f(K_b_fn, lai)
    lai = np.asarray(lai)
    return np.array([f(K_b_fn, v) for v in lai], dtype=float).reshape(lai.shape)


def K_df_fn(K_b_fn, lai_tot, **kwargs):
    """K_d from tau_d at total LAI  (ref common.py:90-95)."""
    return -np.log(tau_df_fn(K_b_fn, lai_tot, **kwargs)) / lai_tot


def mu_bar_fn(G_fn):
    """2s: average inverse diffuse optical depth per unit leaf area  (ref _solve_2s.py:32)."""
    return integrate.quad(lambda sa: math.cos(sa) / G_fn(sa) * -math.sin(sa), math.pi / 2, 0)[0]


def G_sector_integrals(G_fn, mu_s):
    """4s: integrals of G(arccos mu') over [0, mu_s] and [mu_s, 1]  (ref _solve_4s.py:148-149)."""
    g1 = integrate.quad(lambda m: G_fn(np.arccos(m)), 0, mu_s)[0]
    g2 = integrate.quad(lambda m: G_fn(np.arccos(m)), mu_s, 1)[0]
    return g1, g2


def mean_dlai(lai):
    """zq: the single layer thickness used for every layer  (ref _solve_zq.py:50)."""
    dlai = np.diff(lai)
    return np.abs(np.mean(dlai[dlai != 0]))

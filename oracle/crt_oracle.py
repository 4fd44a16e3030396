"""CPU oracle: a numpy/scipy restatement of the reference's solver hot path (zmoon/crt1d).

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module, and there only as the
checker or as the CPU baseline being reported; `crt1d_b200/` never imports it and has no CPU fallback.

Parity status: PINNED.  Every function here is checked element-wise against outputs of the unmodified
reference (imported from /root/reference in the build container by `tests/golden/make_golden.py`),
stored as fixtures under `tests/golden/` (see tests/test_oracle_golden.py), and against the reference's
own known-answer vectors where it has any (`tests/test_spectra.py:25-35` for the band weights; the n79
Bonan golden file is not vendored by the reference and is fetched from the network by its test, so the
n79 pin is a reference-generated fixture of the same SP 14.3 set-up).

Each solver takes exactly the reference's keyword arguments and returns the same dict of
`(n_z, n_wl)` float64 arrays, but is written band-vectorised (no Python loop over bands), so it is an
independent statement of the same arithmetic rather than a transcription.

Third-party arithmetic at the same call sites as the reference (scipy is an unpinned dependency of the
reference, `setup.cfg:14-21`; this image has scipy 1.18.1):
  * `scipy.integrate.quad` (QUADPACK QAGS): mu_bar (2s), G sector integrals (4s), tau_d (bl/n79/zq).
  * `scipy.integrate.solve_bvp`: 4s (4th-order Lobatto IIIA collocation, `tol=1e-6` as shipped).
  * SuperLU `spsolve` in zq is replaced by LAPACK `gtsv` via `scipy.linalg.solve_banded` (both are
    partial-pivoting direct solves of the same tridiagonal matrix; agreement ~1e-14, pinned by golden).
"""
import math

import numpy as np
import scipy.integrate as integrate
from scipy.linalg import solve_banded

RET_KEYS_ALL_SCHEMES = ["I_dr", "I_df_d", "I_df_u", "F"]  # ref solvers/__init__.py:34


# ----------------------------------------------------------------------------------------------
# common transmittance helpers  (ref solvers/common.py:11-95)
# ----------------------------------------------------------------------------------------------
def tau_b_fn(K_b_fn, psi, lai):
    """Direct-beam transmittance exp(-K_b(psi) L)  (ref common.py:11-27)."""
    return np.exp(-K_b_fn(psi) * lai)


def _tau_df_scalar_quad(K_b_fn, L):
    """2 int_0^{pi/2} tau_b(psi) sin cos dpsi by QAGS with epsrel=1e-9  (ref common.py:30-37)."""
    f = lambda psi: np.exp(-K_b_fn(psi) * L) * np.sin(psi) * np.cos(psi)  # noqa: E731
    return 2 * integrate.quad(f, 0, np.pi / 2, epsrel=1e-9)[0]


def _tau_df_scalar_9sky(K_b_fn, L):
    """Nine 10-degree sky sectors  (ref common.py:40-53)."""
    acc = 0
    for deg in (5, 15, 25, 35, 45, 55, 65, 75, 85):
        psi = math.radians(deg)
        acc += np.exp(-K_b_fn(psi) * L) * math.sin(psi) * math.cos(psi)
    return acc * (2 * math.radians(10))


def tau_df_fn(K_b_fn, lai, *, method="quad"):
    """Hemispherical (diffuse) transmittance for scalar or array LAI  (ref common.py:56-87)."""
    if method == "quad":
        f = _tau_df_scalar_quad
    elif method == "9sky":
        f = _tau_df_scalar_9sky
    else:
        raise ValueError("invalid `method`. Valid options are 'quad' and '9sky'.")
    if np.isscalar(lai):
        return f(K_b_fn, lai)
    out = np.zeros_like(lai)
    for i, L in enumerate(lai):
        out[i] = f(K_b_fn, L)
    return out


def K_df_fn(K_b_fn, lai_tot, **kw):
    """K_d = -ln(tau_d)/L  (ref common.py:90-95)."""
    return -np.log(tau_df_fn(K_b_fn, lai_tot, **kw)) / lai_tot


def mu_bar_quad(G_fn):
    """Sellers' mean inverse diffuse optical depth per unit leaf area, same quad call as ref _solve_2s.py:32."""
    return integrate.quad(lambda sa: math.cos(sa) / G_fn(sa) * -math.sin(sa), math.pi / 2, 0)[0]


def G_sector_integrals(G_fn, mu_s):
    """int G(arccos mu') dmu' over [0, mu_s] and [mu_s, 1], same quad calls as ref _solve_4s.py:148-149."""
    g1 = integrate.quad(lambda m: G_fn(np.arccos(m)), 0, mu_s)[0]
    g2 = integrate.quad(lambda m: G_fn(np.arccos(m)), mu_s, 1)[0]
    return g1, g2


# ----------------------------------------------------------------------------------------------
# 2s  Dickinson-Sellers two-stream  (ref solvers/_solve_2s.py:11-163)
# ----------------------------------------------------------------------------------------------
def solve_2s(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn, G_fn, mla):
    K = K_b_fn(psi)  # black-leaf extinction (ref :26, :40)
    mu = math.cos(psi)
    cos2_tl = math.cos(math.radians(mla)) ** 2  # ref :28, :68
    mu_bar = mu_bar_quad(G_fn)  # ref :32

    L = np.asarray(lai, dtype=float)[:, None]  # (nz, 1)
    L_T = lai[0]
    alpha = np.asarray(leaf_r, dtype=float)
    tau = np.asarray(leaf_t, dtype=float)
    rho_s = np.asarray(soil_r, dtype=float)

    omega = alpha + tau  # ref :65
    beta = (0.5 * (alpha + tau + (alpha - tau) * cos2_tl)) / omega  # eq. 3, ref :68
    a_s = omega / 2 * (1 - mu * math.log((mu + 1) / mu))  # ref :73 (spherical form regardless of G_fn)
    beta_0 = (1 + mu_bar * K) / (omega * mu_bar * K) * a_s  # eq. 4, ref :76

    b = 1 - (1 - beta) * omega  # ref :80-85
    c = omega * beta
    d = omega * mu_bar * K * beta_0
    f = omega * mu_bar * K * (1 - beta_0)
    h = np.sqrt(b**2 - c**2) / mu_bar
    sigma = (mu_bar * K) ** 2 + c**2 - b**2

    u1 = b - c / rho_s  # ref :87-97
    u2 = b - c * rho_s
    u3 = f + c * rho_s
    S1 = np.exp(-h * L_T)
    S2 = math.exp(-K * L_T)
    p1 = b + mu_bar * h
    p2 = b - mu_bar * h
    p3 = b + mu_bar * K
    p4 = b - mu_bar * K
    D1 = p1 * (u1 - mu_bar * h) / S1 - p2 * (u1 + mu_bar * h) * S1
    D2 = (u2 + mu_bar * h) / S1 - (u2 - mu_bar * h) * S1

    h1 = -d * p4 - c * f  # ref :99-120
    h2 = 1 / D1 * (
        (d - h1 / sigma * p3) * (u1 - mu_bar * h) / S1
        - p2 * (d - c - h1 / sigma * (u1 + mu_bar * K)) * S2
    )
    h3 = -1 / D1 * (
        (d - h1 / sigma * p3) * (u1 + mu_bar * h) * S1
        - p1 * (d - c - h1 / sigma * (u1 + mu_bar * K)) * S2
    )
    h4 = -f * p3 - c * d
    h5 = -1 / D2 * (
        h4 / sigma * (u2 + mu_bar * h) / S1 + (u3 - h4 / sigma * (u2 - mu_bar * K)) * S2
    )
    h6 = 1 / D2 * (
        h4 / sigma * (u2 - mu_bar * h) * S1 + (u3 - h4 / sigma * (u2 - mu_bar * K)) * S2
    )
    h7 = c / D1 * (u1 - mu_bar * h) / S1
    h8 = -c / D1 * (u1 + mu_bar * h) * S1
    h9 = 1 / D2 * (u2 + mu_bar * h) / S1
    h10 = -1 / D2 * (u2 - mu_bar * h) * S1

    eK = np.exp(-K * L)  # (nz, 1)
    em = np.exp(-h * L)  # (nz, nb)
    ep = np.exp(h * L)
    I_df_u = I_dr0_all * (h1 * eK / sigma + h2 * em + h3 * ep) + I_df0_all * (h7 * em + h8 * ep)  # ref :125-135
    I_df_d = I_dr0_all * (h4 * eK / sigma + h5 * em + h6 * ep) + I_df0_all * (h9 * em + h10 * ep)
    I_dr = I_dr0_all * eK  # ref :150
    F = I_dr / mu + 2 * I_df_u + 2 * I_df_d  # ref :156
    return {"I_dr": I_dr, "I_df_d": I_df_d, "I_df_u": I_df_u, "F": F}


# ----------------------------------------------------------------------------------------------
# bl  Beer-Lambert  (ref solvers/_solve_bl.py:9-93)
# ----------------------------------------------------------------------------------------------
def solve_bl(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, K_b_fn):
    mu = np.cos(psi)
    K_b = K_b_fn(psi)
    lai = np.asarray(lai, dtype=float)
    tau_b = np.exp(-K_b * lai)[:, None]  # ref :31
    tau_d = np.array([tau_df_fn(K_b_fn, L) for L in lai])[:, None]  # ref :35-37
    k_prime = np.sqrt(1 - (leaf_t + leaf_r))  # ref :58-60
    tau_g = np.exp(-(K_b * k_prime) * lai[:, None])  # ref :62-65
    I_dr = I_dr0_all * tau_b  # ref :69
    I_df_d = I_df0_all * tau_d + 0.5 * (I_dr0_all * (tau_g - tau_b))  # ref :70-79
    I_df_u = np.zeros_like(I_dr)  # ref :87
    F = I_dr / mu + 2 * I_df_d  # ref :90
    return dict(I_dr=I_dr, I_df_d=I_df_d, I_df_u=I_df_u, F=F)


# ----------------------------------------------------------------------------------------------
# bf  Bodin & Franklin (2012)  (ref solvers/_solve_bf.py:7-154)
# g77 Goudriaan (1977)        (ref solvers/_solve_g77.py:7-135)
# ----------------------------------------------------------------------------------------------
def _bf_g77_common(psi, lai, leaf_t, leaf_r, K_b_fn):
    k_b = K_b_fn(psi)
    mu = np.cos(psi)
    lai = np.asarray(lai, dtype=float)
    lai_tot = lai[0]
    assert lai_tot == lai.max()  # ref _solve_bf.py:40 / _solve_g77.py:35
    sigma = leaf_r + leaf_t
    k_prime = np.sqrt(1 - sigma)
    rho_c = ((1 - k_prime) / (1 + k_prime)) * (2 / (1 + 1.6 * mu))  # Spitters (1986) eq. 1
    k_d = 0.8 * np.sqrt(1 - sigma)  # B&F eq. 2
    return k_b, mu, lai[:, None], lai_tot, sigma, k_prime, rho_c, k_d


def _bf_g77_absorbed(A_sl, k_d, k_prime, r_l, t_l, I_df, I_sc_u, I_sc_d, k_b, I_dr0):
    """B&F eq. 14 (shaded) and eq. 15 (sunlit)  (ref _solve_bf.py:118-130, _solve_g77.py:99-111)."""
    diff = k_d / k_prime * I_df + k_d / np.sqrt(1 - r_l) * I_sc_u + k_d / np.sqrt(1 - t_l) * I_sc_d
    sh = (1 - A_sl) * diff
    sl = A_sl * (diff + k_b * I_dr0)
    return sl, sh


def solve_bf(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn):
    k_b, mu, L, L_T, sigma, k_prime, rho_c, k_d = _bf_g77_common(psi, lai, leaf_t, leaf_r, K_b_fn)
    eb = np.exp(-k_b * L)  # (nz,1)
    ed = np.exp(-k_d * L)  # (nz,nb)
    I_df = I_df0_all * ed  # ref :86
    I_dr = I_dr0_all * eb  # ref :90
    A_sl = eb  # ref :93
    I_sc_d = I_dr0_all * leaf_t * ((eb - ed) / (k_d - k_b))  # eq. 8, ref :97
    I_sc_u = I_dr0_all * leaf_r * ((eb - np.exp(+k_d * L - (k_b + k_d) * L_T)) / (k_d + k_b))  # eq. 9, ref :101-105
    I_sr = soil_r * (I_dr0_all * A_sl[0] + I_df[0] + I_sc_d[0]) * np.exp(-k_d * (L_T - L))  # eq. 11, ref :114
    sl, sh = _bf_g77_absorbed(A_sl, k_d, k_prime, leaf_r, leaf_t, I_df, I_sc_u, I_sc_d, k_b, I_dr0_all)
    I_df_d = I_sc_d + I_df
    I_df_u = I_sc_u + I_sr
    F = I_dr / mu + 2 * I_df_u + 2 * I_df_d
    return dict(
        I_dr=I_dr, I_df_d=I_df_d, I_df_u=I_df_u, F=F, aI_lsl=sl, aI_lsh=sh, aI_l=sl + sh,
        rho_c=rho_c[-1],  # the reference returns the LAST band's scalar (ref :153)
    )


def solve_g77(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn):
    k_b, mu, L, L_T, sigma, k_prime, rho_c, k_d = _bf_g77_common(psi, lai, leaf_t, leaf_r, K_b_fn)
    eb = np.exp(-k_b * L)
    I_df = I_df0_all * (1 - rho_c) * np.exp(-k_d * L)  # ref :73
    I_dr = I_dr0_all * eb
    A_sl = eb
    I_sc = I_dr0_all * (1 - rho_c) * np.exp(-k_prime * k_b * L) + -I_dr0_all * (1 - sigma) * eb  # eq. 5, ref :84-86
    I_sc_d = 0.5 * I_sc
    I_sc_u = 0.5 * I_sc
    I_sr = soil_r * (I_dr0_all * A_sl[0] + I_df[0] + I_sc_d[0]) * np.exp(-k_d * (L_T - L))  # ref :95
    sl, sh = _bf_g77_absorbed(A_sl, k_d, k_prime, leaf_r, leaf_t, I_df, I_sc_u, I_sc_d, k_b, I_dr0_all)
    I_df_d = I_sc_d + I_df
    I_df_u = I_sc_u + I_sr
    F = I_dr / mu + 2 * I_df_u + 2 * I_df_d
    return dict(I_dr=I_dr, I_df_d=I_df_d, I_df_u=I_df_u, F=F, aI_lsl=sl, aI_lsh=sh, aI_l=sl + sh)


# ----------------------------------------------------------------------------------------------
# n79  Norman (1979) after Bonan SP 14.3  (ref solvers/_solve_n79.py:11-200)
# ----------------------------------------------------------------------------------------------
def thomas_rows(a, b, c, d):
    """Thomas algorithm with the reference's arithmetic order (ref _solve_n79.py:167-200), vectorised
    over trailing axes: a, b, c, d are (n, ...) arrays (sub-, main, super-diagonal, rhs)."""
    n = a.shape[0]
    e = np.zeros_like(d)
    f = np.zeros_like(d)
    e[0] = c[0] / b[0]
    f[0] = d[0] / b[0]
    for i in range(1, n):
        den = b[i] - a[i] * e[i - 1]
        if i < n - 1:
            e[i] = c[i] / den
        f[i] = (d[i] - a[i] * f[i - 1]) / den
    u = np.zeros_like(d)
    u[-1] = f[-1]
    for i in range(n - 2, -1, -1):
        u[i] = f[i] - e[i] * u[i + 1]
    return u


def solve_n79(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn, tau_d_method="quad"):
    K_b = K_b_fn(psi)
    lai = np.asarray(lai, dtype=float)
    rho = np.asarray(leaf_r, dtype=float)
    tau = np.asarray(leaf_t, dtype=float)
    nz, nb = lai.size, tau.size
    dlai = lai[:-1] - lai[1:]  # ref :41
    tb = tau_b_fn(K_b_fn, psi, dlai)  # ref :45
    tbcum = np.exp(-K_b * lai)  # ref :46
    td = tau_df_fn(K_b_fn, dlai, method=tau_d_method)  # ref :53
    omega = rho + tau
    laim = (lai[:-1] + lai[1:]) / 2
    fracsun = np.exp(-K_b * laim)  # ref :56-59
    fracsha = 1 - fracsun

    a = np.zeros((2 * nz, nb))
    b = np.ones((2 * nz, nb))
    c = np.zeros((2 * nz, nb))
    d = np.zeros((2 * nz, nb))

    def up_row(k_td, k_cum, k_tb):
        refld = (1 - td[k_td]) * rho
        trand = (1 - td[k_td]) * tau + td[k_td]
        fiv = refld - trand * trand / refld
        eiv = trand / refld
        return -eiv, -fiv, I_dr0_all * tbcum[k_cum] * (1 - tb[k_tb]) * (rho - tau * eiv)

    def dn_row(k_td, k_cum, k_tb):
        refld = (1 - td[k_td]) * rho
        trand = (1 - td[k_td]) * tau + td[k_td]
        aiv = refld - trand * trand / refld
        biv = trand / refld
        return -aiv, -biv, I_dr0_all * tbcum[k_cum] * (1 - tb[k_tb]) * (tau - rho * biv)

    # soil rows (ref :79-92; the downward row uses index 1 of td/tb/tbcum as shipped)
    c[0] = -soil_r
    d[0] = I_dr0_all * tbcum[0] * soil_r
    a[1], c[1], d[1] = dn_row(1, 1, 1)
    # interior layers (ref :95-119)
    for j in range(nz - 2):
        ju = 2 * (j + 1)
        a[ju], c[ju], d[ju] = up_row(j, j + 1, j)
        a[ju + 1], c[ju + 1], d[ju + 1] = dn_row(j + 1, j + 2, j + 1)
    # top layer, upward (ref :122-130); top boundary, downward (ref :132-135)
    a[-2], c[-2], d[-2] = up_row(-1, -1, -1)
    a[-1] = 0
    c[-1] = 0
    d[-1] = I_df0_all

    u = thomas_rows(a, b, c, d)
    swup = u[::2]
    swdn = u[1::2]

    direct = I_dr0_all * tbcum[1:, None] * (1 - tb)[:, None] * (1 - omega)  # ref :145-148
    diffuse = (swdn[1:] + swup[:-1]) * (1 - td)[:, None] * (1 - omega)
    sun = diffuse * fracsun[:, None] + direct
    shade = diffuse * fracsha[:, None]
    I_dr = I_dr0_all * tbcum[:, None]
    return {
        "I_dr": I_dr,
        "I_df_d": swdn,
        "I_df_u": swup,
        "F": I_dr / np.cos(psi) + 2 * swdn + 2 * swup,  # ref :161
        "aI_lsl": sun / (fracsun * dlai)[:, None],  # ref :154-155
        "aI_lsh": shade / (fracsha * dlai)[:, None],
    }


# ----------------------------------------------------------------------------------------------
# zq  Zhao & Qualls (2005)  (ref solvers/_solve_zq.py:13-229)
# ----------------------------------------------------------------------------------------------
def zq_layer_scalars(psi, lai, K_b_fn):
    """Band-independent prologue: one mean tau_i / tau_psi for all layers, as shipped (ref :50-52)."""
    dlai = np.diff(lai)
    dlai_mean = np.abs(np.mean(dlai[dlai != 0]))
    return tau_df_fn(K_b_fn, dlai_mean), tau_b_fn(K_b_fn, psi, dlai_mean)


def solve_zq(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn, G_fn):
    lai = np.asarray(lai, dtype=float)
    mu = np.cos(psi)
    K = K_b_fn(psi)
    tau_i, t_psi = zq_layer_scalars(psi, lai, K_b_fn)
    m = lai.size
    nb = I_dr0_all.size
    n = 2 * m + 2
    out = {k: np.zeros((m, nb)) for k in ("I_dr", "I_df_d", "I_df_u", "F", "I_df_d_ss", "I_df_u_ss", "F_ss")}
    eK = np.exp(-K * lai)
    li = np.arange(m) + 1

    for i in range(nb):
        beta_L, tau_L, rho = leaf_r[i], leaf_t[i], soil_r[i]
        r = np.full(m + 2, 2.0 / 3 * (beta_L / (beta_L + tau_L)) + 1.0 / 3 * (tau_L / (beta_L + tau_L)))  # eq. 23
        t = np.full(m + 2, tau_i)
        a = np.full(m + 2, 1 - (beta_L + tau_L))
        r[0], r[-1] = 1, 0  # ghost soil / top layers (ref :106-108)
        t[0], t[-1] = 0, 1
        a[0], a[-1] = 1 - rho, 0

        pen = t[li] + (1 - t[li]) * (1 - a[li]) * (1 - r[li])  # layer forward "penetration" term
        s_lo = r[li - 1] * (1 - a[li - 1]) * (1 - t[li - 1])  # back-scatter of the layer below
        s_me = r[li] * (1 - a[li]) * (1 - t[li])
        s_hi = r[li + 1] * (1 - a[li + 1]) * (1 - t[li + 1])
        m_lo = 1 - s_lo * s_me  # ref :114 / :137
        m_hi = 1 - s_me * s_hi  # ref :115 / :141

        # tridiagonal in LAPACK banded storage: ab[0]=super, ab[1]=main, ab[2]=sub  (ref :110-122, p. 8)
        ab = np.zeros((3, n))
        ab[1, 0] = 1
        ab[1, -1] = 1
        ab[2, 2 * li - 2] = -pen                # A[2li-1, 2li-2]
        ab[1, 2 * li - 1] = -s_lo * pen          # A[2li-1, 2li-1]   (written -r[li-1]*pen*(1-a)(1-t))
        ab[0, 2 * li] = m_lo                     # A[2li-1, 2li]
        ab[2, 2 * li - 1] = m_hi                 # A[2li,   2li-1]
        ab[1, 2 * li] = -s_hi * pen              # A[2li,   2li]
        ab[0, 2 * li + 1] = -pen                 # A[2li,   2li+1]

        S = I_dr0_all[i] * eK
        r_psi = 0.5 + 0.3334 * ((beta_L - tau_L) / (beta_L + tau_L)) * np.cos(psi)  # eq. 22
        C = np.zeros(n)
        C[0] = rho * S[0]
        C[2 * li - 1] = m_lo * r_psi * (1 - t_psi) * (1 - a[li]) * S
        C[2 * li] = m_hi * (1 - t_psi) * (1 - a[li]) * (1 - r_psi) * S
        C[-1] = I_df0_all[i]

        x = solve_banded((1, 1), ab, C)
        SWu0 = x[::2]
        SWd0 = x[1::2]
        SWd = np.zeros(m + 1)
        SWu = np.zeros(m + 1)
        SWd[li] = SWd0[li] / m_lo + s_me * SWu0[li - 1] / m_lo  # eq. 24 (ref :178-180)
        SWu[li - 1] = SWu0[li - 1] / m_lo + s_lo * SWd0[li] / m_lo  # eq. 25 (ref :183-185)

        out["I_df_d_ss"][:, i] = SWd0[1:]
        out["I_df_d"][:, i] = SWd[1:]
        out["I_df_u_ss"][:, i] = SWu0[:-1]
        out["I_df_u"][:, i] = SWu[:-1]
        out["F_ss"][:, i] = S / mu + 2 * SWu0[:-1] + 2 * SWd0[1:]
        out["F"][:, i] = S / mu + 2 * SWu[:-1] + 2 * SWd[1:]
        out["I_dr"][:, i] = S
    return out


# ----------------------------------------------------------------------------------------------
# zq_pa  Zhao & Qualls, pyAPES variant  (ref solvers/_solve_zq_pa.py:24-418)
# ----------------------------------------------------------------------------------------------
def solve_zq_pa(*, psi, I_dr0_all, I_df0_all, lai, clump, leaf_t, leaf_r, soil_r, K_b_fn):
    """The zq tridiagonal on M = min(100, n_z) equal-LAI layers, interpolated back to `lai`.  Only what
    reaches the reference's return value is restated (its absorption block, ref :363-401, does not)."""
    lai = np.asarray(lai, dtype=float)
    LAI = lai[0]
    N = lai.size
    M = int(np.minimum(100, N))  # ref :94-95
    L = np.ones(M + 2) * LAI / M  # ref :96-100
    L[0] = L[M + 1] = 0.0
    Kb = K_b_fn(psi)
    taud_layer = tau_df_fn(K_b_fn, LAI / M)  # ref :175 (recomputed per band by the reference; band-independent)
    Lcum_full = np.cumsum(np.flipud(L), 0)  # ref :161
    f_sl = np.flipud(np.exp(-Kb * Lcum_full))  # ref :165
    taub = np.exp(-Kb * L)  # ref :173
    taub[0] = 0.0
    nb = I_dr0_all.size
    out = {k: np.zeros((N, nb)) for k in ("I_dr", "I_df_d", "I_df_u", "F")}
    k = np.arange(1, M + 1)
    n = 2 * M + 2
    for ib in range(nb):
        alb = leaf_r[ib] + leaf_t[ib]
        aL = np.ones(M + 2) * (1 - alb)
        tL = np.ones(M + 2) * leaf_t[ib] / alb
        rL = np.ones(M + 2) * leaf_r[ib] / alb
        aL[0], tL[0], rL[0] = 1.0 - soil_r[ib], 0.0, 1.0  # soil (ref :148-150)
        aL[M + 1], tL[M + 1], rL[M + 1] = 0.0, 1.0, 0.0  # transparent atmosphere (ref :153-155)
        taud = np.full(M + 2, taud_layer)
        taud[0] = 0.0  # ref :178-179
        rb = 0.5 + 0.3334 * (rL - tL) / (rL + tL) * np.cos(psi)  # ref :185-186
        rd = 2.0 / 3.0 * rL / (rL + tL) + 1.0 / 3.0 * tL / (rL + tL)
        rb[0] = rd[0] = 1.0
        rb[M + 1] = rd[M + 1] = 0.0
        Ib = f_sl * I_dr0_all[ib]
        pen = taud[k] + (1 - taud[k]) * (1 - aL[k]) * (1 - rd[k])
        s_lo = rd[k - 1] * (1 - aL[k - 1]) * (1 - taud[k - 1])
        s_me = rd[k] * (1 - aL[k]) * (1 - taud[k])
        s_hi = rd[k + 1] * (1 - aL[k + 1]) * (1 - taud[k + 1])
        D_lo = 1 - s_lo * s_me
        D_hi = 1 - s_me * s_hi
        ab = np.zeros((3, n))  # LAPACK banded storage of the matrix of ref :195-236
        ab[1, 0] = ab[1, -1] = 1.0
        ab[2, 2 * k - 2] = -pen
        ab[1, 2 * k - 1] = -s_lo * pen
        ab[0, 2 * k] = D_lo
        ab[2, 2 * k - 1] = D_hi
        ab[1, 2 * k] = -s_hi * pen
        ab[0, 2 * k + 1] = -pen
        C = np.zeros(n)
        C[0] = soil_r[ib] * Ib[0]  # ref :241
        C[2 * k - 1] = D_lo * rb[k] * (1 - taub[k]) * (1 - aL[k]) * Ib[k]  # ref :243-256
        C[2 * k] = D_hi * (1 - taub[k]) * (1 - aL[k]) * (1 - rb[k]) * Ib[k]  # ref :257-271
        C[-1] = I_df0_all[ib]
        SW = solve_banded((1, 1), ab, C)  # ref :278 (dense LAPACK solve there)
        SWu0, SWd0 = SW[0::2], SW[1::2]
        kk = np.arange(M)  # pairs (kk, kk+1): eq. 24 / 25 (ref :286-345)
        D = 1 - rd[kk] * rd[kk + 1] * (1 - aL[kk]) * (1 - taud[kk]) * (1 - aL[kk + 1]) * (1 - taud[kk + 1])
        SWd = np.zeros(M + 1)
        SWu = np.zeros(M + 1)
        SWd[kk + 1] = SWd0[kk + 1] / D + SWu0[kk] * rd[kk + 1] * (1 - aL[kk + 1]) * (1 - taud[kk + 1]) / D
        SWd[0] = SWd[1]
        SWu[kk] = SWu0[kk] / D + SWd0[kk + 1] * rd[kk] * (1 - aL[kk]) * (1 - taud[kk]) / D
        SWu[M] = SWu[M - 1]
        Lcum = np.flipud(Lcum_full[0:M + 1])  # ref :350
        X, xi = np.flipud(lai), np.flipud(Lcum)
        SWdo = np.flipud(np.interp(X, xi, np.flipud(SWd)))  # ref :359-361
        SWuo = np.flipud(np.interp(X, xi, np.flipud(SWu)))
        SWbo = np.exp(-Kb * lai) * I_dr0_all[ib]  # ref :352-353
        out["I_dr"][:, ib] = SWbo
        out["I_df_u"][:, ib] = SWuo
        out["I_df_d"][:, ib] = SWdo
        out["F"][:, ib] = SWbo / np.cos(psi) + 2 * SWuo + 2 * SWdo
    return out


# ----------------------------------------------------------------------------------------------
# 4s  Tian et al. (2007) four-stream  (ref solvers/_solve_4s.py:8-293)
# ----------------------------------------------------------------------------------------------
def fourstream_system(omega, G_int_1, G_int_2, mu_s, P=1.0):
    """The constant 4x4 matrix M of y' = M y + (direct) v exp(-G x/mu0), y = [R2d, R1d, R1u, R2u],
    and the unit forcing pattern; coefficients from Tian eq. 4 as coded in ref :188-203, rows from
    ref `eqns` :80-95."""
    mu_1 = 0.5 * mu_s**2
    mu_2 = 0.5 * (1 - mu_s**2)
    al = 0.5 * omega * P * (1 - mu_s) * G_int_2
    be = 0.5 * omega * P * (1 - mu_s) * G_int_1
    ga = 0.5 * omega * P * (mu_s - 0) * G_int_1
    k1, k2 = G_int_1, G_int_2
    M = np.array(
        [
            [(al - k2) / mu_2, be / mu_2, be / mu_2, al / mu_2],
            [be / mu_1, (ga - k1) / mu_1, ga / mu_1, be / mu_1],
            [-be / mu_1, -ga / mu_1, -(ga - k1) / mu_1, -be / mu_1],
            [-al / mu_2, -be / mu_2, -be / mu_2, -(al - k2) / mu_2],
        ]
    )
    return M, mu_1, mu_2


def solve_4s(
    *, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn, G_fn, mu_s=0.501,
    tol=1e-6, max_nodes=1000, exact_jacobian=False,
):
    """`tol=1e-6, max_nodes=1000, exact_jacobian=False` is the as-shipped configuration (ref :241, :260);
    `tol=1e-11, max_nodes=200000, exact_jacobian=True` is the tight oracle of SURVEY.md section 8c."""
    K = K_b_fn(psi)
    mu = np.cos(psi)
    lai = np.asarray(lai, dtype=float)
    LAI = lai[0]
    G = G_fn(psi)
    G_int_1, G_int_2 = G_sector_integrals(G_fn, mu_s)
    nz, nb = lai.size, I_dr0_all.size
    I_dr_all = np.zeros((nz, nb))
    I_df_d_all = np.zeros((nz, nb))
    I_df_u_all = np.zeros((nz, nb))
    F_all = np.zeros((nz, nb))

    for i in range(nb):
        R_dr0 = I_dr0_all[i] / (np.pi * mu)  # irradiance -> radiance (ref :169-170)
        R_df0 = I_df0_all[i] / np.pi
        rho = soil_r[i]
        omega = leaf_r[i] + leaf_t[i]
        M, mu_1, mu_2 = fourstream_system(omega, G_int_1, G_int_2, mu_s)
        e1 = 0.25 * omega * R_dr0 * (mu_s - 0)  # eps_{p,m}1 (ref :195-198)
        e2 = 0.25 * omega * R_dr0 * (1 - mu_s)
        v = G * np.array([e2 / mu_2, e1 / mu_1, -e1 / mu_1, -e2 / mu_2])

        def make(direct, R0):
            def fun(x, y):
                out = M @ y
                if direct:
                    out = out + v[:, None] * np.exp(-G * x / mu)
                return out

            def bc(ya, yb):
                top = 0.0 if direct else R0
                I_down = 2 * np.pi * (mu_1 * yb[1] + mu_2 * yb[0])
                R_refl = rho / np.pi * (I_down + direct * mu * np.pi * R0 * np.exp(-G * LAI / mu))  # eq. 8b
                return np.array([ya[0] - top, ya[1] - top, yb[2] - R_refl, yb[3] - R_refl])

            jac = None
            bjac = None
            if exact_jacobian:
                def jac(x, y):
                    return np.repeat(M[:, :, None], x.size, axis=2)

                def bjac(ya, yb):
                    dya = np.zeros((4, 4))
                    dyb = np.zeros((4, 4))
                    dya[0, 0] = 1
                    dya[1, 1] = 1
                    for row, col in ((2, 2), (3, 3)):
                        dyb[row, col] = 1
                        dyb[row, 0] = -rho * 2 * mu_2
                        dyb[row, 1] = -rho * 2 * mu_1
                    return dya, dyb

            return fun, bc, jac, bjac

        x0 = np.linspace(0, LAI, 50)
        y0 = np.ones((4, x0.size))
        tot = np.zeros((4, nz))
        for direct, R0 in ((1, R_dr0), (0, R_df0)):
            fun, bc, jac, bjac = make(direct, R0)
            res = integrate.solve_bvp(fun, bc, x0, y0, tol=tol, max_nodes=max_nodes, fun_jac=jac, bc_jac=bjac)
            tot += res.sol(lai)
        I2d = 2 * np.pi * mu_2 * tot[0]
        I1d = 2 * np.pi * mu_1 * tot[1]
        I1u = 2 * np.pi * mu_1 * tot[2]
        I2u = 2 * np.pi * mu_2 * tot[3]
        I_dr = I_dr0_all[i] * np.exp(-K * lai)
        I_dr_all[:, i] = I_dr
        I_df_d_all[:, i] = I1d + I2d
        I_df_u_all[:, i] = I1u + I2u
        F_all[:, i] = I_dr / mu + 2 * (I1u + I2u) + 2 * (I1d + I2d)
    return dict(I_dr=I_dr_all, I_df_d=I_df_d_all, I_df_u=I_df_u_all, F=F_all)


def solve_4s_tight(**kw):
    """Tight-tolerance 4s oracle (SURVEY.md section 8c two-oracle rule)."""
    return solve_4s(tol=1e-11, max_nodes=200000, exact_jacobian=True, **kw)


# ----------------------------------------------------------------------------------------------
# layer absorption and band reduction  (ref model.py:573-647, diagnostics.py:56-81)
# ----------------------------------------------------------------------------------------------
def calc_absorption(*, lai, K_b, leaf_r, leaf_t, I_dr, I_df_d, I_df_u):
    lai = np.asarray(lai, dtype=float)
    dlai = lai[:-1] - lai[1:]
    leaf_a = 1 - (leaf_r + leaf_t)
    laim = (lai[:-1] + lai[1:]) / 2
    f_sl = np.exp(-K_b * laim)
    f_sh = 1 - f_sl
    a = I_dr[1:] - I_dr[:-1] + I_df_d[1:] - I_df_d[:-1] + I_df_u[:-1] - I_df_u[1:]  # ref :606-609
    a_dr = I_dr[1:, :] * (1 - np.exp(-K_b * dlai))[:, None] * leaf_a  # ref :617-621
    a_df = a - a_dr
    a_df_sl = a_df * f_sl[:, None]
    a_df_sh = a_df * f_sh[:, None]
    return {
        "aI": a, "aI_df": a_df, "aI_dr": a_dr, "aI_sh": a_df_sh, "aI_sl": a_df_sl + a_dr,
        "aI_df_sl": a_df_sl, "aI_df_sh": a_df_sh, "laim": laim, "f_slm": f_sl,
    }


def x_frac_in_bounds(xe, bounds):
    """Loop restatement of the band-weight rule (ref spectra.py:71-126)."""
    xe = np.asarray(xe, dtype=float)
    w = np.zeros(xe.size - 1)
    b1, b2 = bounds
    for i in range(w.size):
        lo, hi = xe[i], xe[i + 1]
        if not (hi >= b1 and lo <= b2):
            continue
        if lo < b1:
            w[i] = (hi - b1) / (hi - lo)
        elif hi > b2:
            w[i] = (b2 - lo) / (hi - lo)
        else:
            w[i] = 1
    return w


def canopy_absorbed_bands(aI, wle, bands=((0.4, 0.7), (0.7, 2.5))):
    """Sum over layers and weighted Sum over wavelength: canopy-integrated absorbed PAR, NIR (W m-2)."""
    return np.array([(aI * x_frac_in_bounds(wle, b)).sum() for b in bands])


SOLVERS = {
    "2s": solve_2s, "4s": solve_4s, "bf": solve_bf, "bl": solve_bl, "g77": solve_g77,
    "n79": solve_n79, "zq": solve_zq, "zq_pa": solve_zq_pa,
}
ARGS = {
    "2s": ["psi", "I_dr0_all", "I_df0_all", "lai", "leaf_t", "leaf_r", "soil_r", "K_b_fn", "G_fn", "mla"],
    "4s": ["psi", "I_dr0_all", "I_df0_all", "lai", "leaf_t", "leaf_r", "soil_r", "K_b_fn", "G_fn"],
    "zq": ["psi", "I_dr0_all", "I_df0_all", "lai", "leaf_t", "leaf_r", "soil_r", "K_b_fn", "G_fn"],
    "bl": ["psi", "I_dr0_all", "I_df0_all", "lai", "leaf_t", "leaf_r", "K_b_fn"],
    "bf": ["psi", "I_dr0_all", "I_df0_all", "lai", "leaf_t", "leaf_r", "soil_r", "K_b_fn"],
    "g77": ["psi", "I_dr0_all", "I_df0_all", "lai", "leaf_t", "leaf_r", "soil_r", "K_b_fn"],
    "n79": ["psi", "I_dr0_all", "I_df0_all", "lai", "leaf_t", "leaf_r", "soil_r", "K_b_fn"],
    "zq_pa": ["psi", "I_dr0_all", "I_df0_all", "lai", "clump", "leaf_t", "leaf_r", "soil_r", "K_b_fn"],
}


def run(scheme, p, **extra):
    """Call an oracle solver from a full parameter dict (as `Model.run` does, ref model.py:305-310)."""
    return SOLVERS[scheme](**{k: p[k] for k in ARGS[scheme]}, **extra)


def smear_tuv(x, y, bins):
    """TUV / F0AM bin averaging: trapezoidally integrated average of y(x) in every bin
    (ref crt1d/spectra.py:221-300, `_smear_tuv_1` + `smear_tuv`).  Restated per bin with the reference's
    trapezoid loop (clipped trapezoids visited in increasing k, `area += (a2 - a1) (b2 + b1) / 2`), vectorised
    over the rows of a 2-D `y`."""
    x = np.asarray(x, dtype=np.float64)
    y2 = np.atleast_2d(np.asarray(y, dtype=np.float64))
    bins = np.asarray(bins, dtype=np.float64)
    out = np.zeros((y2.shape[0], bins.size - 1))
    for i, (xl, xu) in enumerate(zip(bins[:-1], bins[1:])):
        area = np.zeros(y2.shape[0])
        for k in range(x.size - 1):
            if x[k + 1] < xl:
                continue
            if x[k] > xu:
                break
            a1 = max(x[k], xl)
            a2 = min(x[k + 1], xu)
            slope = (y2[:, k + 1] - y2[:, k]) / (x[k + 1] - x[k])
            b1 = y2[:, k] + slope * (a1 - x[k])
            b2 = y2[:, k] + slope * (a2 - x[k])
            area = area + (a2 - a1) * (b2 + b1) / 2
        out[:, i] = area / (xu - xl)
    return out[0] if np.ndim(y) == 1 else out

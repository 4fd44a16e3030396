"""Shared helpers of the test suite."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fp64 parity bar of BASELINE.json / SURVEY.md section 8d: max rel err <= 1e-10 on every returned array
RTOL = 1e-10
# 4s two-oracle rule: <= 1e-9 vs the tight-tolerance reference,
# <= 2e-4 relative or 1e-7 W m-2 absolute vs the as-shipped (tol = 1e-6) reference
RTOL_4S_TIGHT = 1e-9
RTOL_4S_SHIPPED = 2e-4
ATOL_4S_SHIPPED = 1e-7
# The tight oracle is itself a collocation solution (solve_bvp tol = 1e-11 on residuals scaled by
# 1 + |f|): its ABSOLUTE accuracy floor is ~1e-13 of the field's scale, which in the weakest bands
# (irradiance ~1e-5 of the strongest) is just over 1e-9 relative.  So: 1e-9 relative, or within
# 1e-12 x max|field| absolute -- never looser than the oracle's own resolution.
ATOL_4S_TIGHT_FRAC = 1e-12


def golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def relerr(x, ref):
    """max |x - ref| / max(|ref|, tiny), element-wise (cfg 2)."""
    x, ref = np.asarray(x, dtype=float), np.asarray(ref, dtype=float)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    if x.size == 0:
        return 0.0
    return float(np.max(np.abs(x - ref) / np.maximum(np.abs(ref), 1e-300)))


def assert_close(x, ref, rtol, what="", atol=0.0):
    x, ref = np.asarray(x, dtype=float), np.asarray(ref, dtype=float)
    assert x.shape == ref.shape, (what, x.shape, ref.shape)
    assert np.all(np.isfinite(x)) or not np.all(np.isfinite(ref)), f"{what}: non-finite values"
    err = np.abs(x - ref)
    ok = err <= np.maximum(rtol * np.abs(ref), atol)
    if not np.all(ok):
        i = np.unravel_index(np.argmax(err / np.maximum(np.abs(ref), 1e-300)), err.shape)
        raise AssertionError(f"{what}: max rel err {relerr(x, ref):.3e} > {rtol:g} at {i}: got {x[i]!r}, ref {ref[i]!r}")


def assert_close_4s(x, ref, what=""):
    """4s vs the tight-tolerance reference (see ATOL_4S_TIGHT_FRAC)."""
    assert_close(x, ref, RTOL_4S_TIGHT, what, atol=ATOL_4S_TIGHT_FRAC * float(np.max(np.abs(ref))))


def with_callables(p):
    """Add K_b_fn / G / K_b to a case dict, as Model._check_inputs does (ref model.py:291-293)."""
    q = dict(p)
    G_fn = q["G_fn"]
    q["K_b_fn"] = lambda psi_: G_fn(psi_) / np.cos(psi_)
    q["G"] = G_fn(q["psi"])
    q["K_b"] = q["K_b_fn"](q["psi"])
    return q


def variant_case(nz, sza_deg):
    """The inputs of a `ref_variants.npz` entry: default case at nz levels, sza, every 9th band."""
    from crt1d_b200 import cases

    q = dict(cases.load_default_case(nz))
    q["psi"] = np.deg2rad(sza_deg)
    for k in ("leaf_t", "leaf_r", "soil_r", "I_dr0_all", "I_df0_all", "wl", "dwl", "wl_leafsoil"):
        q[k] = q[k][::9].copy()
    return with_callables(q)


VARIANTS = [(2, 20), (3, 20), (10, 60), (200, 20), (60, 0), (60, 60), (60, 85)]


def sigma_rel_2s(batch, mu_bar):
    """Conditioning of the Sellers two-stream formulas, per (scenario, band): |sigma| / (mu_bar K)^2 with
    sigma = (mu_bar K)^2 + c^2 - b^2 (ref _solve_2s.py:85).  sigma -> 0 (direct-beam extinction K equal to
    the diffuse eigenvalue h) is a removable singularity of the closed form: h1/sigma and h4/sigma blow up
    and cancel against the h2, h3 / h5, h6 terms, so ANY float64 evaluation -- the reference's included --
    loses ~eps/sigma_rel relative accuracy there (measured: reference 1.8e-11, this kernel 1.2e-11 off the
    50-digit value at sigma_rel = 4e-5)."""
    la = batch.leaf_angle
    K = np.array([la.K_b_fn(p) for p in batch.psi])[:, None]
    r = batch.leaf_r_lib[batch.leaf_idx]
    t = batch.leaf_t_lib[batch.leaf_idx]
    cos2 = np.cos(np.radians(batch.mla)) ** 2
    omega = r + t
    beta = 0.5 * (r + t + (r - t) * cos2) / omega
    b = 1 - (1 - beta) * omega
    c = omega * beta
    mk2 = (mu_bar * K) ** 2
    return np.abs(mk2 + c * c - b * b) / mk2


def assert_close_conditioned(x, ref, base_rtol, sigma_rel, what=""):
    """Per-column tolerance max(base_rtol, 64 eps / sigma_rel): x, ref are (S, n_z, n_wl), sigma_rel (S, n_wl).
    (64 = a few ulp from each side's exponentials and reciprocals, incl. the <= LV-ulp drift of the level
    recurrence, times the ~1/sigma_rel amplification; it only matters for sigma_rel < 1.4e-4.)"""
    tol = np.maximum(base_rtol, 64 * 2.220446049250313e-16 / sigma_rel)[:, None, :]
    x, ref = np.asarray(x), np.asarray(ref)
    err = np.abs(x - ref) / np.maximum(np.abs(ref), 1e-300)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err / tol), err.shape)
        raise AssertionError(f"{what}: rel err {err[i]:.3e} > tol {tol[i[0], 0, i[2]]:.3e} at {i} "
                             f"(sigma_rel {sigma_rel[i[0], i[2]]:.2e}); {bad.sum()} elements out of tolerance")


def deep_case(n_bands=6):
    """BASELINE.json configs[4]: deep canopy, n_z = 1000, LAI = 6 (SURVEY.md section 8d cfg 5)."""
    from crt1d_b200 import cases

    q = dict(cases.load_default_case(1000))
    q["lai"] = np.linspace(1, 0, 1000) * 6.0
    step = 107 // n_bands
    for k in ("leaf_t", "leaf_r", "soil_r", "I_dr0_all", "I_df0_all", "wl", "dwl", "wl_leafsoil"):
        q[k] = q[k][::step][:n_bands].copy()
    return with_callables(q)


def ragged_case(nz, kind="cluster", n_bands=8):
    """Default case on `nz` levels with a NON-uniform cumulative-LAI axis, for the checkpointed tridiagonal
    sweeps (segment boundaries at every multiple of the checkpoint spacing) and zq_pa's streamed
    interpolation: `cluster` packs the levels near the canopy top (several caller levels inside one M-grid
    interval, empty intervals lower down), `quad` is the reference's own beta-like shape."""
    from crt1d_b200 import cases

    q = dict(cases.load_default_case(nz))
    x = np.linspace(1, 0, nz)
    q["lai"] = (x ** 4 if kind == "cluster" else x ** 2) * 4.0
    q["psi"] = np.deg2rad(35.0)
    step = max(1, 107 // n_bands)
    for k in ("leaf_t", "leaf_r", "soil_r", "I_dr0_all", "I_df0_all", "wl", "dwl", "wl_leafsoil"):
        q[k] = q[k][::step][:n_bands].copy()
    return with_callables(q)


RAGGED_NZ = (7, 8, 9, 10, 11, 16, 17, 20, 21, 30, 31, 99, 100, 101, 110)

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


# ---- 4s edge cases: omega -> 1 (vanishing / negative eigenvalue^2 of the 4-stream operator) and
# ---- kappa = lambda_k resonances (ref _solve_4s.py:48-97 integrates the ODE and is indifferent to both)
EDGE_4S_PSI_DEG = (10.0, 45.0, 75.0)
EDGE_4S_MU_S = (0.501, 0.33998)
EDGE_4S_OMEGA = (0.99, 0.995, 0.9954, 0.997, 0.9999, 1.0 - 1e-9)


def fourstream_N(omega, G1, G2, mu_s):
    """The 2x2 matrix N = (P - Q)(P + Q) of s'' = N s + ... for s = D + U (sum of the downward and upward
    radiance pairs of ref `eqns`, _solve_4s.py:80-95), index 0 <-> sector 2, index 1 <-> sector 1."""
    m1, m2 = 0.5 * mu_s**2, 0.5 * (1 - mu_s**2)
    al = 0.5 * omega * (1 - mu_s) * G2
    be = 0.5 * omega * (1 - mu_s) * G1
    ga = 0.5 * omega * mu_s * G1
    q0, q1 = G2 / m2, G1 / m1
    T = np.array([[(2 * al - G2) / m2, 2 * be / m2], [2 * be / m1, (2 * ga - G1) / m1]])
    return -np.diag([q0, q1]) @ T


def fourstream_l2(omega, G1, G2, mu_s):
    """Eigenvalues lambda^2 of N, descending."""
    return np.sort(np.linalg.eigvals(fourstream_N(omega, G1, G2, mu_s)).real)[::-1]


def edge_4s_case(psi_deg, mu_s, n_z=20, lai_tot=4.0):
    """Inputs of one `ref_4s_edge.npz` entry: the default canopy's leaf-angle function at `psi_deg`, `n_z`
    equally spaced levels, and one band per entry of: EDGE_4S_OMEGA; omega* where the smaller eigenvalue^2
    crosses zero (and omega* (1 +- 1e-7)); the omega at which kappa = K_b equals lambda_0 or lambda_1 when
    such an omega exists in (0.02, 0.999) (and that omega (1 + 1e-9), (1 + 1e-5)).  r = 0.55 omega,
    t = 0.45 omega."""
    from scipy.optimize import brentq
    import scipy.integrate as integ

    from crt1d_b200 import cases

    q = dict(cases.load_default_case(n_z))
    G_fn = q["G_fn"]
    G1 = integ.quad(lambda m: G_fn(np.arccos(m)), 0, mu_s)[0]
    G2 = integ.quad(lambda m: G_fn(np.arccos(m)), mu_s, 1)[0]
    psi = np.deg2rad(psi_deg)
    kap = G_fn(psi) / np.cos(psi)
    om = list(EDGE_4S_OMEGA)
    f1 = lambda w: fourstream_l2(w, G1, G2, mu_s)[1]  # noqa: E731
    if f1(0.9) * f1(1.0) < 0:
        ws = brentq(f1, 0.9, 1.0, xtol=1e-15, rtol=1e-15)
        om += [ws * (1 - 1e-7), ws, min(1.0, ws * (1 + 1e-7))]
    for k in (0, 1):
        fk = lambda w, k=k: fourstream_l2(w, G1, G2, mu_s)[k] - kap * kap  # noqa: E731
        if fk(0.02) * fk(0.999) < 0:
            wr = brentq(fk, 0.02, 0.999, xtol=1e-15, rtol=1e-15)
            om += [wr, wr * (1 + 1e-9), wr * (1 + 1e-5)]
    om = np.array(om)
    n = om.size
    q["psi"] = psi
    q["lai"] = np.linspace(1, 0, n_z) * lai_tot
    q["leaf_r"] = 0.55 * om
    q["leaf_t"] = 0.45 * om
    q["soil_r"] = np.linspace(0.05, 0.3, n)
    q["I_dr0_all"] = np.linspace(80.0, 120.0, n)
    q["I_df0_all"] = np.linspace(40.0, 20.0, n)
    q["wl"] = np.linspace(0.4, 2.4, n)
    q["dwl"] = np.full(n, 2.0 / max(n - 1, 1))
    q["wl_leafsoil"] = q["wl"]
    return with_callables(q)


def edge_4s_fixture_cases():
    """(tag, mu_s, params, shipped, tight) for every entry of ref_4s_edge.npz, inputs taken from the fixture itself
    (the GPU box needs neither scipy.optimize nor the reference)."""
    from crt1d_b200 import cases

    g = golden("ref_4s_edge.npz")
    base = cases.load_default_case(20)
    out = []
    for psi_deg in EDGE_4S_PSI_DEG:
        for mu_s in EDGE_4S_MU_S:
            tag = f"psi{int(psi_deg)}_mus{int(round(mu_s * 1000))}"
            q = dict(base)
            for k in ("psi", "lai", "leaf_r", "leaf_t", "soil_r", "I_dr0_all", "I_df0_all"):
                q[k] = g[f"{tag}__in__{k}"]
            q["psi"] = float(q["psi"])
            n = q["leaf_r"].size
            q["wl"] = np.linspace(0.4, 2.4, n)
            q["dwl"] = np.full(n, 2.0 / max(n - 1, 1))
            q["wl_leafsoil"] = q["wl"]
            q = with_callables(q)
            ship = {k: g[f"{tag}__shipped__{k}"] for k in ("I_dr", "I_df_d", "I_df_u", "F")}
            tight = {k: g[f"{tag}__tight__{k}"] for k in ("I_dr", "I_df_d", "I_df_u", "F")}
            out.append((tag, mu_s, q, ship, tight))
    return out


def assert_close_4s_shipped(x, ref, what=""):
    """4s vs the reference as shipped (solve_bvp tol = 1e-6): 2e-4 relative or 1e-7 W m-2."""
    assert_close(x, ref, RTOL_4S_SHIPPED, what, atol=ATOL_4S_SHIPPED)


def tune(monkeypatch, name, value):
    """Set (or, with None, delete) a CRT1D_B200_* kernel-selection variable for this test and make the library
    re-read them: they are read once at load time, not on the launch path (conftest restores them afterwards)."""
    from crt1d_b200 import _lib

    if value is None:
        monkeypatch.delenv(name, raising=False)
    else:
        monkeypatch.setenv(name, value)
    _lib.load().crt1d_reload_tuning()


def random_4s_batch(seed=7):
    """Seeded random 4s scenarios for `ref_4s_random.npz`: zenith angles to 86 deg, omega = r + t up to 1 (a third of
    the bands above 0.99), soil albedo to 0.5, one equally spaced and one irregular LAI profile, thin and deep."""
    from crt1d_b200.leaf_angle import LeafAngle
    from crt1d_b200.scenarios import ScenarioBatch

    rng = np.random.default_rng(seed)
    S, nw, nz = 5, 8, 17
    om = np.concatenate([rng.uniform(0.02, 0.99, (3, nw - 3)), rng.uniform(0.99, 1.0, (3, 3))], axis=1)
    frac = rng.uniform(0.3, 0.7, (3, nw))
    steps = rng.uniform(0.2, 1.8, nz - 1)
    irr = np.concatenate([np.cumsum(steps[::-1])[::-1], [0.0]])
    lai_lib = np.stack([np.linspace(1, 0, nz) * 5.5, irr / irr[0] * 2.2, np.linspace(1, 0, nz) * 0.3])
    return ScenarioBatch(
        psi=np.radians([3.0, 30.0, 55.0, 72.0, 86.0]), lai_lib=lai_lib, leaf_r_lib=om * frac, leaf_t_lib=om * (1 - frac),
        soil_r_lib=rng.uniform(0.02, 0.5, (2, nw)), I_dr0_lib=rng.uniform(0.0, 9.0, (2, nw)),
        I_df0_lib=rng.uniform(0.1, 5.0, (2, nw)), lai_idx=[0, 1, 2, 0, 1], leaf_idx=[0, 1, 2, 2, 0],
        soil_idx=[0, 1, 0, 1, 0], sky_idx=[0, 1, 1, 0, 0], leaf_angle=LeafAngle(), mla=57.0,
    )


# ---------------------------------------------------------------------------------------------------------------
# Non-beta LAI generators and the non-uniform cumulative-LAI axes they produce
BORDEN95_CDD = dict(lai_tot=3.044, lai_frac=[0.608, 0.392], h_canopy=22.0, h_max_lad=[15.4, 6.16], h_bot=[12.1, 1.375],
                    h_top=[22.0, 12.0], lad_h_top=[0, 0.065])
LEAF_AREA_CASES = [
    # (tag, generator name, positional args, keyword args); `z20` / `z60` grids are rebuilt by leaf_area_args()
    ("wz_pine20", "distribute_lai_weibull_z", (np.linspace(0, 10.5, 20), 5, 10), dict(hb=2, species="pine")),
    ("wz_spruce20", "distribute_lai_weibull_z", (np.linspace(0, 10.5, 20), 5, 10), dict(hb=2, species="spruce")),
    ("wz_birch60", "distribute_lai_weibull_z", (np.linspace(0, 20.5, 60), 4, 20), dict(hb=0.5, species="birch")),
    ("wz_bc60", "distribute_lai_weibull_z", (np.linspace(0, 20.5, 60), 6.5, 20), dict(hb=3.0, b=1.2, c=2.4)),
    ("w_pine20", "distribute_lai_weibull", (10, 5, 20), dict(h_min=2, species="pine")),
    ("w_spruce60", "distribute_lai_weibull", (20, 4, 60), dict(species="spruce")),
    ("w_birch1000", "distribute_lai_weibull", (30, 6.5, 1000), dict(h_min=3, species="birch")),
    ("gamma10", "distribute_lai_gamma", (20, 4, 10), {}),
    ("gamma60", "distribute_lai_gamma", (20, 4, 60), {}),
    ("gamma1000", "distribute_lai_gamma", (35, 7.5, 1000), {}),
    ("cdd20", "distribute_lai_from_cdd", (BORDEN95_CDD, 20), {}),
    ("cdd60", "distribute_lai_from_cdd", (BORDEN95_CDD, 60), {}),
    ("cdd61", "distribute_lai_from_cdd", (BORDEN95_CDD, 61), {}),
]
NONUNIFORM_AXES = ("wz_birch60", "wz_pine20", "gamma60", "gamma10")


def nonuniform_case(tag):
    """Default case (every 9th band, SZA 35 deg) on the cumulative-LAI axis of LEAF_AREA_CASES[tag], generated by the
    PRODUCT's leaf_area module (bit-identical to the reference's, see test_leaf_area_generators)."""
    from crt1d_b200 import leaf_area

    _, fn, args, kw = next(c for c in LEAF_AREA_CASES if c[0] == tag)
    prof = getattr(leaf_area, fn)(*args, **kw)
    q = variant_case(len(prof.lai), 35)
    q["lai"], q["z"] = np.asarray(prof.lai, dtype=np.float64), np.asarray(prof.z, dtype=np.float64)
    return with_callables(q)


def assert_close_same_nans(x, ref, rtol, what="", atol=0.0):
    """`assert_close` where the reference itself returns NaN / inf (n79 sunlit/shaded absorption per leaf area on
    zero-thickness layers: 0/0 and x/0 in ref _solve_n79.py): the non-finite values must be the same ones at the same
    places, everything else must agree."""
    x, ref = np.asarray(x, dtype=float), np.asarray(ref, dtype=float)
    assert x.shape == ref.shape, (what, x.shape, ref.shape)
    assert np.array_equal(np.isnan(x), np.isnan(ref)), f"{what}: NaN pattern differs from the reference's"
    inf = np.isinf(ref)
    assert np.array_equal(np.isinf(x), inf) and np.array_equal(x[inf], ref[inf]), f"{what}: inf pattern differs"
    m = np.isfinite(ref)
    assert_close(x[m], ref[m], rtol, what, atol=atol)


def zq_pa_adversarial_case(nz, sza_deg, lai_tot, seed):
    """Default case with leaf / soil optics that stress zq_pa's closed M-grid solution: omega from 1e-6 to 1 - 1e-12
    (degenerate eigenvalue of the layer transfer matrix), bands placed ON the beam / diffuse-mode resonance
    lam taub = 1 and next to it, black and bright soil, zero direct beam."""
    from crt1d_b200 import cases
    rng = np.random.default_rng(seed)
    q = dict(cases.load_default_case(nz))
    n = 96
    om = np.concatenate([10.0 ** rng.uniform(-6, -1, 16), rng.uniform(0.02, 0.98, 40), 1.0 - 10.0 ** rng.uniform(-12, -2, 24),
                         np.full(16, 0.5)])
    fr = rng.uniform(0.05, 0.95, n)
    q["leaf_r"], q["leaf_t"] = om * fr, om * (1 - fr)
    q["soil_r"] = rng.choice([0.0, 0.1, 0.3, 0.95], n)  # incl. a perfectly black soil (zero pivot of a pivot-free sweep)
    q["I_dr0_all"] = rng.choice([0.0, 1e-8, 0.7, 1.3], n)
    q["I_df0_all"] = rng.uniform(1e-6, 1.0, n)
    q["wl"] = np.linspace(0.4, 2.5, n)
    q["dwl"] = np.full(n, q["wl"][1] - q["wl"][0])
    q["wl_leafsoil"] = q["wl"]
    q["psi"] = np.deg2rad(sza_deg)
    q["lai"] = np.linspace(1, 0, nz) * lai_tot
    q = with_callables(q)
    # the last 16 bands: solve lam(omega) taub = 1 for omega by bisection (same formulas as the kernel), then detune
    M = min(100, nz)
    dl = lai_tot / M
    from crt1d_b200.solvers import common

    taud = common.tau_df_fn(q["K_b_fn"], dl)
    taub = np.exp(-q["K_b_fn"](q["psi"]) * dl)

    def lam(o, f):
        rL, tL = f, 1 - f
        rd = 2 / 3 * rL + 1 / 3 * tL
        pen = taud + (1 - taud) * o * (1 - rd)
        sc = rd * o * (1 - taud)
        hm1 = ((1 - taud) * (1 - o)) * ((1 - taud) * (1 - o) + 2 * sc) / (2 * pen)
        return 1 + hm1 + np.sqrt(hm1 * (hm1 + 2))

    for i, det in enumerate((0.0, 1e-9, -1e-7, 1e-6, -1e-5, 5e-5, -9e-5, 1.1e-4, -2e-4, 1e-3, -1e-3, 1e-2, 3e-9, -3e-8, 2e-4, -5e-4)):
        f = fr[80 + i]
        lo, hi = 1e-9, 1.0  # lam decreases with omega; resonance needs lam = 1/taub
        if not (lam(hi, f) < 1 / taub < lam(lo, f)):
            continue
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            lo, hi = (mid, hi) if lam(mid, f) > 1 / taub else (lo, mid)
        o = min(1.0, max(1e-9, 0.5 * (lo + hi) * (1 + det)))
        q["leaf_r"][80 + i], q["leaf_t"][80 + i] = o * f, o * (1 - f)
    return q

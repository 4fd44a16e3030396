"""CPU tier: the arithmetic the CUDA kernels execute (crt1d_b200/csrc/*.cuh, compiled for the host by
tests/_hostcheck) vs the oracle and the reference-generated goldens.  Catches math regressions in the
kernels in a container without a GPU; the real parity tests are tests/test_gpu_parity.py."""
import numpy as np
import pytest

import crt_oracle as oracle
import hostcheck
from crt1d_b200.engine import host_prologue
from crt1d_b200.scenarios import ScenarioBatch
from util import RTOL
from util import RTOL_4S_TIGHT
from util import VARIANTS
from util import assert_close
from util import assert_close_4s
from util import assert_close_4s_shipped
from util import edge_4s_case
from util import edge_4s_fixture_cases
from util import deep_case as _deep_case
from util import golden
from util import variant_case
from util import zq_pa_adversarial_case as _zq_pa_adversarial_case

SCHEMES = ("2s", "bf", "bl", "g77", "n79", "zq", "zq_pa")


def _solve(p, scheme, vec=None, **kw):
    batch = ScenarioBatch.from_params(p)
    pro = host_prologue(batch, scheme, K_b_fn=p["K_b_fn"], G_fn=p["G_fn"], **kw)
    out = hostcheck.solve(batch, scheme, pro, vec=vec, mu_s=kw.get("mu_s", 0.501))
    sol = {k: v[0] for k, v in out.items()}
    if "rho_c" in sol:
        sol["rho_c"] = sol["rho_c"][-1]
    return sol


@pytest.mark.parametrize("scheme", SCHEMES)
def test_kernel_math_default_case(scheme, default_p):
    ref = golden(f"ref_default_{scheme}.npz")
    sol = _solve(default_p, scheme, vec=1)
    for k in ref:
        assert_close(sol[k], ref[k], RTOL, f"{scheme}.{k}")


@pytest.mark.parametrize("scheme", SCHEMES + ("4s",))
def test_vec2_equals_vec1(scheme, default_p):
    """The 2-bands-per-thread instantiation computes the same columns (even band count needed)."""
    sub = {k: (v[:106].copy() if isinstance(v, np.ndarray) and v.shape == (107,) else v) for k, v in default_p.items()}
    a = _solve(sub, scheme, vec=1)
    b = _solve(sub, scheme, vec=2)
    for k in a:
        assert np.array_equal(a[k], b[k]), (scheme, k)


def test_kernel_math_4s(default_p):
    for mu_s, tag in ((0.501, "4s"), (0.33998, "4s_mus034")):
        ref = golden(f"ref_default_{tag}_tight.npz")
        sol = _solve(default_p, "4s", vec=1, mu_s=mu_s)
        for k in ref:
            assert_close_4s(sol[k], ref[k], f"4s[{mu_s}].{k}")


@pytest.mark.parametrize("nz,sza", VARIANTS)
def test_kernel_math_variants(nz, sza):
    g = golden("ref_variants.npz")
    q = variant_case(nz, sza)
    tag = f"nz{nz}_sza{sza}"
    for scheme in SCHEMES + ("4s_tight",):
        name = "4s" if scheme == "4s_tight" else scheme
        keys = [k for k in g if k.startswith(f"{tag}__{scheme}__") and not k.endswith("__raises")]
        if not keys:
            continue
        sol = _solve(q, name)
        for full in keys:
            k = full.split("__")[-1]
            if name == "4s":
                assert_close_4s(sol[k], g[full], f"{tag} {scheme}.{k}")
            else:
                assert_close(sol[k], g[full], RTOL, f"{tag} {scheme}.{k}")


@pytest.mark.parametrize("kind", ["cluster", "quad"])
def test_kernel_math_checkpointed_sweeps_ragged_levels(kind):
    """Every level count around the checkpoint spacing (8 here, 10 on the device) and around zq_pa's
    100-layer cap, on non-uniform level axes."""
    from util import RAGGED_NZ, ragged_case

    for nz in RAGGED_NZ:
        q = ragged_case(nz, kind)
        for scheme in ("zq", "zq_pa", "n79"):
            kw = {"tau_d_method": "9sky"} if scheme == "n79" else {}
            ref = oracle.run(scheme, q, **kw)
            for vec in (1, 2):
                sol = _solve(q, scheme, vec=vec, **kw)
                for k in ref:
                    assert_close(sol[k], ref[k], RTOL, f"ragged {kind} nz={nz} {scheme}.{k} vec{vec}")


def test_kernel_math_multi_scenario_indexing(default_p):
    """Library/index plumbing: scenarios pick different rows; each must equal its own single solve."""
    rng = np.random.default_rng(3)
    nz = 12
    lai_lib = np.stack([np.linspace(1, 0, nz) * 2.0, np.linspace(1, 0, nz) ** 2 * 6.0])
    sc = rng.uniform(0.7, 1.1, (3, 1))
    b = ScenarioBatch(
        psi=np.radians([5.0, 40.0, 70.0, 84.0, 20.0]), lai_lib=lai_lib, leaf_r_lib=default_p["leaf_r"] * sc,
        leaf_t_lib=default_p["leaf_t"] * sc[::-1], soil_r_lib=np.stack([default_p["soil_r"], default_p["soil_r"] * 2]),
        I_dr0_lib=np.stack([default_p["I_dr0_all"], default_p["I_dr0_all"] * 0.3]),
        I_df0_lib=np.stack([default_p["I_df0_all"], default_p["I_df0_all"] * 1.9]),
        lai_idx=[0, 1, 1, 0, 1], leaf_idx=[0, 1, 2, 1, 0], soil_idx=[0, 1, 0, 1, 1], sky_idx=[0, 0, 1, 1, 0],
        leaf_angle=default_p["leaf_angle"], mla=57.0, wl=default_p["wl"], dwl=default_p["dwl"],
    )
    for scheme in SCHEMES + ("4s",):
        pro = host_prologue(b, scheme)
        out = hostcheck.solve(b, scheme, pro, band_w=np.ones((2, b.n_wl)))
        for s in range(b.n_scen):
            q = b.scenario_params(s)
            ref = (oracle.solve_4s_tight if scheme == "4s" else oracle.SOLVERS[scheme])(**{k: q[k] for k in oracle.ARGS[scheme]})
            for k in ("I_dr", "I_df_d", "I_df_u", "F"):
                if scheme == "4s":
                    assert_close_4s(out[k][s], ref[k], f"{scheme}[{s}].{k}")
                else:
                    assert_close(out[k][s], ref[k], RTOL, f"{scheme}[{s}].{k}")
            ab = oracle.calc_absorption(lai=q["lai"], K_b=q["K_b"], leaf_r=q["leaf_r"], leaf_t=q["leaf_t"],
                                        I_dr=ref["I_dr"], I_df_d=ref["I_df_d"], I_df_u=ref["I_df_u"])
            assert_close(out["absorbed"][s, 0], ab["aI"].sum(), 1e-9, f"{scheme}[{s}] absorbed")


def test_zq_pa_closed_form_vs_oracle_and_thomas():
    """zq_pa's closed M-grid solution (crt_core.cuh zq_pa_closed_coef: constant-coefficient layer transfer matrix, two
    modes + beam term) against the oracle (the reference's dense solve) AND against the same kernel source built with
    the closed form disabled (checkpointed Thomas sweep on every column), on optics that hit both fall-back windows
    (degenerate eigenvalue, beam resonance) and their edges."""
    T = hostcheck.build_variant("zqpa_thomas", ["CRT_ZQPA_NO_CLOSED"])
    n_closed = 0
    for nz, sza, lai_tot, seed in ((60, 30.0, 4.0, 0), (10, 65.0, 7.5, 1), (150, 5.0, 1.0, 2), (4, 80.0, 3.0, 3), (33, 50.0, 0.2, 4)):
        q = _zq_pa_adversarial_case(nz, sza, lai_tot, seed)
        ref = oracle.run("zq_pa", q)
        a = _solve(q, "zq_pa")
        with hostcheck.use_lib(T):
            b = _solve(q, "zq_pa")
        for k in ("I_dr", "I_df_d", "I_df_u", "F"):
            atol = 1e-14 * np.max(np.abs(ref[k]), axis=0, keepdims=True)  # the dense solve's own absolute floor per column
            assert_close(b[k], ref[k], RTOL, f"thomas zq_pa nz={nz}.{k}", atol=atol)
            assert_close(a[k], ref[k], RTOL, f"closed zq_pa nz={nz}.{k}", atol=atol)
        n_closed += int(np.sum(np.any(a["I_df_u"] != b["I_df_u"], axis=0)))
    assert n_closed > 300  # most of the 480 columns take the closed form (bit-different from the Thomas sweep)


@pytest.mark.parametrize("rho", [0.0, 1e-300, 1e-30])
def test_black_soil_zq_family(rho, default_p):
    """A perfectly black (or numerically black) soil zeroes the main-diagonal entry of row 1 of the zq / zq_pa system --
    a pivot of the pivot-free Thomas sweep (NaN/Inf before the pivot floor), while the reference's pivoting solvers
    (`spsolve`, `np.linalg.solve`: _solve_zq.py:147, _solve_zq_pa.py:278) return finite values.  Both builds (closed
    M-grid solution and Thomas sweep) against the oracle."""
    T = hostcheck.build_variant("zqpa_thomas", ["CRT_ZQPA_NO_CLOSED"])
    q = dict(default_p)
    q["soil_r"] = np.full_like(default_p["soil_r"], rho)
    for scheme in ("zq", "zq_pa"):
        ref = oracle.run(scheme, q)
        a = _solve(q, scheme)
        with hostcheck.use_lib(T):
            b = _solve(q, scheme)
        for k in ref:
            atol = 1e-14 * np.max(np.abs(ref[k]), axis=0, keepdims=True)  # I_df_u at the ground is ~rho: the reference solve's absolute floor
            assert_close(a[k], ref[k], RTOL, f"black soil {scheme}.{k}", atol=atol)
            assert_close(b[k], ref[k], RTOL, f"black soil thomas {scheme}.{k}", atol=atol)


def test_kernel_math_extreme_inputs_fuzz():
    """Seeded fuzz over inputs no default case reaches: omega from 1e-6 to 1 - 1e-12, soil_r from 1e-12 to 1, zero / tiny /
    huge sky irradiances, sun at the zenith and 0.1 degree above the horizon, LAI from 0.01 to 20, 3 to 130 levels on linear,
    quadratic and clustered axes.  (Known limit, outside this range: for VANISHING layers -- total LAI 1e-6 on 130 quartic
    levels, layer LAI ~1e-8 -- the pivot-free zq sweep is 1.4e-8 off where the reference's pivoting solver is exact to 1e-12.)  Wherever the reference (oracle) is finite the kernel arithmetic must be finite and
    within the bar; the absolute floor is 1e-12 of the column's scale (its largest value or the sky irradiance): the
    reference's own solvers do not resolve less.  2s gets 1e-9 / 1e-10 of the column scale: at omega -> 1 its closed form
    (the reference's as well) loses digits like 1/h -- e.g. the upward flux over a soil of reflectance 1e-12 comes
    out as 5.2e-10 here, 1.1e-9 in the reference and is 4.0e-10."""
    from crt1d_b200 import cases
    from util import with_callables

    rng = np.random.default_rng(11)
    for trial in range(16):
        nz = int(rng.choice([3, 4, 7, 20, 60, 130]))
        q = dict(cases.load_default_case(nz))
        n = 24
        om = np.concatenate([10.0 ** rng.uniform(-6, -1, 6), rng.uniform(0.02, 0.98, 10), 1.0 - 10.0 ** rng.uniform(-12, -2, 8)])
        fr = rng.uniform(0.02, 0.98, n)
        q["leaf_r"], q["leaf_t"] = om * fr, om * (1 - fr)
        q["soil_r"] = rng.choice([1e-12, 1e-3, 0.1, 0.5, 1.0], n)
        q["I_dr0_all"] = rng.choice([0.0, 1e-12, 0.7, 300.0], n)
        q["I_df0_all"] = rng.choice([0.0, 1e-9, 0.5, 100.0], n)
        q["wl"] = np.linspace(0.4, 2.5, n)
        q["dwl"] = np.full(n, q["wl"][1] - q["wl"][0])
        q["wl_leafsoil"] = q["wl"]
        q["psi"] = np.deg2rad(rng.choice([0.0, 1e-6, 30.0, 60.0, 85.0, 89.0, 89.9]))
        x = np.linspace(1, 0, nz)
        q["lai"] = x ** int(rng.choice([1, 2, 4])) * float(rng.choice([1e-2, 0.5, 3.0, 8.0, 20.0]))
        q = with_callables(q)
        scale = (q["I_dr0_all"] + q["I_df0_all"])[None, :]
        for scheme in SCHEMES:
            kw = {"tau_d_method": "9sky"} if scheme == "n79" else {}
            with np.errstate(all="ignore"):
                ref = oracle.run(scheme, q, **kw)
            sol = _solve(q, scheme, **kw)
            for k in ("I_dr", "I_df_d", "I_df_u", "F"):
                fin = np.isfinite(ref[k])
                assert np.all(np.isfinite(sol[k][fin])), f"fuzz {trial} {scheme}.{k}: reference finite, kernel math not"
                r, v = np.where(fin, ref[k], 0.0), np.where(fin, sol[k], 0.0)
                atol = (1e-10 if scheme == "2s" else 1e-12) * np.maximum(np.max(np.abs(r), axis=0, keepdims=True),
                                                                         scale * (4.0 if k == "F" else 1.0))
                assert_close(v, r, 1e-9 if scheme == "2s" else RTOL, f"fuzz {trial} nz={nz} {scheme}.{k}", atol=atol)


def test_kernel_math_4s_extreme_inputs_fuzz():
    """The same kind of seeded fuzz for 4s against the reference run at tight `solve_bvp` tolerance (slow: three trials of
    eight bands here; 12 trials were run when the test was written): omega from 1e-4 to 1 - 1e-9 (both sides of the vanishing
    eigenvalue), black / bright soil, zero and huge irradiances, sun from the zenith to 1 degree above the horizon."""
    from crt1d_b200 import cases
    from util import with_callables

    rng = np.random.default_rng(0)
    for trial in range(3):
        nz = int(rng.choice([3, 7, 20, 60]))
        q = dict(cases.load_default_case(nz))
        n = 8
        om = np.concatenate([10.0 ** rng.uniform(-4, -1, 2), rng.uniform(0.02, 0.98, 3), 1.0 - 10.0 ** rng.uniform(-9, -2, 3)])
        fr = rng.uniform(0.05, 0.95, n)
        q["leaf_r"], q["leaf_t"] = om * fr, om * (1 - fr)
        q["soil_r"] = rng.choice([0.0, 1e-6, 0.1, 0.5, 1.0], n)
        q["I_dr0_all"] = rng.choice([0.0, 0.7, 300.0], n)
        q["I_df0_all"] = rng.choice([0.0, 0.5, 100.0], n)
        q["wl"] = np.linspace(0.4, 2.5, n)
        q["dwl"] = np.full(n, q["wl"][1] - q["wl"][0])
        q["wl_leafsoil"] = q["wl"]
        q["psi"] = np.deg2rad(rng.choice([0.0, 30.0, 60.0, 85.0, 89.0]))
        lai_tot = float(rng.choice([1e-2, 0.5, 3.0, 8.0, 15.0]))
        q["lai"] = np.linspace(1, 0, nz) ** int(rng.choice([1, 2])) * lai_tot
        q = with_callables(q)
        with np.errstate(all="ignore"):
            ref = oracle.solve_4s_tight(**{k: q[k] for k in oracle.ARGS["4s"]})
        sol = _solve(q, "4s")
        for k in ref:
            assert np.all(np.isfinite(sol[k][np.isfinite(ref[k])])), f"4s fuzz {trial}.{k}: reference finite, kernel math not"
            assert_close_4s(sol[k], ref[k], f"4s fuzz {trial} nz={nz}.{k}")


def test_leaf_angle_device_functions():
    from crt1d_b200.leaf_angle import LeafAngle
    from crt1d_b200.solvers import common

    L = hostcheck.lib()
    psi = np.radians(np.linspace(0, 89.5, 180))
    for fam, par in (("spherical", 0), ("horizontal", 0), ("vertical", 0), ("ellipsoidal_approx", 0.9632),
                     ("ellipsoidal", 2.5), ("ellipsoidal", 0.5), ("ellipsoidal", 1.0), ("ellipsoidal_approx_bonan", 0.25),
                     ("ellipsoidal_approx_bonan", 0.9)):
        la = LeafAngle(fam, par)
        got = np.array([L.hostcheck_leaf_G(la.family_id, la.param, a) for a in psi])
        assert_close(got, la.G_fn(psi) * np.ones_like(psi), 1e-14, f"G {fam}")
    la = LeafAngle("ellipsoidal_approx", 0.9632)
    # Device rule: Gauss-Legendre, 32 nodes x 12 panels graded towards psi = pi/2.  It self-converges to
    # ~1e-16; QUADPACK(epsrel=1e-9), which the reference uses, is itself within ~5e-12 of that.
    for Lv in (0.004, 0.0678, 0.5, 2.0, 6.0, 10.0):
        got = L.hostcheck_tau_d(la.family_id, la.param, 32, 12, Lv)
        assert abs(got - common.tau_df_fn(la.K_b_fn, Lv)) < 2e-11, Lv
        assert abs(got - L.hostcheck_tau_d(la.family_id, la.param, 64, 12, Lv)) < 2e-15, Lv
        assert abs(got - L.hostcheck_tau_d(la.family_id, la.param, 128, 12, Lv)) < 2e-15, Lv
        assert_close(L.hostcheck_tau_d(la.family_id, la.param, 0, 1, Lv), common.tau_df_fn(la.K_b_fn, Lv, method="9sky"),
                     1e-14, "9sky")
    g1, g2 = common.G_sector_integrals(la.G_fn, 0.501)
    assert abs(L.hostcheck_leaf_integral(la.family_id, la.param, 0.501, 32, 0) - common.mu_bar_fn(la.G_fn)) < 1e-13
    assert abs(L.hostcheck_leaf_integral(la.family_id, la.param, 0.501, 32, 1) - g1) < 1e-13
    assert abs(L.hostcheck_leaf_integral(la.family_id, la.param, 0.501, 32, 2) - g2) < 1e-13


def test_kernel_math_deep_canopy_nz1000():
    q = _deep_case()
    for scheme in ("2s", "bf", "g77", "zq", "zq_pa", "n79"):
        kw = {"tau_d_method": "9sky"} if scheme == "n79" else {}  # 999 quad calls are slow; 9sky is exact to compare
        ref = oracle.run(scheme, q, **kw)
        sol = _solve(q, scheme, **kw)
        for k in ref:
            # 2 x 1000 unknowns per tridiagonal column: measured 3e-13 (n79), 2e-13 (zq), 2e-14 (zq_pa) against the
            # oracle; the oracle itself moves by up to 3e-12 (n79 aI_lsh) when `lai` is perturbed by one ulp
            assert_close(sol[k], ref[k], RTOL, f"deep {scheme}.{k}")
    ref = oracle.solve_4s_tight(**{k: q[k] for k in oracle.ARGS["4s"]})  # all six bands (2.6 s of solve_bvp at tol = 1e-11)
    sol = _solve(q, "4s")
    for k in ref:
        assert_close_4s(sol[k], ref[k], f"deep 4s.{k}")


def test_exp_pm_accuracy():
    """exp_pm (shared-range-reduction e^-x, e^+x used by every level sweep) vs numpy: <= 2 ulp on [0, 700]."""
    import ctypes as C

    L = hostcheck.lib()
    L.hostcheck_exp_pm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(0, 30, 400000), rng.uniform(0, 700, 100000),
                        [0.0, 1e-300, 1e-17, 0.34657359027997264, 0.3465735902799727, 699.999, 700.0]])
    em, ep = np.empty_like(x), np.empty_like(x)
    L.hostcheck_exp_pm(x.size, x.ctypes.data, em.ctypes.data, ep.ctypes.data)
    assert np.max(np.abs(em - np.exp(-x)) / np.exp(-x)) < 4.5e-16
    assert np.max(np.abs(ep - np.exp(x)) / np.exp(x)) < 4.5e-16
    big = np.array([700.5, 745.0, 800.0, np.inf])  # beyond the fast path: libm semantics (underflow / overflow)
    em, ep = np.empty_like(big), np.empty_like(big)
    L.hostcheck_exp_pm(big.size, big.ctypes.data, em.ctypes.data, ep.ctypes.data)
    with np.errstate(over="ignore"):
        assert np.array_equal(em, np.exp(-big)) and np.array_equal(ep, np.exp(big))


def test_4s_irregular_levels_and_huge_lai(default_p):
    """4s paths the default cases never reach: irregularly spaced levels (no recurrence) vs the tight
    oracle, and lam * LAI >= 600 (unscaled coefficients + direct exponentials): finite, and the boundary
    conditions of ref _solve_4s.py:99-138 hold (downward diffuse at the top = sky diffuse; upward at the
    ground = soil reflection of what arrives there)."""
    sub = {k: (v[::40].copy() if isinstance(v, np.ndarray) and v.shape == (107,) else v) for k, v in default_p.items()}
    q = dict(sub, lai=np.linspace(1, 0, 37) ** 2 * 5.0)
    ref = oracle.solve_4s_tight(**{k: q[k] for k in oracle.ARGS["4s"]})
    sol = _solve(q, "4s")
    for k in ref:
        assert_close_4s(sol[k], ref[k], f"irregular 4s.{k}")
    big = dict(sub, lai=np.linspace(1, 0, 50) * 400.0)
    sol = _solve(big, "4s")
    for k in ("I_dr", "I_df_d", "I_df_u", "F"):
        assert np.all(np.isfinite(sol[k])), k
    assert_close(sol["I_df_d"][-1], big["I_df0_all"], 1e-9, "top boundary: I_df_d = I_df0")
    assert_close(sol["I_df_u"][0], big["soil_r"] * (sol["I_df_d"][0] + sol["I_dr"][0]), 1e-9, "soil boundary", atol=1e-300)
    # moderately deep canopy straddling the switch: same physics from both code paths
    a = _solve(dict(sub, lai=np.linspace(1, 0, 50) * 300.0), "4s")   # lam0 * LAI < 600 for these bands? checked below
    b = _solve(dict(sub, lai=np.linspace(1, 0, 50) * 300.0 * (1 + 1e-12)), "4s")
    assert_close(a["I_df_d"][-5:], b["I_df_d"][-5:], 1e-8, "continuity in LAI")


def test_smear_tuv_device_function_matches_reference():
    """`smear_tuv_bin` (bisection start + the reference's trapezoid loop) compiled for the host vs the
    reference-generated vectors; without FMA contraction it is bit-identical."""
    g = golden("ref_smear_tuv.npz")
    for x, y, bins, res in (("a_x", "a_y", "a_bins10", "a_bins10_res"), ("a_x", "a_y", "a_bins37", "a_bins37_res"),
                            ("a_x", "a_y", "a_bins_off", "a_bins_off_res"), ("b_x", "b_y", "b_bins", "b_res"),
                            ("c_x", "c_y", "c_bins", "c_res")):
        got = hostcheck.smear_tuv(g[x], g[y], g[bins])
        assert np.array_equal(got, g[res]), bins


@pytest.mark.parametrize("vec", (1, 2))
def test_4s_edge_cases_vs_reference(vec):
    """omega -> 1 (the smaller eigenvalue^2 of the four-stream operator vanishes at omega* = 0.99536 and is negative
    beyond: oscillatory mode), omega* itself, and kappa = lambda_k resonances -- against reference-generated
    fixtures (tests/golden/make_golden_4s_edge.py): the reference integrates the ODE (ref _solve_4s.py:48-97,
    235-262) and is finite and smooth through all of them."""
    for tag, mu_s, q, ship, tight in edge_4s_fixture_cases():
        n = q["leaf_r"].size
        if vec == 2 and n % 2:
            q = {k: (v[:n - 1].copy() if isinstance(v, np.ndarray) and v.shape == (n,) else v) for k, v in q.items()}
            ship = {k: v[:, :n - 1] for k, v in ship.items()}
            tight = {k: v[:, :n - 1] for k, v in tight.items()}
        sol = _solve(q, "4s", vec=vec, mu_s=mu_s)
        for k in tight:
            assert np.all(np.isfinite(sol[k])), (tag, k)
            assert_close_4s(sol[k], tight[k], f"{tag} tight {k}")
            assert_close_4s_shipped(sol[k], ship[k], f"{tag} shipped {k}")


def test_4s_edge_inputs_match_fixture():
    """The fixture's stored inputs are what util.edge_4s_case builds (root finding for omega* and the resonances)."""
    for tag, mu_s, q, _, _ in edge_4s_fixture_cases():
        psi_deg = float(tag.split("_")[0][3:])
        q2 = edge_4s_case(psi_deg, mu_s)
        for k in ("lai", "soil_r", "I_dr0_all", "I_df0_all"):
            assert np.array_equal(q[k], q2[k]), (tag, k)
        for k in ("leaf_r", "leaf_t"):
            np.testing.assert_allclose(q[k], q2[k], rtol=1e-13, err_msg=f"{tag} {k}")


def test_4s_thin_canopy_resonance_and_conservative_limit():
    """The entire-basis variants no default-LAI case reaches: kappa = lambda_1 in a thin canopy (kappa LAI < 0.5:
    series particular solution), omega = 1 exactly, and a sweep of omega across omega* -- vs the tight oracle."""
    for lai_tot, sel in ((0.5, (5, 7, 9, 10, 11)), (0.05, (0, 5, 9))):
        q = edge_4s_case(10.0, 0.501, lai_tot=lai_tot)
        n = q["leaf_r"].size
        q = {k: (v[list(sel)].copy() if isinstance(v, np.ndarray) and v.shape == (n,) else v) for k, v in q.items()}
        ref = oracle.solve_4s_tight(**{k: q[k] for k in oracle.ARGS["4s"]})
        sol = _solve(q, "4s", vec=1)
        for k in ref:
            assert_close_4s(sol[k], ref[k], f"thin {lai_tot} {k}")
    q = edge_4s_case(45.0, 0.501)
    n = q["leaf_r"].size
    om = np.array([1.0, 0.9953, 0.99537])
    q = {k: (v[:3].copy() if isinstance(v, np.ndarray) and v.shape == (n,) else v) for k, v in q.items()}
    q["leaf_r"], q["leaf_t"] = 0.5 * om, 0.5 * om
    ref = oracle.solve_4s_tight(**{k: q[k] for k in oracle.ARGS["4s"]})
    sol = _solve(q, "4s", vec=1)
    for k in ref:
        assert_close_4s(sol[k], ref[k], f"omega=1 {k}")


def test_4s_resonance_window_accuracy():
    """How wide the kappa ~ lambda_k window of coef_4s must be.  Bands whose omega brings lambda_0^2 (then lambda_1^2)
    to within delta = |kappa^2 - lambda_k^2| / kappa^2 of kappa^2, delta from 1e-1 down to 1e-9 on both sides, through
    three builds of the same kernel source: a wide window (0.05: the resonance-safe form wherever the ordinary one loses
    more than 4e-13), the ordinary form everywhere (window = 0), and the default.  The ordinary form loses ~2e-14 / delta;
    the default must stay within 5e-11 of the wide build for every delta (it switches at CRT_4S_RESONANCE_WIDTH = 1e-3)."""
    from scipy.optimize import brentq

    from crt1d_b200.leaf_angle import LeafAngle
    from crt1d_b200.solvers import common
    from util import fourstream_l2

    safe = hostcheck.build_variant("4s_wide", ["CRT_4S_ENTIRE_BELOW=-1.0", "CRT_4S_RESONANCE_WIDTH=0.05"])
    plain = hostcheck.build_variant("4s_plain", ["CRT_4S_ENTIRE_BELOW=-1.0", "CRT_4S_RESONANCE_WIDTH=-1.0"])
    la = LeafAngle.from_mla(57)
    mu_s = 0.501
    G1, G2 = common.G_sector_integrals(la.G_fn, mu_s)
    worst_default, worst_plain_outside = 0.0, 0.0
    for target, which in ((3.0, 0), (0.3, 1)):
        psi = brentq(lambda p: la.K_b_fn(p) ** 2 - target, 0.01, 1.55)
        k2 = la.K_b_fn(psi) ** 2
        om0 = brentq(lambda o: fourstream_l2(o, G1, G2, mu_s)[which] - k2, 0.01, 0.99)
        d = np.logspace(-1, -9, 120)
        om = np.r_[om0 - d * om0, om0 + d * (1 - om0) * 0.5]
        om = om[(om > 0.01) & (om < 0.99)]
        delta = np.array([abs(k2 - fourstream_l2(o, G1, G2, mu_s)[which]) / k2 for o in om])
        nw = om.size
        for lai_tot in (0.5, 3.0, 8.0):
            b = ScenarioBatch(psi=[psi], lai_lib=np.linspace(1, 0, 30) * lai_tot, leaf_r_lib=0.55 * om, leaf_t_lib=0.45 * om,
                              soil_r_lib=np.full(nw, 0.2), I_dr0_lib=np.full(nw, 1.0), I_df0_lib=np.full(nw, 0.3),
                              lai_idx=[0], leaf_idx=0, soil_idx=0, sky_idx=0, leaf_angle=la)
            pro = host_prologue(b, "4s")
            dflt = hostcheck.solve(b, "4s", pro, vec=1)
            with hostcheck.use_lib(safe):
                ref = hostcheck.solve(b, "4s", pro, vec=1)
            with hostcheck.use_lib(plain):
                ordn = hostcheck.solve(b, "4s", pro, vec=1)
            for k in ("I_df_d", "I_df_u"):
                scale = np.max(np.abs(ref[k][0]), axis=0)
                e_d = np.max(np.abs(dflt[k][0] - ref[k][0]), axis=0) / scale
                e_o = np.max(np.abs(ordn[k][0] - ref[k][0]), axis=0) / scale
                worst_default = max(worst_default, float(e_d.max()))
                worst_plain_outside = max(worst_plain_outside, float(e_o[delta > 1e-3].max()))
                assert e_o[delta < 1e-6].max() > 1e-9  # the ordinary form really does fail close to the resonance
    assert worst_default < 5e-11, worst_default
    assert worst_plain_outside < 5e-11, worst_plain_outside

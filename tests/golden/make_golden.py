"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run HERE (build container; needs /root/reference):   python tests/golden/make_golden.py
The reference is Python and cannot travel to the GPU box, so its outputs on fixed inputs are committed
as small .npz fixtures together with this script.  Inputs come from the product's own host-side case
builders (`crt1d_b200.cases`, `crt1d_b200.sweep`), whose checksums are pinned to SURVEY.md section 8c.

Fixtures written:
  default_inputs.npz          default case (n_z = 60, n_wl = 107) inputs + band-independent prologue scalars
  ref_default_<id>.npz        every array each reference scheme returns on the default case
                              (ids: 2s 4s bf bl g77 n79 zq zq_pa; plus n79_9sky, 4s_tight, 4s_mus034,
                              4s_mus034_tight)
  ref_variants.npz            n_z in {2,3,10,200}, psi in {0,20,60,85} deg on a 12-band subset
  ref_bonan_n79.npz           the reference's Bonan SP 14.3 set-up (tests/test_n79.py) with "9sky"
  ref_absorption.npz          Model.run + _calc_absorption of the reference (2s, bf, zq default; n79 Bonan) (recipe B)
  ref_sweep_2s.npz            a strided sample of the config-3 synthetic sweep through reference 2s/4s-tight
  band_weight_kat.npz         the reference test-suite's own known answers for _x_frac_in_bounds
"""
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from _refimport import import_reference_solvers  # noqa: E402
from _refimport import import_reference_spectra  # noqa: E402
from _refimport import tight_4s_solver  # noqa: E402

from crt1d_b200 import cases  # noqa: E402  (host-side only; no CUDA needed)


def full_params(p):
    """What Model._check_inputs derives (ref model.py:236-293) that the solvers consume."""
    q = dict(p)
    G_fn = q["G_fn"]
    q["K_b_fn"] = lambda psi_: G_fn(psi_) / np.cos(psi_)
    q["G"] = G_fn(q["psi"])
    q["K_b"] = q["K_b_fn"](q["psi"])
    return q


def run_ref(S, scheme, p, **extra):
    sd = S.AVAILABLE_SCHEMES[scheme]
    return sd["solver"](**{k: p[k] for k in sd["args"]}, **extra)


def savez(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"  wrote {name}: {os.path.getsize(path)/1024:.0f} KB")


def main():
    S, LA, LAREA = import_reference_solvers()
    t0 = time.time()

    # ---------------------------------------------------------------- default case
    p = full_params(cases.load_default_case(60))
    # the product's host-side case builder must agree with the reference's own pieces
    ref_prof = LAREA.distribute_lai_beta(20.0, 4.0, 60)
    assert np.array_equal(ref_prof.lai, p["lai"]) and np.allclose(ref_prof.z, p["z"], rtol=0, atol=0)
    assert float(LA.mla_to_x_approx(57)) == float(p["orient"])
    assert abs(p["I_dr0_all"].sum() - 765.9542860103) < 1e-9 and abs(p["I_df0_all"].sum() - 216.5547763843) < 1e-9
    ref_G = lambda psi_: LA.G_ellipsoidal_approx(psi_, p["orient"])  # noqa: E731
    for ang in (0.0, 0.3, 1.0, 1.5):
        assert ref_G(ang) == p["G_fn"](ang)

    import scipy.integrate as integ
    import math
    mu_bar = integ.quad(lambda sa: math.cos(sa) / p["G_fn"](sa) * -math.sin(sa), math.pi / 2, 0)[0]
    savez(
        "default_inputs.npz",
        **{k: p[k] for k in ("lai", "z", "psi", "mla", "orient", "clump", "leaf_t", "leaf_r", "soil_r",
                             "I_dr0_all", "I_df0_all", "wl", "dwl", "G", "K_b")},
        mu_bar=mu_bar,
    )

    for scheme in S.AVAILABLE_SCHEMES:
        t1 = time.time()
        sol = run_ref(S, scheme, p)
        print(f"{scheme}: {time.time()-t1:.2f} s")
        savez(f"ref_default_{scheme}.npz", **sol)
    savez("ref_default_n79_9sky.npz", **run_ref(S, "n79", p, tau_d_method="9sky"))
    solve_4s_tight = tight_4s_solver()
    args4 = {k: p[k] for k in S.AVAILABLE_SCHEMES["4s"]["args"]}
    t1 = time.time()
    savez("ref_default_4s_tight.npz", **solve_4s_tight(**args4))
    print(f"4s tight: {time.time()-t1:.1f} s")
    savez("ref_default_4s_mus034.npz", **run_ref(S, "4s", p, mu_s=0.33998))
    savez("ref_default_4s_mus034_tight.npz", **solve_4s_tight(**args4, mu_s=0.33998))

    # ---------------------------------------------------------------- variants (12-band subset)
    sub = slice(0, None, 9)
    var = {}
    var_list = [(2, 20), (3, 20), (10, 60), (200, 20), (60, 0), (60, 60), (60, 85)]
    for nz, sza in var_list:
        q = dict(cases.load_default_case(nz))
        q["psi"] = np.deg2rad(sza)
        for k in ("leaf_t", "leaf_r", "soil_r", "I_dr0_all", "I_df0_all", "wl", "dwl", "wl_leafsoil"):
            q[k] = q[k][sub].copy()
        q = full_params(q)
        tag = f"nz{nz}_sza{sza}"
        var[f"{tag}__lai"] = q["lai"]
        var[f"{tag}__psi"] = q["psi"]
        for scheme in ("2s", "bf", "bl", "g77", "n79", "zq", "zq_pa", "4s_tight"):
            try:
                if scheme == "4s_tight":
                    if nz > 60:
                        continue
                    sol = solve_4s_tight(**{k: q[k] for k in S.AVAILABLE_SCHEMES["4s"]["args"]})
                else:
                    sol = run_ref(S, scheme, q)
            except Exception as e:  # the reference itself fails (e.g. n79 needs n_z >= 3)
                var[f"{tag}__{scheme}__raises"] = np.array(type(e).__name__)
                print(f"  {tag} {scheme}: reference raises {type(e).__name__}: {e}")
                continue
            for k, v in sol.items():
                var[f"{tag}__{scheme}__{k}"] = v
    for k in ("leaf_t", "leaf_r", "soil_r", "I_dr0_all", "I_df0_all", "wl", "dwl"):
        var[f"sub__{k}"] = p[k][sub]
    savez("ref_variants.npz", **var)

    # ---------------------------------------------------------------- Bonan SP 14.3 n79 case
    qb = full_params(cases.load_bonan_sp1403_case())
    ref_b = LAREA.distribute_lai_beta_bonan(20, 6, 61)
    assert np.array_equal(ref_b.lai, qb["lai"]) and np.array_equal(ref_b.z, qb["z"])
    sol = run_ref(S, "n79", qb, tau_d_method="9sky")
    savez("ref_bonan_n79.npz", lai=qb["lai"], z=qb["z"], **sol)

    # ---------------------------------------------------------------- band weights KAT (ref tests/test_spectra.py:25-35)
    SP = import_reference_spectra()
    kat = {}
    for i, (xe, bounds, expected) in enumerate(
        [(np.r_[0, 1, 2, 3], (0, 3), [1, 1, 1]),
         (np.r_[0, 1, 2, 3], (0.5, 2.2), [0.5, 1, 0.2]),
         (np.r_[0, 1, 2, 3], (0.5, 2.0), [0.5, 1, 0])]
    ):
        got = SP._x_frac_in_bounds(xe, bounds)
        np.testing.assert_allclose(got, expected)
        kat[f"k{i}_xe"], kat[f"k{i}_bounds"], kat[f"k{i}_w"] = xe, np.array(bounds), got
    wle = np.r_[p["wl"][0] - 0.5 * p["dwl"][0], p["wl"] + 0.5 * p["dwl"]]
    kat["default_wle"] = wle
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kat["default_w_PAR"] = SP._x_frac_in_bounds(wle, SP.BAND_DEFNS_UM["PAR"])
        kat["default_w_NIR"] = SP._x_frac_in_bounds(wle, SP.BAND_DEFNS_UM["NIR"])
    savez("band_weight_kat.npz", **kat)

    # ---------------------------------------------------------------- sweep sample (config 3)
    from crt1d_b200 import sweep as sw

    spec = sw.synthetic_sweep_spec(seed=0)
    idx = np.array([0, 123456, 250250, 499999, 777777, 999999])
    bsub = slice(0, None, 50)
    out = {"scenario_index": idx, "band_subset_step": np.array(50)}
    for n, s in enumerate(idx):
        q = full_params(spec.scenario_params(int(s)))
        sol = run_ref(S, "2s", q)
        for k, v in sol.items():
            out[f"s{n}__2s__{k}"] = v[:, bsub]
        if n in (1, 4):
            qs = dict(q)
            for k in ("leaf_t", "leaf_r", "soil_r", "I_dr0_all", "I_df0_all"):
                qs[k] = q[k][bsub].copy()
            sol = solve_4s_tight(**{k: qs[k] for k in S.AVAILABLE_SCHEMES["4s"]["args"]})
            for k, v in sol.items():
                out[f"s{n}__4s_tight__{k}"] = v
    savez("ref_sweep_2s.npz", **out)

    # ---------------------------------------------------------------- Model.run + _calc_absorption (recipe B, own process)
    subprocess.check_call([sys.executable, os.path.join(HERE, "make_golden_absorption.py")])
    print(f"done in {time.time()-t0:.0f} s")


if __name__ == "__main__":
    main()

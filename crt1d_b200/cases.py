"""Complete input-parameter sets (host side): the reference's default case and the n79/Bonan case.

`load_default_case` restates ref `crt1d/cases.py:15-58` + `crt1d/data/__init__.py:90-217` in numpy
(the reference needs xarray for this).  Spectra come from the packed `data/spectra_lib.npz`
(see tools/make_data_lib.py).
"""
import os

import numpy as np

from .leaf_angle import G_ellipsoidal_approx
from .leaf_angle import G_spherical
from .leaf_angle import LeafAngle
from .leaf_angle import mla_to_x_approx
from .leaf_area import distribute_lai_beta
from .leaf_area import distribute_lai_beta_bonan
from .leaf_area import distribute_lai_from_cdd

_LIB_PATH = os.path.join(os.path.dirname(__file__), "data", "spectra_lib.npz")
_lib_cache = None


def spectra_lib():
    """The packed sample spectra (dict of float64 arrays)."""
    global _lib_cache
    if _lib_cache is None:
        with np.load(_LIB_PATH) as f:
            _lib_cache = {k: np.array(f[k], dtype=np.float64) for k in f.files}
    return _lib_cache


def default_spectra():
    """Default toc irradiance + ideal-green-leaf + two-value soil on SPCTRAL2 midpoints, rows with a
    missing leaf value dropped  (ref data/__init__.py:23-35, 90-147, 188-217; cases.py:25)."""
    lib = spectra_lib()
    t = lib["ideal_t"].copy()
    r = lib["ideal_r"].copy()
    t[t == 0] = 1e-10  # before midpointing, as the reference does (data/__init__.py:96-102)
    r[r == 0] = 1e-10
    t = 0.5 * (t[:-1] + t[1:])
    r = 0.5 * (r[:-1] + r[1:])
    wl_leaf = 0.5 * (lib["ideal_wl_um"][:-1] + lib["ideal_wl_um"][1:])

    wl0 = lib["sp2_wl_um"]
    dwl = np.diff(wl0)
    wl = wl0[:-1] + 0.5 * dwl
    I_dr = 0.5 * (lib["sp2_SI_dr"][:-1] + lib["sp2_SI_dr"][1:]) * dwl
    I_df = 0.5 * (lib["sp2_SI_df"][:-1] + lib["sp2_SI_df"][1:]) * dwl
    assert np.allclose(wl, wl_leaf)

    rs = np.where(wl_leaf <= 0.7, 0.1100, 0.2250)
    keep = np.isfinite(t) & np.isfinite(r) & np.isfinite(I_dr) & np.isfinite(I_df)
    return dict(
        wl=wl[keep], dwl=dwl[keep], leaf_t=t[keep], leaf_r=r[keep], soil_r=rs[keep],
        I_dr0_all=I_dr[keep], I_df0_all=I_df[keep],
    )


def load_default_case(nlayers):
    """h_c = 20 m, LAI = 4, mla = 57 deg (ellipsoidal-approx G), SZA = 20 deg  (ref cases.py:15-58)."""
    prof = distribute_lai_beta(20.0, 4.0, nlayers)
    sp = default_spectra()
    mla = 57
    orient = mla_to_x_approx(mla)
    G_fn = lambda psi_: G_ellipsoidal_approx(psi_, orient)  # noqa: E731
    return dict(
        lai=prof.lai, z=prof.z, green=1.0, mla=mla, clump=1.0, orient=orient, G_fn=G_fn,
        leaf_angle=LeafAngle("ellipsoidal_approx", float(orient)),
        psi=np.deg2rad(20),
        leaf_t=sp["leaf_t"], leaf_r=sp["leaf_r"], soil_r=sp["soil_r"], wl_leafsoil=sp["wl"],
        I_dr0_all=sp["I_dr0_all"], I_df0_all=sp["I_df0_all"], wl=sp["wl"], dwl=sp["dwl"],
    )


def load_bonan_sp1403_case():
    """Inputs of the reference's only known-answer test, Bonan (2019) SP 14.3  (ref tests/test_n79.py:11-44)."""
    prof = distribute_lai_beta_bonan(20, 6, 61)
    wl = np.r_[0.55, 1.6]
    return dict(
        lai=prof.lai, z=prof.z, psi=30 * (np.pi / 180),
        leaf_r=np.r_[0.1, 0.45], leaf_t=np.r_[0.05, 0.25], soil_r=np.r_[0.1, 0.2],
        I_dr0_all=np.r_[0.8, 0.8], I_df0_all=np.r_[0.2, 0.2],
        wl=wl, wl_leafsoil=wl, dwl=np.r_[0.3, 1.8], clump=1.0, G_fn=G_spherical,
        leaf_angle=LeafAngle("spherical", 0.0), mla=60.0, green=1.0, orient=1.0,
    )


# The reference's default canopy description (Borden 1995 field study, late June: ref data/default_canopy_descrip.csv),
# storeys listed uppermost first
BORDEN95_CANOPY_DESCRIP = dict(
    lai_tot=3.044, lai_frac=[0.608, 0.392], h_ref=34.2, h_canopy=22.0, h_max_lad=[15.4, 6.16], h_bot=[12.1, 1.375],
    h_top=[22.0, 12.0], lad_h_top=[0.0, 0.065], mla=57.4, lat=44.31666, lon=80.93333, green=1.0, clump=[0.930, 0.930],
)


def load_canopy_descrip(fpath):
    """Canopy-description CSV (`varname,value,...` rows; `;`-separated lists) -> dict  (ref cases.py:61-85)."""
    import csv

    d = {}
    with open(fpath, newline="") as fh:
        rows = csv.reader(fh)
        next(rows)  # header
        for row in rows:
            if len(row) < 2 or not row[0].strip():
                continue
            name, val = row[0].strip(), row[1]
            d[name] = [float(x) for x in val.split(";")] if ";" in val else float(val)
    assert np.array(d["lai_frac"]).sum() == 1.0
    return d


def borden95_profile(nlayers, cdd=None):
    """Equal-LAI-increment levels of the two-storey Borden canopy (`distribute_lai_from_cdd`).  The reference's
    `load_Borden95_default_case` stops right after this step with NotImplementedError (ref cases.py:88-93)."""
    return distribute_lai_from_cdd(dict(cdd or BORDEN95_CANOPY_DESCRIP), nlayers)

"""Rows f-2 / f-4 of SURVEY.md section 8: the dataset layout `Model.to_xr` produces, the band / photon-flux reductions
and the canopy energy balance, pinned to fixtures written by the reference's OWN `Model.to_xr`, `diagnostics.band` and
`diagnostics.compare_ebal` (tests/golden/make_golden_ebal.py runs them unmodified on the xarray stand-in
tests/golden/_xr_standin.py; xarray is not in this image)."""
import json
import os
import sys
import warnings

import numpy as np
import pytest

from util import GOLDEN
from util import assert_close
from util import golden

sys.path.insert(0, GOLDEN)
import _xr_standin as xrs  # noqa: E402

BANDS = ("PAR", "NIR", "solar", "UV")


def _meta():
    with open(os.path.join(GOLDEN, "ref_to_xr.json")) as fh:
        return json.load(fh)


def test_variable_metadata_matches_reference_datasets():
    """Every variable of every reference dataset: the product's table gives the same dims and attrs."""
    from crt1d_b200.variables import VMD
    from crt1d_b200.variables import da_attrs
    from crt1d_b200.variables import dims_of

    meta = _meta()
    for scheme in ("2s", "bf", "bl", "g77", "zq", "n79"):
        for name, v in meta[scheme]["variables"].items():
            base = name[: -len("_scheme")] if name.endswith("_scheme") else name
            assert base in VMD, base
            assert da_attrs(VMD[base]) == v["attrs"], (scheme, name)
            if not name.endswith("_scheme"):  # scheme arrays take their dims from their shape
                assert list(dims_of(VMD[base].shape)) == v["dims"], (scheme, name)


def test_band_and_pfd_weights_match_reference_band():
    """`diagnostics.band(ds, calc_PFD=True)` of the reference (diagnostics.py:39-108) on its own 2s dataset vs the
    product's host-side weights applied to the same profiles: W m-2 sums and photon flux densities, four bands."""
    from crt1d_b200 import spectra

    g = golden("ref_ebal.npz")
    wl, dwl, wle = g["toxr__2s__wl"], g["toxr__2s__dwl"], g["toxr__2s__wle"]
    assert np.array_equal(spectra.edges_from_centers_widths(wl, dwl), wle)
    for bn in BANDS:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # UV bounds extend outside the data range (the reference warns too)
            w = spectra.band_weights(wle, bn)
            wp = spectra.pfd_band_weights(wl, dwl, bn)
        for k in ("I_dr", "I_df_d", "I_df_u", "I_d", "aI", "aI_sl", "aI_df_sh"):
            prof = g[f"toxr__2s__{k}"]
            assert_close((prof * w).sum(axis=-1), g[f"band__2s__{bn}__{k}"], 1e-13, f"{bn} {k}", atol=1e-300)
            assert_close((prof * wp).sum(axis=-1), g[f"band__2s__{bn}__{k.replace('I', 'PFD')}"], 1e-13, f"{bn} PFD {k}", atol=1e-300)
        # the reference's `vn.replace("I", "PFD")` leaves "F" unchanged, so its band dataset holds the PHOTON actinic flux
        # under the name F (units relabelled, see ref_to_xr.json band_attrs) -- reproduced, not "fixed"
        assert_close((g["toxr__2s__F"] * wp).sum(axis=-1), g[f"band__2s__{bn}__F"], 1e-13, f"{bn} F")


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["2s", "bf", "n79", "zq"])
def test_to_xr_layout_matches_reference(scheme):
    """`Model.run().calc_absorption().to_xr()` materialised through the stand-in: same coords, variables, dims, shapes,
    attrs and values as the reference's dataset."""
    import crt1d_b200 as crt

    meta = _meta()[scheme]
    g = golden("ref_ebal.npz")
    xrs.install()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ds = crt.Model(scheme, nlayers=60).run().calc_absorption().to_xr(info="golden")
    assert isinstance(ds, xrs.Dataset)
    assert list(ds.coord_names) == meta["coords"]
    attrs = dict(ds.attrs)
    assert attrs.pop("crt1d_version") == crt.__version__
    assert attrs == meta["attrs"]
    assert set(ds.variables) == set(meta["variables"])
    for name, v in meta["variables"].items():
        a = ds[name]
        assert list(a.dims) == v["dims"] and list(a.values.shape) == v["shape"] and a.attrs == v["attrs"], name
        key = f"toxr__{scheme}__{name}"
        if key in g:
            tol = 1e-9 if name.startswith("aI") else 1e-10  # absorption: differences of profiles (cancellation)
            assert_close(a.values, g[key], tol, f"{scheme} to_xr {name}", atol=1e-12 if name.startswith("aI") else 0.0)


@pytest.mark.gpu
def test_energy_balance_matches_reference_compare_ebal():
    """`crt1d_energy_balance` (+ the absorption kernel for the layer-wise sum) vs the reference's `compare_ebal`
    (diagnostics.py:476-530) for six schemes x four bands: all six columns of its DataFrame."""
    import torch

    import crt1d_b200 as crt
    from crt1d_b200 import engine
    from crt1d_b200 import spectra

    meta, g = _meta(), golden("ref_ebal.npz")
    assert meta["ebal_columns"] == ["incoming", "outgoing (reflected)", "soil absorbed", "layerwise abs sum", "in-out-soil", "canopy abs"]
    assert engine.EBAL_COLUMNS == tuple(meta["ebal_columns"][i] for i in (0, 1, 2, 5))
    models = []
    for scheme in meta["ebal_index"]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            models.append(crt.Model(scheme, nlayers=60).run().calc_absorption())
    p = models[0]._p
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bw = np.stack([spectra.band_weights(p["wle"], b) for b in BANDS])
    prof = {k: torch.as_tensor(np.stack([m.out[k] for m in models])).cuda() for k in ("I_dr", "I_df_d", "I_df_u")}
    eb = engine.energy_balance(prof["I_dr"], prof["I_df_d"], prof["I_df_u"], bw).cpu().numpy()  # (scheme, band, 4)
    for b, bn in enumerate(BANDS):
        ref = g[f"ebal__{bn}"]  # (scheme, 6)
        scale = ref[:, 0:1]
        got = np.empty_like(ref)
        got[:, [0, 1, 2, 5]] = eb[:, b, :]
        got[:, 3] = [(m.absorption["aI"] * bw[b]).sum() for m in models]
        got[:, 4] = eb[:, b, 0] - eb[:, b, 1] - eb[:, b, 2]
        # sums of ~100 in-band terms of O(1): 1e-10 relative, or 1e-12 of the incoming flux for the columns that are
        # differences (soil, canopy) and for bl's identically zero reflected flux
        assert np.all(np.abs(got - ref) <= np.maximum(1e-10 * np.abs(ref), 1e-12 * scale)), (bn, np.abs(got - ref).max())


@pytest.mark.gpu
def test_sensitivity_dataset_materialised():
    """`run_sensitivity` -> `sensitivity_to_xr` through the stand-in: one leading dim per swept parameter in front of the
    reference's z / zm / wl dims; every slice equals the dataset of the corresponding single run."""
    import crt1d_b200 as crt
    from crt1d_b200.model import sensitivity_to_xr

    xrs.install()
    m = crt.Model("bf", nlayers=30)
    psis = [0.2, 0.9, 1.3]
    lais = [m._p["lai"], m._p["lai"] * 0.5]
    p_sets = {"psi": psis, "lai": lais}
    res = crt.run_sensitivity(m, p_sets)
    ds = sensitivity_to_xr(res, m, p_sets)
    assert isinstance(ds, xrs.Dataset)
    assert ds.dims["psi"] == 3 and ds.dims["lai_case"] == 2 and ds.dims["z"] == 30 and ds.dims["wl"] == m.nwl
    assert np.array_equal(ds["psi"].values, psis)
    assert tuple(ds["I_df_d"].dims) == ("psi", "lai_case", "z", "wl")
    assert tuple(ds["absorbed"].dims) == ("psi", "lai_case", "band") and list(ds["band"].values) == ["PAR", "NIR"]
    assert ds["I_dr"].attrs == {"long_name": "Direct beam irradiance (binned)", "units": "W m-2"}
    for i, psi in enumerate(psis):
        for j, lai in enumerate(lais):
            one = crt.Model("bf", nlayers=30, psi=psi, lai=lai).run().to_xr()
            for k in ("I_dr", "I_df_d", "I_df_u", "F", "aI_l_scheme", "aI_lsl_scheme"):
                name = k[: -len("_scheme")] if k.endswith("_scheme") else k
                assert tuple(ds[name].dims)[2:] == tuple(one[k].dims)
                # batched device prologue vs plugin host prologue: identical closed-form scalars for bf
                assert_close(ds[name].values[i, j], one[k].values, 1e-10, f"sens[{i},{j}].{k}", atol=1e-300)

"""ref_ebal.npz + ref_to_xr.json: the reference's own export and diagnostics, UNMODIFIED --
`Model.run().calc_absorption().to_xr()` (ref model.py:338-447), `diagnostics.band(ds, calc_PFD=True)`
(diagnostics.py:39-108) and `diagnostics.compare_ebal` (diagnostics.py:476-530) -- executed on the stand-in for
xarray in `_xr_standin.py` (xarray / matplotlib are not in this image; pandas is).  Own process (recipe B of SURVEY
appendix B).  Run HERE:   python tests/golden/make_golden_ebal.py"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = os.environ.get("CRT1D_REFERENCE", "/root/reference")

import _xr_standin  # noqa: E402

_xr_standin.install()
for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[m] = types.ModuleType(m)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
v = types.ModuleType("crt1d._version")
v.version = "0+ref"
sys.modules["crt1d._version"] = v
sys.path.insert(0, REF)
import crt1d  # noqa: E402
import crt1d.model  # noqa: E402
from crt1d import diagnostics  # noqa: E402

from crt1d_b200 import cases  # noqa: E402

SCHEMES = ("2s", "bf", "bl", "g77", "zq", "n79")
BANDS = ("PAR", "NIR", "solar", "UV")


def _case(nlayers):
    p = cases.load_default_case(nlayers)
    p.pop("leaf_angle")
    return p


def main():
    crt1d.model.load_default_case = _case
    out, meta = {}, {}
    dsets = []
    for scheme in SCHEMES:
        m = crt1d.Model(scheme, nlayers=60).run().calc_absorption()
        ds = m.to_xr(info="golden")
        dsets.append(ds)
        meta[scheme] = {
            "coords": list(ds.coord_names),
            "attrs": {k: (v if isinstance(v, str) else float(v)) for k, v in ds.attrs.items() if k != "crt1d_version"},
            "variables": {k: {"dims": list(a.dims), "shape": list(a.values.shape), "attrs": a.attrs} for k, a in ds.variables.items()},
        }
        for k, a in ds.variables.items():
            if scheme == "2s" or k in ("I_d", "aI"):
                out[f"toxr__{scheme}__{k}"] = a.values
        for bn in BANDS:
            b = diagnostics.band(ds, band_name=bn, calc_PFD=True)
            for k, a in b.variables.items():
                if k == "F" or "I" in k or "PFD" in k:
                    out[f"band__{scheme}__{bn}__{k}"] = a.values
            if scheme == "2s":
                meta.setdefault("band_attrs", {})[bn] = {
                    "ds": {k: (list(x) if isinstance(x, tuple) else x) for k, x in b.attrs.items() if k.startswith("band")},
                    "variables": {k: {"dims": list(a.dims), "attrs": a.attrs} for k, a in b.variables.items() if "I" in k or "PFD" in k or k == "F"}}
    for bn in BANDS:
        df = diagnostics.compare_ebal(dsets, band_name=bn)
        out[f"ebal__{bn}"] = df.to_numpy(dtype=np.float64)
        meta["ebal_columns"] = list(df.columns)
        meta["ebal_index"] = list(df.index)
    path = os.path.join(HERE, "ref_ebal.npz")
    np.savez_compressed(path, **out)
    with open(os.path.join(HERE, "ref_to_xr.json"), "w") as fh:
        json.dump(meta, fh, indent=1, ensure_ascii=False, sort_keys=True)
    print(f"  wrote ref_ebal.npz: {os.path.getsize(path) / 1024:.0f} KB; ref_to_xr.json")
    print(df)


if __name__ == "__main__":
    main()

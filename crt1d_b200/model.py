"""`Model`: the reference's front end for the solver hot path, with a batched scenario axis.

Mirrors the part of `crt1d.Model` that sits directly on the hot path (ref crt1d/model.py:49-336,
573-664): parameter handling and validation (`update_p`, `_check_inputs`), scheme dispatch through
`AVAILABLE_SCHEMES` (`run`), the layer-absorption post-step (`calc_absorption`) and -- new -- the
batched axis (`run_batch`, `run_sensitivity`, which the reference leaves as `NotImplementedError`,
ref model.py:650-664).  Plotting stays out of scope; `to_xr` is kept but imports xarray lazily.
"""
import itertools
import warnings
from collections import namedtuple
from copy import deepcopy

import numpy as np

from .cases import load_default_case
from .leaf_angle import LeafAngle
from .scenarios import ScenarioBatch
from .solvers import AVAILABLE_SCHEMES
from .solvers import RET_KEYS_ALL_SCHEMES
from .spectra import BAND_DEFNS_UM
from .spectra import x_frac_in_bounds
from .variables import VMD
from .variables import da_attrs

__all__ = ("Model", "run_sensitivity")

CANOPY_DESCRIPTION_KEYS = [
    "lai", "z", "dlai", "lai_tot", "lai_eff", "mla", "clump", "leaf_t", "leaf_r", "soil_r", "wl_leafsoil",
    "orient", "G_fn",
]
CanopyDescription = namedtuple("CanopyDescription", " ".join(CANOPY_DESCRIPTION_KEYS))

ABSORPTION_KEYS = ("aI", "aI_df", "aI_dr", "aI_sh", "aI_sl", "aI_df_sl", "aI_df_sh")


class Model:
    """1-D canopy radiative transfer with a selectable scheme, solved on the GPU."""

    required_input_keys = tuple(
        [k for k in CANOPY_DESCRIPTION_KEYS if k not in ("dlai", "lai_tot", "lai_eff")]
        + ["I_dr0_all", "I_df0_all", "wl", "dwl", "psi"]
    )
    _optional_input_keys = ("leaf_angle", "green")
    _schemes = AVAILABLE_SCHEMES
    vmd = VMD

    def __init__(self, scheme="2s", nlayers=60, **p_kwargs):
        self.nlayers = nlayers
        self.p_default = load_default_case(nlayers=self.nlayers)
        self._p = deepcopy(self.p_default)
        self.assign_scheme(scheme)
        if p_kwargs:
            self.update_p(**p_kwargs)
        else:
            self._check_inputs()
        self._run_count = 0
        self.absorption = None
        self.out = {}
        self.out_extra = {}

    # ------------------------------------------------------------------ parameters
    @property
    def p(self):
        print(
            "Please update parameters using `.update_p()`! Changes to `.p` will not be stored!\n"
            "Extract (copy) the parameters using `.copy_p()` or summarize using `.print_p()`."
        )

    def print_p(self):
        import pprint

        with np.printoptions(precision=3, threshold=7):
            pprint.PrettyPrinter(indent=1).pprint(self._p)

    def copy_p(self):
        return deepcopy(self._p)

    @property
    def cd(self):
        return CanopyDescription(**{k: v for k, v in self._p.items() if k in CANOPY_DESCRIPTION_KEYS})

    def __repr__(self):
        return f"Model(scheme={self.scheme['name']!r}, psi={self._p['psi']:.4g})"

    def assign_scheme(self, scheme_name, *, verbose=False):
        """Select the scheme; an unknown id prints the valid ones and falls back to `2s`, like the
        reference (ref model.py:151-170)."""
        try:
            self.scheme = AVAILABLE_SCHEMES[scheme_name]
            if verbose:
                print("\n\n" + "=" * 40 + f"\nscheme: {self.scheme['name']}\n" + "-" * 40)
        except KeyError:
            print(f"{scheme_name!r} is not a valid scheme name/ID!")
            print(f"The valid ones are: {', '.join(AVAILABLE_SCHEMES)}.")
            print("Defaulting to Dickinson-Sellers two-stream.\n")
            self.scheme = AVAILABLE_SCHEMES["2s"]
        return self

    def update_p(self, **kwargs):
        """Update inputs; on any validation failure warn and revert (ref model.py:172-203)."""
        import traceback

        saved = deepcopy(self._p)
        try:
            for k, v in kwargs.items():
                if k not in Model.required_input_keys and k not in Model._optional_input_keys:
                    warnings.warn(f"{k!r} is not intended as an input and will be ignored")
                    continue
                self._p[k] = v
                if k == "G_fn" and "leaf_angle" not in kwargs:
                    self._p.pop("leaf_angle", None)  # a custom callable invalidates the parametric family
                if k == "leaf_angle" and "G_fn" not in kwargs and v is not None:
                    self._p["G_fn"] = v.G_fn  # one canopy for run() (G_fn) and run_batch() (leaf_angle)
            self._check_inputs()
        except Exception:
            warnings.warn(
                f"Updating parameters failed. Full traceback:\n\n{traceback.format_exc()}\nReverting."
            )
            self._p = saved
        return self

    def _check_inputs(self):
        """Validate the LAI profile / wavelength grids and derive dependent inputs (ref model.py:222-294)."""
        p = self._p
        for key in Model.required_input_keys:
            if key not in p:
                raise Exception(f"required key {key} is not present. Set it using `update_p`.")
        lai, z = np.asarray(p["lai"], dtype=float), np.asarray(p["z"], dtype=float)
        assert z.size == lai.size
        self.nlev = lai.size
        assert z[-1] > z[0]
        assert lai[0] > lai[-1]
        assert lai[-1] == 0
        dz = np.diff(z)
        dlai = lai[:-1] - lai[1:]
        p["lai_tot"] = lai[0]
        p["lai_eff"] = lai * p["clump"]
        p["dlai"] = dlai
        p["dlai_eff"] = dlai * p["clump"]
        p["zm"] = z[:-1] + 0.5 * dz
        p["dz"] = dz
        psi = p["psi"]
        if "mu" in p:
            if p["mu"] != np.cos(psi):
                warnings.warn(
                    "Provided `mu` not consistent with provided `psi`. "
                    "`mu` will be updated based on the value of `psi`."
                )
                p["mu"] = np.cos(psi)  # (the reference warns but leaves the stale value, ref model.py:258-267)
        else:
            p["mu"] = np.cos(psi)
        wl_toc, wl_op = np.asarray(p["wl"]), np.asarray(p["wl_leafsoil"])
        assert wl_toc.size == wl_op.size
        if not np.allclose(wl_toc, wl_op):
            warnings.warn(
                "Provided wavelengths for optical props (`wl_leafsoil`) and toc BC (`wl`) "
                f"appear to be incompatible:\n`wl - wl_leafsoil`:\n{wl_toc - wl_op}"
            )
        self.nwl = wl_toc.size
        assert p["wl"].size == p["dwl"].size
        p["wle"] = np.r_[p["wl"][0] - 0.5 * p["dwl"][0], p["wl"] + 0.5 * p["dwl"]]
        G_fn = p["G_fn"]
        la = p.get("leaf_angle")
        if la is not None:  # the plugin path uses G_fn, the batched path the parametric family: they must be one canopy
            probe = (0.0, 0.4, 0.9, 1.3)
            if not np.allclose([la.G_fn(a) for a in probe], [G_fn(a) for a in probe], rtol=1e-12, atol=0.0):
                warnings.warn(
                    "`leaf_angle` does not describe the same canopy as `G_fn`; dropping `leaf_angle` "
                    "(the batched path needs `update_p(leaf_angle=...)`, which also sets `G_fn`)."
                )
                del p["leaf_angle"]
        p["K_b_fn"] = lambda psi_: G_fn(psi_) / np.cos(psi_)
        p["G"] = G_fn(psi)
        p["K_b"] = p["K_b_fn"](psi)

    # ------------------------------------------------------------------ single-scenario run (plugin path)
    def run(self, **extra_solver_kwargs):
        """Run the selected scheme on the current parameters (ref model.py:296-320)."""
        self._check_inputs()
        scheme, p = self.scheme, self._p
        sol = scheme["solver"](**{k: p[k] for k in scheme["args"]}, **extra_solver_kwargs)
        self.out.update({k: v for k, v in sol.items() if k in RET_KEYS_ALL_SCHEMES})
        self.out_extra.update({f"{k}_scheme": v for k, v in sol.items() if k not in RET_KEYS_ALL_SCHEMES})
        self._run_count += 1
        return self

    @property
    def out_all(self):
        return {**self.out, **self.out_extra}

    def calc_absorption(self):
        """Layerwise absorption from the stored profiles, on the GPU (ref model.py:327-336, 573-647)."""
        if self._run_count == 0:
            raise Exception("Must run the model first.")
        self.absorption = _calc_absorption(self)
        return self

    # ------------------------------------------------------------------ batched axis
    def scenario_batch(self, **overrides):
        """The current parameters as a one-scenario `ScenarioBatch` (starting point for sweeps)."""
        p = {**self._p, **overrides}
        la = p.get("leaf_angle")
        if la is None:
            raise ValueError(
                "the batched path needs a parametric leaf-angle family: pass `leaf_angle=LeafAngle(...)` "
                "(an arbitrary Python `G_fn` cannot be evaluated inside a CUDA kernel)"
            )
        return ScenarioBatch.from_params(p, leaf_angle=la)

    def run_batch(self, batch, *, bands=("PAR", "NIR"), profiles=True, **solver_options):
        """Run the selected scheme on every scenario of `batch` in one launch.

        Returns a dict of numpy arrays with a leading scenario axis: the scheme's profiles
        `(S, n_z, n_wl)` if `profiles`, and `absorbed` `(S, len(bands))` -- canopy-integrated absorbed
        irradiance in each named band (fused reduction; needs `batch.wl`/`batch.dwl`)."""
        from . import engine

        name = self.scheme["name"]
        band_w = None
        if bands and batch.wl is not None:
            wle = np.r_[batch.wl[0] - 0.5 * batch.dwl[0], batch.wl + 0.5 * batch.dwl]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                band_w = np.stack([x_frac_in_bounds(wle, BAND_DEFNS_UM[b]) for b in bands])
        res = engine.solve(batch, name, band_w=band_w, **solver_options)
        keep = res if profiles else {k: v for k, v in res.items() if k == "absorbed"}
        return {k: v.cpu().numpy() for k, v in keep.items()}

    # ------------------------------------------------------------------ export
    def to_xr(self, *, info=""):
        """Pack inputs/outputs into an `xarray.Dataset` laid out exactly as the reference's (ref model.py:338-447):
        coords z, wl, zm, wle; data variables I_dr, I_df_d, I_df_u, F, I_d, dwl, lai, dlai, the post-processed
        absorption set, the scheme's own `aI*_scheme` arrays, and the scalars psi, sza, G, K_b; attrs info / scheme
        names / version -- so the reference's `diagnostics.band` / `compare_ebal` accept it unchanged.
        xarray is imported here, not at module import: it is not part of this image."""
        try:
            import xarray as xr
        except ImportError as e:
            raise ImportError("Model.to_xr needs xarray, which is not installed") from e
        return xr.Dataset(**self.dataset_description(info=info))

    def dataset_description(self, *, info=""):
        """The keyword arguments of `xarray.Dataset` for this run (`coords`, `data_vars`, `attrs`)."""
        if self._run_count == 0:
            raise Exception("Must run the model before creating the dataset.")
        from . import __version__

        p = self._p
        out = self.out_all
        tup = VMD.dv_tuple
        dv = {
            "I_dr": tup("I_dr", out["I_dr"]), "I_df_d": tup("I_df_d", out["I_df_d"]), "I_df_u": tup("I_df_u", out["I_df_u"]),
            "F": tup("F", out["F"]), "I_d": tup("I_d", out["I_dr"] + out["I_df_d"]),
            "dwl": tup("dwl", p["dwl"]), "lai": tup("lai", p["lai"]), "dlai": tup("dlai", p["dlai"]),
        }
        if self.absorption is not None:
            dv.update({k: tup(k, v) for k, v in self.absorption.items()})
        for k, v in self.out_extra.items():
            if k[:2] != "aI":
                continue
            if v.shape[0] == p["z"].size:  # some schemes provide absorption on interface levels
                dims = ("z", "wl")
            elif v.shape[0] == p["zm"].size:
                dims = ("zm", "wl")
            else:
                raise ValueError("Scheme absorption output has too many or too few levels.")
            base = k[: -len("_scheme")]
            if base not in VMD:
                raise Exception(f"Scheme absorbance variable {base} not found in vmd.")
            dv[k] = (dims, v, da_attrs(VMD[base]))
        dv.update({"psi": tup("psi", p["psi"]), "sza": tup("sza", np.rad2deg(p["psi"])), "G": tup("G", p["G"]),
                   "K_b": tup("K_b", p["K_b"])})
        coords = {"z": tup("z", p["z"]), "wl": tup("wl", p["wl"]), "zm": tup("zm", p["zm"]), "wle": tup("wle", p["wle"])}
        attrs = {"info": info, "scheme_name": self.scheme["name"], "scheme_long_name": self.scheme["long_name"],
                 "scheme_short_name": self.scheme["short_name"], "crt1d_version": __version__}
        return {"coords": coords, "data_vars": dv, "attrs": attrs}


def _calc_absorption(m):
    """Layerwise absorption through the CUDA kernel `crt1d_calc_absorption` (ref model.py:573-647)."""
    import torch

    from . import engine

    p, out = m._p, m.out
    batch = ScenarioBatch.from_params({**p, "leaf_angle": p.get("leaf_angle") or LeafAngle()})
    db = engine.DeviceBatch(batch, "2s", prologue={"K_b": np.array([p["K_b"]], dtype=np.float64)})
    dev = {k: torch.as_tensor(np.ascontiguousarray(out[k])[None]).to(db.device) for k in ("I_dr", "I_df_d", "I_df_u")}
    res = engine.calc_absorption(db, dev["I_dr"], dev["I_df_d"], dev["I_df_u"])
    absorption = {k: res[k][0].cpu().numpy() for k in ABSORPTION_KEYS}
    lai = p["lai"]
    laim = (lai[:-1] + lai[1:]) / 2
    absorption["laim"] = laim
    absorption["f_slm"] = np.exp(-p["K_b"] * laim)
    assert np.allclose(absorption["aI_sl"] + absorption["aI_sh"], absorption["aI"])  # ref model.py:635
    return absorption


_SWEEPABLE = ("psi", "lai", "leaf_r", "leaf_t", "soil_r", "I_dr0_all", "I_df0_all")


def run_sensitivity(m0, p_sets, *, bands=("PAR", "NIR"), profiles=True):
    """Run the cross product of the value lists in `p_sets` as ONE batched GPU solve.

    The reference declares this function but never implemented it (ref model.py:650-664).  `m0` is the
    base case; `p_sets` maps a parameter name (`psi`, `lai`, `leaf_r`, `leaf_t`, `soil_r`, `I_dr0_all`,
    `I_df0_all`) to a list of values.  Returns a dict: every output array gets one leading axis per key
    of `p_sets` (in its order), e.g. `I_df_d` has shape `(len(psi_list), len(lai_list), n_z, n_wl)`;
    `"dims"` lists the swept names.  With xarray installed, `to_dataset(result)` is straightforward --
    it is not imported here."""
    for k in p_sets:
        if k not in _SWEEPABLE:
            raise KeyError(f"cannot sweep {k!r}; sweepable parameters: {', '.join(_SWEEPABLE)}")
    m0._check_inputs()
    p = m0._p
    la = p.get("leaf_angle")
    if la is None:
        raise ValueError("run_sensitivity needs a parametric `leaf_angle` (see Model.scenario_batch)")
    keys = list(p_sets)
    sizes = [len(p_sets[k]) for k in keys]
    vals = {k: list(p_sets[k]) for k in keys}

    def lib(names):
        """Library rows = cross product of the swept members of `names`; returns (rows per name, index fn)."""
        swept = [n for n in names if n in vals]
        combos = list(itertools.product(*[range(len(vals[n])) for n in swept])) or [()]
        rows = {n: [] for n in names}
        for c in combos:
            pick = dict(zip(swept, c))
            for n in names:
                rows[n].append(np.asarray(vals[n][pick[n]] if n in pick else p[n], dtype=np.float64))
        lut = {c: i for i, c in enumerate(combos)}
        return rows, swept, lut

    lai_rows, lai_sw, lai_lut = lib(["lai"])
    leaf_rows, leaf_sw, leaf_lut = lib(["leaf_r", "leaf_t"])
    soil_rows, soil_sw, soil_lut = lib(["soil_r"])
    sky_rows, sky_sw, sky_lut = lib(["I_dr0_all", "I_df0_all"])
    psi, i_lai, i_leaf, i_soil, i_sky = [], [], [], [], []
    for combo in itertools.product(*[range(n) for n in sizes]):
        pick = dict(zip(keys, combo))
        psi.append(vals["psi"][pick["psi"]] if "psi" in pick else p["psi"])
        i_lai.append(lai_lut[tuple(pick[n] for n in lai_sw)])
        i_leaf.append(leaf_lut[tuple(pick[n] for n in leaf_sw)])
        i_soil.append(soil_lut[tuple(pick[n] for n in soil_sw)])
        i_sky.append(sky_lut[tuple(pick[n] for n in sky_sw)])
    batch = ScenarioBatch(
        psi=psi, lai_lib=np.stack(lai_rows["lai"]), leaf_r_lib=np.stack(leaf_rows["leaf_r"]),
        leaf_t_lib=np.stack(leaf_rows["leaf_t"]), soil_r_lib=np.stack(soil_rows["soil_r"]),
        I_dr0_lib=np.stack(sky_rows["I_dr0_all"]), I_df0_lib=np.stack(sky_rows["I_df0_all"]),
        lai_idx=i_lai, leaf_idx=i_leaf, soil_idx=i_soil, sky_idx=i_sky, leaf_angle=la, mla=float(p["mla"]),
        wl=p["wl"], dwl=p["dwl"],
    )
    flat = m0.run_batch(batch, bands=bands, profiles=profiles)
    res = {k: v.reshape(tuple(sizes) + v.shape[1:]) for k, v in flat.items()}
    res["dims"] = keys
    return res


def sensitivity_dataset(res, m0, p_sets, *, bands=("PAR", "NIR")):
    """Describe a `run_sensitivity` result as a dataset: `{"coords": ..., "data_vars": ..., "attrs": ...}` with
    every entry a `(dims, data, attrs)` tuple, i.e. exactly the arguments of `xarray.Dataset` -- "a new
    dimension for each key in `p_sets`" (ref model.py:650-664) in front of the reference's `z`/`zm`/`wl`
    dims (ref model.py:338-447, variables.yml:4-9).  Scalar parameters (psi) become coordinate values;
    array-valued ones (lai, spectra) get an integer case coordinate `<name>_case`.  No xarray needed."""
    p = m0._p
    keys = list(res["dims"])
    nlev = len(p["lai"])

    def tup(name, data, dims):
        m = VMD[name] if name in VMD else None
        return (tuple(dims), data, da_attrs(m) if m else {})

    coords = {"z": tup("z", p["z"], ["z"]), "zm": tup("zm", p["zm"], ["zm"]), "wl": tup("wl", p["wl"], ["wl"])}
    sw_dims = []
    for k in keys:
        vals = [np.asarray(v, dtype=np.float64) for v in p_sets[k]]
        if all(v.ndim == 0 for v in vals):
            sw_dims.append(k)
            coords[k] = tup(k, np.array([float(v) for v in vals]), [k])
        else:
            d = f"{k}_case"
            sw_dims.append(d)
            coords[d] = ((d,), np.arange(len(vals)), {"long_name": f"index into the swept values of {k}"})
    dv = {}
    for k, v in res.items():
        if k == "dims":
            continue
        if k == "absorbed":
            coords["band"] = (("band",), np.array(list(bands)), {"long_name": "spectral band of the absorbed-irradiance diagnostic"})
            dv[k] = (tuple(sw_dims) + ("band",), v, {"units": "W m-2", "long_name": "canopy-absorbed irradiance in band"})
            continue
        if v.ndim != len(sw_dims) + 2:
            dv[k] = (tuple(sw_dims) + tuple(f"{k}_dim{i}" for i in range(v.ndim - len(sw_dims))), v, {})
            continue
        lev = "z" if v.shape[-2] == nlev else "zm"
        dv[k] = tup(k, v, sw_dims + [lev, "wl"])
    attrs = {"scheme_name": m0.scheme["name"], "scheme_long_name": m0.scheme["long_name"],
             "scheme_short_name": m0.scheme["short_name"], "swept": ", ".join(keys)}
    return {"coords": coords, "data_vars": dv, "attrs": attrs}


def sensitivity_to_xr(res, m0, p_sets, *, bands=("PAR", "NIR")):
    """`run_sensitivity` result as the `xr.Dataset` the reference's docstring promises (xarray imported lazily)."""
    try:
        import xarray as xr
    except ImportError as e:
        raise ImportError("sensitivity_to_xr needs xarray, which is not installed") from e
    return xr.Dataset(**sensitivity_dataset(res, m0, p_sets, bands=bands))

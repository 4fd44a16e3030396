"""Variable metadata: names, shapes, units and in/out intent of the solver interface.

The contract is the reference's `crt1d/variables.yml` (names and `intent` flags, e.g. :40-208) as
consumed by `solvers/__init__.py:32,40,95` (argument validation) and `model.py:384-446` (dataset
attributes).  Only the part the solver boundary needs is restated, as a Python table.
"""
from collections import namedtuple

VarMeta = namedtuple("VarMeta", "name shape units long_name intent units_long", defaults=(None,))

_Z, _ZM, _WL, _Z_WL, _ZM_WL = "(n_z,)", "(n_z-1,)", "(n_wl,)", "(n_z, n_wl)", "(n_z-1, n_wl)"
_E = "W m-2"
_LEAF, _PER_GROUND = "W (m2 leaf)-1", "(m2 leaf) (m2 ground area)-1"

# name, shape, units, long_name, intent[, units_long] -- the strings a dataset built by `Model.to_xr` carries must be
# the reference's (they label its plots and key `diagnostics.band`'s PFD conversion on `units == "W m-2"`)
_TABLE = [
    # scheme inputs (intent "in"): exactly the reference's CANOPY_RAD_STATE_INPUT_KEYS
    ("psi", "", "radians", "Solar zenith angle", "in"),
    ("I_dr0_all", _WL, _E, "Incoming direct irradiance at top-of-canopy", "in"),
    ("I_df0_all", _WL, _E, "Incoming diffuse irradiance at top-of-canopy", "in"),
    ("lai", _Z, "m2 m-2", "Leaf area index (cumulative)", "in", _PER_GROUND),
    ("clump", "", "1", "Clump factor", "in"),
    ("leaf_t", _WL, "1", "Leaf transmittance", "in"),
    ("leaf_r", _WL, "1", "Leaf reflectance", "in"),
    ("soil_r", _WL, "1", None, "in"),
    ("K_b", "", "", "Black leaf attenuation coefficient", "in"),
    ("K_b_fn", "", "", None, "in"),
    ("G", "", "", "Fractional leaf area in the psi direction", "in"),
    ("G_fn", "", "", None, "in"),
    ("mla", "", "deg", "Mean leaf angle", "in"),
    # scheme standard outputs (intent "out")
    ("I_dr", _Z_WL, _E, "Direct beam irradiance (binned)", "out"),
    ("I_df_d", _Z_WL, _E, "Downward diffuse irradiance (binned)", "out"),
    ("I_df_u", _Z_WL, _E, "Upward diffuse irradiance (binned)", "out"),
    ("F", _Z_WL, _E, "Actinic flux (binned)", "out"),
    # coordinates and derived quantities used by Model / to_xr
    ("I_d", _Z_WL, _E, "Downward irradiance", "none"),
    ("aI", _ZM_WL, _E, "Absorbed irradiance", "none"),
    ("aI_l", _ZM_WL, _E, "Absorbed irradiance", "none", _LEAF),
    ("aI_dr", _ZM_WL, _E, "Absorbed direct irradiance", "none"),
    ("aI_df", _ZM_WL, _E, "Absorbed diffuse irradiance", "none"),
    ("aI_sl", _ZM_WL, _E, "Absorbed irradiance by sunlit leaves", "none"),
    ("aI_lsl", _ZM_WL, _E, "Absorbed irradiance by sunlit leaves", "none", _LEAF),
    ("aI_sh", _ZM_WL, _E, "Absorbed irradiance by shaded leaves", "none"),
    ("aI_lsh", _ZM_WL, _E, "Absorbed irradiance by shaded leaves", "none", _LEAF),
    ("aI_df_sl", _ZM_WL, _E, "Absorbed diffuse irradiance by sunlit leaves", "none"),
    ("aI_df_lsl", _ZM_WL, _E, "Absorbed irradiance by sunlit leaves", "none", _LEAF),
    ("aI_df_sh", _ZM_WL, _E, "Absorbed diffuse irradiance by shaded leaves", "none"),
    ("aI_df_lsh", _ZM_WL, _E, "Absorbed irradiance by shaded leaves", "none", _LEAF),
    ("dlai", _ZM, "m2 m-2", "Leaf area index in layer", "none", _PER_GROUND),
    ("lad", _ZM, "m2 m-3", "Leaf area density", "none", _PER_GROUND + " m-1"),
    ("wl", _WL, "μm", "Wavelength", "none"),
    ("dwl", _WL, "μm", "Wavelength band width", "none"),
    ("wle", "(n_wl+1,)", "μm", "Wavelength of irradiance band edges", "none"),
    ("z", _Z, "m", "Height above ground", "none"),
    ("zm", _ZM, "m", "Height above ground", "none"),
    ("f_slm", _ZM, "1", "Sunlit leaf fraction", "none"),
    ("laim", _ZM, "m2 m-2", "Leaf area index (cumulative)", "none", _PER_GROUND),
    ("sza", "", "deg", "Solar zenith angle", "none"),
    ("mu", "", "", "cos(psi)", "none"),
]


def dims_of(shape):
    """Dataset dims of a shape string: `n_z-1` -> `zm` (layer midpoints), `n_wl+1` -> `wle` (band edges)
    (ref variables.py:191-210)."""
    if not shape:
        return ()
    special = {"z-1": "zm", "wl+1": "wle"}
    parts = [t.strip() for t in shape.strip("()").split(",") if t.strip()]
    return tuple(special.get(t[2:], t[2:]) for t in parts)


def da_attrs(m):
    """Attributes of a variable's DataArray (ref variables.py:44-59)."""
    attrs = {"long_name": m.long_name, "units": m.units}
    if m.units_long:
        attrs["units_long"] = m.units_long
    return attrs


class _VMD:
    def __init__(self, rows):
        self.variables = {r[0]: VarMeta(*r) for r in rows}

    def __getitem__(self, name):
        return self.variables[name]

    def __contains__(self, name):
        return name in self.variables

    def dv_tuple(self, name, data):
        """`(dims, data, attrs)` as `xarray.Dataset` takes it (ref variables.py:61-63)."""
        m = self.variables[name]
        return (dims_of(m.shape), data, da_attrs(m))

    def intent(self, intent="in"):
        """Variables with the given intent ("in", "out", "none"; None/"all" for everything)."""
        if intent is None or intent == "all":
            return dict(self.variables)
        return {k: v for k, v in self.variables.items() if v.intent == intent}


VMD = _VMD(_TABLE)

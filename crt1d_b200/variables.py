"""Variable metadata: names, shapes, units and in/out intent of the solver interface.

The contract is the reference's `crt1d/variables.yml` (names and `intent` flags, e.g. :40-208) as
consumed by `solvers/__init__.py:32,40,95` (argument validation) and `model.py:384-446` (dataset
attributes).  Only the part the solver boundary needs is restated, as a Python table.
"""
from collections import namedtuple

VarMeta = namedtuple("VarMeta", "name shape units long_name intent")

_Z, _ZM, _WL, _Z_WL, _ZM_WL = "(n_z,)", "(n_z-1,)", "(n_wl,)", "(n_z, n_wl)", "(n_z-1, n_wl)"
_E = "W m-2"

_TABLE = [
    # scheme inputs (intent "in"): exactly the reference's CANOPY_RAD_STATE_INPUT_KEYS
    ("psi", "", "radians", "Solar zenith angle", "in"),
    ("I_dr0_all", _WL, _E, "Incoming direct irradiance at top-of-canopy", "in"),
    ("I_df0_all", _WL, _E, "Incoming diffuse irradiance at top-of-canopy", "in"),
    ("lai", _Z, "m2 m-2", "Cumulative leaf area index", "in"),
    ("clump", "", "", "Clumping index", "in"),
    ("leaf_t", _WL, "", "Leaf element transmittance", "in"),
    ("leaf_r", _WL, "", "Leaf element reflectance", "in"),
    ("soil_r", _WL, "", "Soil reflectivity", "in"),
    ("K_b", "", "", "Black leaf extinction coefficient", "in"),
    ("K_b_fn", "", "", "K_b(psi) function", "in"),
    ("G", "", "", "Leaf angle projection factor", "in"),
    ("G_fn", "", "", "G(psi) function", "in"),
    ("mla", "", "deg", "Mean leaf inclination angle", "in"),
    # scheme standard outputs (intent "out")
    ("I_dr", _Z_WL, _E, "Direct beam irradiance", "out"),
    ("I_df_d", _Z_WL, _E, "Downward diffuse irradiance", "out"),
    ("I_df_u", _Z_WL, _E, "Upward diffuse irradiance", "out"),
    ("F", _Z_WL, _E, "Actinic flux", "out"),
    # coordinates and derived quantities used by Model / to_xr
    ("z", _Z, "m", "Height above ground", "none"),
    ("zm", _ZM, "m", "Layer midpoint height", "none"),
    ("dlai", _ZM, "m2 m-2", "Layer leaf area index", "none"),
    ("wl", _WL, "μm", "Wavelength", "none"),
    ("dwl", _WL, "μm", "Wavelength band width", "none"),
    ("wle", "(n_wl+1,)", "μm", "Wavelength band edges", "none"),
    ("laim", _ZM, "m2 m-2", "Cumulative LAI at layer midpoints", "none"),
    ("f_slm", _ZM, "", "Sunlit fraction at layer midpoints", "none"),
    ("aI", _ZM_WL, _E, "Absorbed irradiance", "none"),
    ("aI_df", _ZM_WL, _E, "Absorbed diffuse irradiance", "none"),
    ("aI_dr", _ZM_WL, _E, "Absorbed direct irradiance", "none"),
    ("aI_sh", _ZM_WL, _E, "Absorbed irradiance by shaded leaves", "none"),
    ("aI_sl", _ZM_WL, _E, "Absorbed irradiance by sunlit leaves", "none"),
    ("aI_df_sl", _ZM_WL, _E, "Absorbed diffuse irradiance by sunlit leaves", "none"),
    ("aI_df_sh", _ZM_WL, _E, "Absorbed diffuse irradiance by shaded leaves", "none"),
    ("aI_l", _Z_WL, "W (m2 leaf)-1", "Absorbed irradiance per unit leaf area", "none"),
    ("aI_lsl", _Z_WL, "W (m2 leaf)-1", "Absorbed irradiance per unit sunlit leaf area", "none"),
    ("aI_lsh", _Z_WL, "W (m2 leaf)-1", "Absorbed irradiance per unit shaded leaf area", "none"),
]


class _VMD:
    def __init__(self, rows):
        self.variables = {r[0]: VarMeta(*r) for r in rows}

    def __getitem__(self, name):
        return self.variables[name]

    def __contains__(self, name):
        return name in self.variables

    def intent(self, intent="in"):
        """Variables with the given intent ("in", "out", "none"; None/"all" for everything)."""
        if intent is None or intent == "all":
            return dict(self.variables)
        return {k: v for k, v in self.variables.items() if v.intent == intent}


VMD = _VMD(_TABLE)

"""Host-side description of a batch of canopy scenarios (the batched axis `Model.run` gains).

A *scenario* is one call of a reference solver: one solar zenith angle, one cumulative-LAI profile, one
set of leaf / soil spectra and one top-of-canopy direct/diffuse irradiance spectrum.  Sweeps are cross
products of a few hundred distinct profiles and spectra, so the batch stores small *libraries* (rows
of float64) plus per-scenario int32 row indices -- 24 bytes per scenario instead of 5 x n_wl doubles.
The layout is exactly what the C ABI takes (`crt1d_batch` in include/crt1d_b200.h).
"""
from dataclasses import dataclass
from dataclasses import field

import numpy as np

from .leaf_angle import LeafAngle


def _rows(a, name):
    a = np.ascontiguousarray(np.atleast_2d(np.asarray(a, dtype=np.float64)))
    if a.ndim != 2:
        raise ValueError(f"{name} must be 1-D or 2-D")
    return a


@dataclass
class ScenarioBatch:
    """S scenarios over shared libraries.  All libraries are C-order float64, band (or level) fastest."""

    psi: np.ndarray  # (S,) solar zenith angle, radians
    lai_lib: np.ndarray  # (n_lai, n_z) cumulative LAI, index 0 = ground (total), -1 = top (0)
    leaf_r_lib: np.ndarray  # (n_leaf, n_wl)
    leaf_t_lib: np.ndarray  # (n_leaf, n_wl)
    soil_r_lib: np.ndarray  # (n_soil, n_wl)
    I_dr0_lib: np.ndarray  # (n_sky, n_wl) in-band W m-2
    I_df0_lib: np.ndarray  # (n_sky, n_wl)
    lai_idx: np.ndarray  # (S,) int32 rows of lai_lib
    leaf_idx: np.ndarray  # (S,) int32
    soil_idx: np.ndarray  # (S,) int32
    sky_idx: np.ndarray  # (S,) int32
    leaf_angle: LeafAngle = field(default_factory=LeafAngle)
    mla: float = 57.0  # mean leaf angle (deg), 2s only
    wl: np.ndarray = None  # (n_wl,) band centres, micrometres (for PAR/NIR weights)
    dwl: np.ndarray = None  # (n_wl,) band widths

    def __post_init__(self):
        self.psi = np.ascontiguousarray(np.atleast_1d(np.asarray(self.psi, dtype=np.float64)))
        for k in ("lai_lib", "leaf_r_lib", "leaf_t_lib", "soil_r_lib", "I_dr0_lib", "I_df0_lib"):
            setattr(self, k, _rows(getattr(self, k), k))
        S = self.psi.size
        for k in ("lai_idx", "leaf_idx", "soil_idx", "sky_idx"):
            v = np.array(np.broadcast_to(np.asarray(getattr(self, k), dtype=np.int32), (S,)), dtype=np.int32, order="C")
            setattr(self, k, v)
        n_wl = self.leaf_r_lib.shape[1]
        for k in ("leaf_t_lib", "soil_r_lib", "I_dr0_lib", "I_df0_lib"):
            if getattr(self, k).shape[1] != n_wl:
                raise ValueError(f"{k} has {getattr(self, k).shape[1]} bands, expected {n_wl}")
        if self.leaf_t_lib.shape[0] != self.leaf_r_lib.shape[0]:
            raise ValueError("leaf_r_lib and leaf_t_lib need the same number of rows")
        if self.I_dr0_lib.shape[0] != self.I_df0_lib.shape[0]:
            raise ValueError("I_dr0_lib and I_df0_lib need the same number of rows")
        for k, lib in (("lai_idx", self.lai_lib), ("leaf_idx", self.leaf_r_lib),
                       ("soil_idx", self.soil_r_lib), ("sky_idx", self.I_dr0_lib)):
            v = getattr(self, k)
            if v.size and (v.min() < 0 or v.max() >= lib.shape[0]):
                raise IndexError(f"{k} out of range for a library of {lib.shape[0]} rows")
        # same structural checks as Model._check_inputs (ref model.py:240-246) on every profile
        L = self.lai_lib
        if L.shape[1] < 2 or not (np.all(L[:, 0] > L[:, -1]) and np.all(L[:, -1] == 0)):
            raise AssertionError("each LAI profile must decrease from lai[0] = total to lai[-1] == 0")

    @property
    def n_scen(self):
        return self.psi.size

    @property
    def n_z(self):
        return self.lai_lib.shape[1]

    @property
    def n_wl(self):
        return self.leaf_r_lib.shape[1]

    def scenario_params(self, s):
        """The reference-style parameter dict of scenario `s` (what one `solve_<id>` call receives)."""
        la = self.leaf_angle
        d = dict(
            psi=float(self.psi[s]),
            lai=self.lai_lib[self.lai_idx[s]].copy(),
            leaf_r=self.leaf_r_lib[self.leaf_idx[s]].copy(),
            leaf_t=self.leaf_t_lib[self.leaf_idx[s]].copy(),
            soil_r=self.soil_r_lib[self.soil_idx[s]].copy(),
            I_dr0_all=self.I_dr0_lib[self.sky_idx[s]].copy(),
            I_df0_all=self.I_df0_lib[self.sky_idx[s]].copy(),
            mla=self.mla, clump=1.0, G_fn=la.G_fn, K_b_fn=la.K_b_fn, leaf_angle=la,
        )
        d["G"] = la.G_fn(d["psi"])
        d["K_b"] = la.K_b_fn(d["psi"])
        if self.wl is not None:
            d["wl"] = self.wl
            d["dwl"] = self.dwl
        return d

    def slice(self, lo, hi):
        """Scenarios [lo, hi) over the same libraries (no copies of the libraries)."""
        import copy

        out = copy.copy(self)
        out.psi = self.psi[lo:hi]
        for k in ("lai_idx", "leaf_idx", "soil_idx", "sky_idx"):
            setattr(out, k, getattr(self, k)[lo:hi])
        return out

    def take(self, idx):
        """The scenarios `idx` (any index array, in that order) over the same libraries."""
        import copy

        idx = np.asarray(idx, dtype=np.int64)
        if idx.ndim != 1 or (idx.size and (idx.min() < 0 or idx.max() >= self.n_scen)):
            raise ValueError("idx must be a 1-D array of scenario indices in range")
        out = copy.copy(self)
        out.psi = np.ascontiguousarray(self.psi[idx])
        for k in ("lai_idx", "leaf_idx", "soil_idx", "sky_idx"):
            setattr(out, k, np.ascontiguousarray(getattr(self, k)[idx]))
        return out

    def permuted(self, order):
        """The same scenarios in another order (`order[i]` = index of the scenario that comes i-th) over the same
        libraries -- a batch is index arrays, so this costs 20 bytes per scenario."""
        import copy

        order = np.asarray(order, dtype=np.int64)
        if order.shape != (self.n_scen,) or not np.array_equal(np.sort(order), np.arange(self.n_scen)):
            raise ValueError("order must be a permutation of range(n_scen)")
        out = copy.copy(self)
        out.psi = np.ascontiguousarray(self.psi[order])
        for k in ("lai_idx", "leaf_idx", "soil_idx", "sky_idx"):
            setattr(out, k, np.ascontiguousarray(getattr(self, k)[order]))
        return out

    @classmethod
    def from_params(cls, p, leaf_angle=None):
        """A one-scenario batch from a reference-style parameter dict."""
        return cls(
            psi=[p["psi"]], lai_lib=p["lai"], leaf_r_lib=p["leaf_r"], leaf_t_lib=p["leaf_t"],
            soil_r_lib=p.get("soil_r", np.zeros_like(p["leaf_r"])), I_dr0_lib=p["I_dr0_all"],
            I_df0_lib=p["I_df0_all"], lai_idx=[0], leaf_idx=[0], soil_idx=[0], sky_idx=[0],
            leaf_angle=leaf_angle or p.get("leaf_angle") or LeafAngle(), mla=float(p.get("mla", 57.0)),
            wl=p.get("wl"), dwl=p.get("dwl"),
        )

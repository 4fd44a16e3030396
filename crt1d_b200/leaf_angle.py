"""Host-side leaf-angle projection functions G(psi) and the parametric family ids the CUDA kernels know.

Mirrors the *interface* of the reference's `crt1d/leaf_angle.py:118-202` (same function names and
argument meaning) so a `G_fn` built from these behaves identically in the plugin path, where `G_fn` /
`K_b_fn` are arbitrary Python callables evaluated on the host.  The batched path
cannot call Python from a kernel; it uses the closed parametric families below, identified by
`LeafAngle(family, param)` and evaluated on the device by `csrc/crt_leafangle.cu`.
"""
import math
from dataclasses import dataclass

import numpy as np

# family ids shared with include/crt1d_b200.h (CRT1D_G_*)
G_FAMILY_IDS = {
    "spherical": 0,
    "horizontal": 1,
    "vertical": 2,
    "ellipsoidal_approx": 3,
    "ellipsoidal": 4,
    "ellipsoidal_approx_bonan": 5,
}


def G_horizontal(psi):
    """Horizontal leaves: projection is cos(psi)  (ref leaf_angle.py:118-120)."""
    return np.cos(psi)


def G_spherical(psi):
    """Spherical distribution: 1/2 independent of psi  (ref leaf_angle.py:123-125)."""
    return 0.5


def G_vertical(psi):
    """Vertical leaves: (2/pi) sin(psi)  (ref leaf_angle.py:128-130)."""
    return 2.0 / math.pi * np.sin(psi)


def G_ellipsoidal(psi, x):
    """Campbell (1986) exact ellipsoidal G  (ref leaf_angle.py:133-165)."""
    if x == 1:
        out = np.full_like(psi, 0.5, dtype=float)
        return float(out) if out.size == 1 else out
    elev = math.pi / 2 - psi
    num = np.sqrt(x**2 + 1.0 / (np.tan(elev) ** 2))
    if x > 1:
        e1 = np.sqrt(1 - x**-2)
        den = x + 1.0 / (2 * e1 * x) * np.log((1 + e1) / (1 - e1))
    else:
        e2 = np.sqrt(1 - x**2)
        den = x + np.arcsin(e2) / e2
    return num / den * np.cos(psi)


def G_ellipsoidal_approx(psi, x):
    """Campbell (1990) approximate ellipsoidal G -- the default-case G_fn  (ref leaf_angle.py:168-180)."""
    num = np.sqrt(x**2 + np.tan(psi) ** 2)
    den = x + 1.774 * (x + 1.182) ** -0.733
    return num / den * np.cos(psi)


def G_ellipsoidal_approx_bonan(psi, xl):
    """Ross-Goudriaan form used by Bonan; `xl` is chi_l, clipped to [-0.4, 0.6]  (ref leaf_angle.py:183-202)."""
    chil = min(max(xl, -0.4), 0.6)
    phi1 = 0.5 - 0.633 * chil - 0.330 * chil**2
    phi2 = 0.877 * (1 - 2 * phi1)
    return phi1 + phi2 * np.cos(psi)


def mla_to_x_approx(mla):
    """Mean leaf angle (deg) -> ellipsoidal x, Campbell (1990) eq. 16 inverted  (ref leaf_angle.py:222-230)."""
    x = (np.deg2rad(mla) / 9.65) ** (-1.0 / 1.65) - 3.0
    assert x > 0
    return x


def x_to_mla_approx(x):
    """Ellipsoidal x -> mean leaf angle (deg)  (ref leaf_angle.py:205-211)."""
    return np.rad2deg(9.65 * (3 + x) ** (-1.65))


_HOST_FNS = {
    "spherical": lambda psi, p: G_spherical(psi) + 0.0 * np.asarray(psi),
    "horizontal": lambda psi, p: G_horizontal(psi),
    "vertical": lambda psi, p: G_vertical(psi),
    "ellipsoidal_approx": G_ellipsoidal_approx,
    "ellipsoidal": G_ellipsoidal,
    "ellipsoidal_approx_bonan": G_ellipsoidal_approx_bonan,
}


@dataclass(frozen=True)
class LeafAngle:
    """A parametric leaf-angle family the device kernels can evaluate: `family` name + one parameter
    (`x` for the ellipsoidal forms, `chi_l` for the Bonan form, unused otherwise)."""

    family: str = "ellipsoidal_approx"
    param: float = 1.0

    def __post_init__(self):
        if self.family not in G_FAMILY_IDS:
            raise ValueError(f"unknown leaf-angle family {self.family!r}; valid: {sorted(G_FAMILY_IDS)}")

    @property
    def family_id(self):
        return G_FAMILY_IDS[self.family]

    def G_fn(self, psi):
        """Host evaluation (same arithmetic as the reference functions above)."""
        return _HOST_FNS[self.family](psi, self.param)

    def K_b_fn(self, psi):
        """K_b = G/cos(psi)  (ref model.py:291)."""
        return self.G_fn(psi) / np.cos(psi)

    @classmethod
    def from_mla(cls, mla):
        return cls("ellipsoidal_approx", float(mla_to_x_approx(mla)))

"""Cumulative-LAI profile generators used to set up cases (host side, O(n_z) once per canopy).

Interface follows the reference's `crt1d/leaf_area.py:42-146`: index 0 = ground (`lai[0]` = total LAI),
index -1 = canopy top (`lai[-1] == 0`); `n` counts interface LEVELS.
"""
from collections import namedtuple

import numpy as np
from scipy.stats import beta as _beta

LeafAreaProfile = namedtuple("LeafAreaProfile", "lai lad z")


def distribute_lai_beta(h_c, LAI, n, *, h_min=0.5):
    """Beta-distributed leaf area density with its mode at 0.7 h_c  (ref leaf_area.py:42-93).

    Cumulative LAI is uniform in level index (`linspace(1, 0, n) * LAI`); only z is non-uniform.
    """
    mode_depth = (h_c - 0.7 * h_c) / h_c
    b = 3
    a = -((b - 2) * mode_depth + 1) / (mode_depth - 1)
    frac = np.linspace(1.0, 0, n)
    z = (h_c - h_min) * (1 - _beta(a, b).ppf(frac)) + h_min
    lai = frac * LAI
    zrel = (z - h_min) / (h_c - h_min)
    lad = LAI / (h_c - h_min) * _beta.pdf(zrel, b, a)
    return LeafAreaProfile(lai, lad, z)


def distribute_lai_beta_bonan(h_c, LAI, n, *, h_min=0.5, p=3.5, q=2.0):
    """Bonan (2019) SP 2.1 beta profile, used by the n79 known-answer case  (ref leaf_area.py:101-146)."""
    dist = _beta(p, q)
    h = h_c - h_min
    frac = np.linspace(1.0, 0, n)
    z = (h * dist.ppf(frac) + h_min)[::-1]
    lai = frac * LAI
    lad = LAI / h * dist.pdf((z - h_min) / h)
    return LeafAreaProfile(lai, lad, z)

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


# ----------------------------------------------------------------------------------------------------------------
# Profiles that are NOT `linspace * LAI`  (ref leaf_area.py:156-342, 395-641).  All host-side, O(n) once per canopy;
# they feed `ScenarioBatch.lai_lib`.  Only `lai` reaches the solvers -- and for `distribute_lai_weibull_z` and
# `distribute_lai_gamma` it is not equally spaced (weibull_z even has zero-thickness layers below the crown base),
# which is what the kernels' direct-exponential paths and the parity tests on these axes are for.
# ----------------------------------------------------------------------------------------------------------------
WEIBULL_SPECIES = {"pine": (0.906, 2.145), "spruce": (2.375, 1.289), "birch": (0.557, 1.914)}  # Teske & Thistle (2004)


def _lai_above(lad, z):
    """Cumulative LAI above each level by the trapezoid rule, accumulated from the top down, in the operation order
    of `-cumtrapz(lad[::-1], z[::-1], initial=0)[::-1]` (ref leaf_area.py:216) so the result is bit-identical."""
    yr, xr = lad[::-1], z[::-1]
    acc = np.concatenate(([0.0], np.cumsum(np.diff(xr) * (yr[1:] + yr[:-1]) / 2.0)))
    return -1 * acc[::-1]


def distribute_lai_weibull_z(z, LAI, h, hb=0.0, *, b=None, c=None, species=None):
    """Weibull leaf-area density between crown base `hb` and canopy height `h` on a GIVEN equally spaced height grid
    `z` (which must extend above `h`); cumulative LAI by trapezoid integration of the density, so it is neither
    equally spaced nor strictly decreasing (ref leaf_area.py:156-218)."""
    z = np.array(z, dtype=np.float64)
    if z.max() <= h or h <= hb:
        raise ValueError("h must be lower than uppermost gridpoint")
    if b is None or c is None:
        b, c = WEIBULL_SPECIES[species]
    dz = abs(z[1] - z[0])
    crown = (z > hb) & (z <= h)
    u = 1.0 - np.linspace(0, 1, int(crown.sum()))  # depth below the crown top, normalised
    shape = np.zeros(z.size)
    shape[crown] = -(c / b) * (u / b) ** (c - 1.0) * np.exp(-((u / b) ** c)) / (1.0 - np.exp(-((1.0 / b) ** c)))
    shape = np.abs(shape / sum(shape * dz))  # builtin sum: the reference's left-to-right accumulation
    lad = LAI * shape
    return LeafAreaProfile(_lai_above(lad, z), lad, z)


def _interp_open(xq, xp, fp):
    """Linear interpolation on an ascending or descending abscissa `xp`, NaN outside its range: what
    `scipy.interpolate.interp1d(..., bounds_error=False)` does for 1-D float data (ref leaf_area.py:265-271) -- a
    stable sort of the abscissa, `numpy.interp` on it (so, of tied abscissae, the later one wins), NaN fill."""
    xp, fp, xq = np.asarray(xp, float), np.asarray(fp, float), np.asarray(xq, float)
    order = np.argsort(xp, kind="mergesort")
    xs, fs = xp[order], fp[order]
    out = np.interp(xq, xs, fs)
    out[(xq < xs[0]) | (xq > xs[-1])] = np.nan
    return out


def distribute_lai_weibull(h_c, LAI, n, *, h_min=0.5, b=None, c=None, species=None):
    """Equal-LAI-increment levels whose heights follow the Weibull profile: `distribute_lai_weibull_z` on
    `linspace(0, h_c + 0.5, n)`, then heights and density interpolated to `linspace(LAI, 0, n)`
    (ref leaf_area.py:229-284).  `lai` IS equally spaced here; z and lad are not mutually consistent (as upstream)."""
    z0 = np.linspace(0, h_c + 0.5, n)
    base = distribute_lai_weibull_z(z0, LAI, h_c, hb=h_min, b=b, c=c, species=species)
    lai = np.linspace(LAI, 0, n)
    above = z0 >= h_min
    z = np.r_[h_min, _interp_open(lai[1:], base.lai[above], z0[above])]
    lad = np.r_[0, _interp_open(z[1:], z0, base.lad)]
    k = LAI / lai[0]
    return LeafAreaProfile(lai * k, lad * k, z)


def distribute_lai_gamma(h_c, LAI, n):
    """Levels at fixed cumulative-LAI fractions `[1, 0.99 ... 0.1, 0]` whose depths are quantiles of a gamma law with
    shape `0.3 h_c / 3.5 + 1` SHIFTED by 3.5 m (the reference passes 3.5 as scipy's `loc`, ref leaf_area.py:320-342);
    no density is returned."""
    from scipy.special import gammaincinv

    shift = 3.5
    shape = (h_c - 0.7 * h_c) / shift + 1
    frac = np.linspace(0.99, 0.1, n - 2)
    lai = np.zeros(n)
    z = np.zeros(n)
    lai[0], lai[1:-1] = LAI, frac * LAI
    z[1:-1] = h_c - (gammaincinv(shape, frac) + shift)
    z[-1] = h_c
    return LeafAreaProfile(lai, None, z)


class Storey:
    """One storey of a multi-storey canopy between `h1` (density `lad_h1`) and `h2` (density `lad_h2`) holding `LAI`,
    with the density maximum at `hmax`: a sine arc `sin(1.3 s)` below the maximum and the parabola `1 - x^2` above
    (ref leaf_area.py:395-498, class `layer`).  The reference finds the amplitude with `fsolve` over `quad`; the
    density is affine in the amplitude, so here it is solved exactly, and `lai_above` is the closed-form integral."""

    _ARC = 1.3

    def __init__(self, h1, lad_h1, hmax, LAI, h2, lad_h2):
        self.h1, self.lad_h1, self.hmax, self.LAI, self.h2, self.lad_h2 = h1, lad_h1, hmax, LAI, h2, lad_h2
        w_lo, w_up = hmax - h1, h2 - hmax
        s, c = np.sin(self._ARC), np.cos(self._ARC)
        # LAI(m) = m * [w_lo (1 - cos 1.3) / 1.3 + 2/3 w_up sin 1.3] + lad_h1 (w_lo + 2/3 w_up) + lad_h2 w_up / 3
        # with m = amplitude / (sin 1.3 + lad_h1) the factor in front of the arc
        gain = w_lo * (1 - c) / self._ARC + (2.0 / 3) * w_up * s
        fixed = lad_h1 * (w_lo + (2.0 / 3) * w_up) + lad_h2 * w_up / 3.0
        self.m = (LAI - fixed) / gain
        self.lai_mult = self.m * (s + lad_h1)  # the reference's `lai_mult`
        self.peak = self.m * s + lad_h1  # density at hmax
        if self.lai_mult <= 0:
            print("desired LAI too small")

    def pdf(self, h):
        """Leaf area density at height h (NaN outside [h1, h2))."""
        if h >= self.h2 or h < self.h1:
            return np.nan
        if h >= self.hmax:
            x = (h - self.hmax) / (self.h2 - self.hmax)
            return (1 - x * x) * (self.peak - self.lad_h2) + self.lad_h2
        return self.m * np.sin(self._ARC * (h - self.h1) / (self.hmax - self.h1)) + self.lad_h1

    def _upper_above(self, h):
        x = (h - self.hmax) / (self.h2 - self.hmax)
        a, bb, w = self.peak - self.lad_h2, self.lad_h2, self.h2 - self.hmax
        return w * ((a + bb) * (1 - x) - a * (1 - x ** 3) / 3.0)

    def cdf(self, h):
        """LAI of this storey above height h."""
        if h >= self.h2:
            return 0.0
        if h >= self.hmax:
            return self._upper_above(h)
        if h >= self.h1:
            w = self.hmax - self.h1
            s = self._ARC * (h - self.h1) / w
            lower = self.m * w / self._ARC * (np.cos(s) - np.cos(self._ARC)) + self.lad_h1 * (self.hmax - h)
            return lower + self._upper_above(self.hmax)
        return self.LAI


class CanopyLaiDist:
    """Stack of storeys, bottom-most first (ref leaf_area.py:501-586, class `canopy_lai_dist`); `layers` are dicts
    with `h_max`, `h_top`, `lad_h_top`, `fLAI`."""

    def __init__(self, h_bottom, layers, LAI):
        self.h_bottom, self.LAItot = h_bottom, LAI
        self.storeys = []
        h1, lad1 = h_bottom, 0
        for ld in layers:
            self.storeys.append(Storey(h1, lad1, ld["h_max"], ld["fLAI"] * LAI, ld["h_top"], ld["lad_h_top"]))
            h1, lad1 = ld["h_top"], ld["lad_h_top"]
        self.h_tops = np.array([s.h2 for s in self.storeys])
        self.LAIlayers = np.array([s.LAI for s in self.storeys])

    def _which(self, h):
        return int(np.where(self.h_tops >= h)[0].min())

    def pdf(self, h):
        if h > self.h_tops.max() or h < self.h_bottom:
            return 0
        return self.storeys[self._which(h)].pdf(h)

    def cdf(self, h):
        """Canopy LAI above height h."""
        if h > self.h_tops.max():
            return 0
        k = self._which(h)
        return self.storeys[k].cdf(h) + float(np.sum(self.LAIlayers[k + 1:]))

    def height_of(self, lai_above):
        """The height above which the canopy holds `lai_above` (exact inverse of `cdf`, bracketing root search)."""
        from scipy.optimize import brentq

        top = float(self.h_tops.max())
        if lai_above <= 0:
            return top
        if lai_above >= self.cdf(self.h_bottom):
            return float(self.h_bottom)
        return brentq(lambda h: self.cdf(h) - lai_above, self.h_bottom, top, xtol=1e-13, rtol=1e-14)

    def inv_cdf(self, ub, lai):
        """Lower bound below `ub` that encloses `lai` (the reference integrates the density numerically and uses
        `fsolve`, ref leaf_area.py:577-586; here through the closed-form cdf)."""
        return np.array([self.height_of(self.cdf(ub) + lai)])


def distribute_lai_from_cdd(cdd, n):
    """Equal-LAI-increment levels of a multi-storey canopy described by a canopy-description dict (keys as in
    `cases.load_canopy_descrip` / the reference's default CSV; ref leaf_area.py:589-641).  Heights agree with the
    reference to its `fsolve`/`quad` accuracy (~1e-5 m; worst at the canopy bottom where the density vanishes);
    `lai` is bit-identical, including the reference's round-off-dependent treatment of the lowest level."""
    LAI = cdd["lai_tot"]
    h_bottom = cdd["h_bot"][-1]
    layers = [dict(h_max=cdd["h_max_lad"][i], h_top=cdd["h_top"][i], lad_h_top=cdd["lad_h_top"][i], fLAI=cdd["lai_frac"][i])
              for i in range(len(cdd["lai_frac"]))]
    cld = CanopyLaiDist(h_bottom, layers[::-1], LAI)
    h_canopy = cdd["h_canopy"]
    dlai = float(LAI) / (n - 1)
    lai = np.zeros(n)
    z = h_canopy * np.ones(n)
    got = 0
    for i in range(n - 2, -1, -1):  # top down; `got` accumulates in floating point exactly like the reference's LAIcum
        if LAI - got < dlai:
            assert i == 0
            z[0], lai[0] = h_bottom, LAI
        else:
            lai[i] = lai[i + 1] + dlai
            got += dlai
            z[i] = cld.height_of(lai[i])
    return LeafAreaProfile(lai, None, z)

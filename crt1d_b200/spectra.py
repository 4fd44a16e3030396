"""Band definitions and the fractional-overlap band weights used by the fused PAR/NIR reduction.

Only the piece of the reference's `crt1d/spectra.py` that is on the hot path's epilogue:
`BAND_DEFNS_UM` (ref spectra.py:22-27) and `_x_frac_in_bounds` (ref spectra.py:71-126), which
`diagnostics.band` (ref diagnostics.py:56-81) multiplies into a Sum over wavelength bands.
"""
import warnings

import numpy as np

BAND_DEFNS_UM = {
    "PAR": (0.4, 0.7),
    "NIR": (0.7, 2.5),
    "UV": (0.01, 0.4),
    "solar": (0.3, 5.0),
}


def x_frac_in_bounds(xe, bounds):
    """Weight in [0, 1] of every bin (edges `xe`) that falls inside `bounds`; partial bins get their
    overlapped fraction.  Vectorised restatement of ref spectra.py:71-126 (identical results,
    including the `>=`/`<=` edge-touching rule that yields weight 0 for a bin that only touches)."""
    xe = np.asarray(xe, dtype=float)
    lo, hi = xe[:-1], xe[1:]
    b1, b2 = bounds
    if (b1 < lo[0] or b2 > hi[-1]) and tuple(bounds) != BAND_DEFNS_UM["solar"]:
        warnings.warn(
            f"`bounds` ({b1:.3g}, {b2:.3g}) extend outside the data range "
            f"defined by `xe` ({lo[0]:.3g}, {hi[-1]:.3g})"
        )
    inside = (hi >= b1) & (lo <= b2)
    width = hi - lo
    w = np.ones_like(lo)
    left = lo < b1
    right = (~left) & (hi > b2)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(left, (hi - b1) / width, w)
        w = np.where(right, (b2 - lo) / width, w)
    return np.where(inside, w, 0.0)


_x_frac_in_bounds = x_frac_in_bounds  # the reference's private name


def band_weights(wle, band_name):
    """Weights for a named band ("PAR", "NIR", ...) on bins with edges `wle` (micrometres)."""
    return x_frac_in_bounds(wle, BAND_DEFNS_UM[band_name])


def edges_from_centers_widths(wl, dwl):
    """Band edges as `Model._check_inputs` builds them  (ref model.py:286-287)."""
    wl = np.asarray(wl, dtype=float)
    dwl = np.asarray(dwl, dtype=float)
    return np.r_[wl[0] - 0.5 * dwl[0], wl + 0.5 * dwl]


def e_wl_umol(wl_um):
    """Photon energy, J per micromole of photons, at wavelength `wl_um` (ref spectra.py:30-39).  Dividing
    band weights by it turns the fused W m-2 reductions into photon flux density (umol m-2 s-1), which --
    as the reference notes (diagnostics.py:49-51) -- has to happen per wavelength, before integrating."""
    h, c, N_A = 6.62607015e-34, 299792458.0, 6.02214076e23
    return h * c / (np.asarray(wl_um, dtype=float) * 1e-6) * N_A * 1e-6


def pfd_band_weights(wl, dwl, band_name="PAR"):
    """Band weights that integrate irradiance to photon flux density in a named band."""
    return band_weights(edges_from_centers_widths(wl, dwl), band_name) / e_wl_umol(wl)


def smear_tuv(x, y, bins, *, device=None):
    r"""Smear `y`\(`x`) into `bins` with the TUV method: each value is the trapezoidally integrated average of
    y(x) in its bin (ref spectra.py:261-300, `_smear_tuv_1` :221-258).  Same arguments as the reference, plus:
    `y` may be 2-D `(n_rows, n_x)` -- a whole spectra library is binned in one CUDA launch
    (`crt1d_smear_tuv`).  Returns float64 `(n_bins,)` or `(n_rows, n_bins)`."""
    import torch

    from . import _lib

    x = np.ascontiguousarray(x, dtype=np.float64)
    y2 = np.ascontiguousarray(np.atleast_2d(np.asarray(y, dtype=np.float64)))
    bins = np.ascontiguousarray(bins, dtype=np.float64)
    if x.ndim != 1 or y2.shape[1] != x.size:
        raise ValueError("`x` and `y` must be the same size along the spectral axis")
    if bins.ndim != 1 or bins.size < 2:
        raise ValueError("`bins` needs at least two edges")
    lib = _lib.load()
    dev = torch.device(device if device is not None else "cuda")
    xd, yd, bd = (torch.as_tensor(a).to(dev) for a in (x, y2, bins))
    out = torch.empty((y2.shape[0], bins.size - 1), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.crt1d_smear_tuv(y2.shape[0], x.size, xd.data_ptr(), yd.data_ptr(), bins.size - 1, bd.data_ptr(),
                                       out.data_ptr(), st))
    res = out.cpu().numpy()
    return res[0] if np.ndim(y) == 1 else res


def rebin_batch(batch, wle_new, *, device=None):
    """A copy of `batch` on coarser (or shifted) spectral bins with edges `wle_new` (micrometres): leaf and soil
    optical properties are bin averages; the in-band irradiances are converted to spectral irradiance at the
    band centres, smeared, and multiplied by the new band widths, as the reference's `smear_si` does
    (ref spectra.py:529-573).  Sweeps over band resolution then reuse one set of 1-nm libraries."""
    import copy

    if batch.wl is None or batch.dwl is None:
        raise ValueError("rebin_batch needs batch.wl and batch.dwl")
    wle_new = np.asarray(wle_new, dtype=np.float64)
    dwl_new = np.diff(wle_new)
    names = ("leaf_r_lib", "leaf_t_lib", "soil_r_lib")
    rows = [getattr(batch, k) for k in names] + [batch.I_dr0_lib / batch.dwl, batch.I_df0_lib / batch.dwl]
    sizes = [r.shape[0] for r in rows]
    sm = smear_tuv(batch.wl, np.concatenate(rows, axis=0), wle_new, device=device)
    parts = np.split(sm, np.cumsum(sizes)[:-1], axis=0)
    out = copy.copy(batch)
    for k, v in zip(names, parts[:3]):
        setattr(out, k, np.ascontiguousarray(v))
    out.I_dr0_lib = np.ascontiguousarray(parts[3] * dwl_new)
    out.I_df0_lib = np.ascontiguousarray(parts[4] * dwl_new)
    out.wl = 0.5 * (wle_new[:-1] + wle_new[1:])
    out.dwl = dwl_new
    return out

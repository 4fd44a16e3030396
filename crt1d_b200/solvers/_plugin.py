"""Shared body of the plugin-path solver functions `solve_<id>` (one scenario per call).

Flow: validate like the reference -> host prologue with the caller's Python callables (same scipy
calls as the reference solver's own prologue) -> ONE C-ABI call `crt1d_solve_host` with host buffers
(H2D, the sm_100a kernel, D2H inside) -> fresh float64 `(n_z, n_wl)` numpy arrays in the reference's
return-dict layout.  No CPU fallback: a missing library or GPU raises.
"""
import ctypes
import os

import numpy as np

from .. import _abi
from .. import _lib
from ..engine import EXTRA_NAMES
from ..engine import _abi_array
from ..engine import host_prologue
from ..scenarios import ScenarioBatch


def _device_index():
    return int(os.environ.get("CRT1D_B200_DEVICE", "0"))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def solve_batch_host(batch, scheme, prologue, *, mu_s=0.501, band_w=None, device=None, alloc=np.empty, status=False,
                     fields=None):
    """`crt1d_solve_host` on a ScenarioBatch with a host prologue dict; returns numpy arrays with a
    leading scenario axis.  `alloc(shape)` makes the float64 output arrays (default: fresh pageable numpy
    arrays like the reference's; a page-locked allocator lets the D2H copies run at PCIe speed).
    `status=True` adds the per-scenario int32 status words (`_abi.STATUS_NONFINITE`) as `out["status"]`.
    `fields`: names of the profiles to return (default all of the scheme's); the others are neither computed into
    host memory nor copied (NULL pointers in `crt1d_out`)."""
    lib = _lib.load()
    S, nz, nw = batch.n_scen, batch.n_z, batch.n_wl
    keep = {}
    cb = _abi.Batch()
    cb.n_scen, cb.n_z, cb.n_wl = S, nz, nw
    cb.n_lai, cb.n_leaf = batch.lai_lib.shape[0], batch.leaf_r_lib.shape[0]
    cb.n_soil, cb.n_sky = batch.soil_r_lib.shape[0], batch.I_dr0_lib.shape[0]
    for k in ("psi", "lai_lib", "leaf_r_lib", "leaf_t_lib", "soil_r_lib", "I_dr0_lib", "I_df0_lib",
              "lai_idx", "leaf_idx", "soil_idx", "sky_idx"):
        keep[k] = _abi_array(batch, k)
        setattr(cb, k, _ptr(keep[k]))
    for k, v in prologue.items():
        keep[k] = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
        setattr(cb, k, _ptr(keep[k]))
    cb.mla_deg = float(batch.mla)
    cb.mu_s = float(mu_s)

    main = ("I_dr", "I_df_d", "I_df_u", "F")
    extra = EXTRA_NAMES.get(scheme, ())
    if fields is not None:
        unknown = set(fields) - set(main) - set(extra) - {"rho_c"}
        if unknown:
            raise KeyError(f"{scheme}: no such output field(s): {', '.join(sorted(unknown))}")
    want = lambda k: fields is None or k in fields  # noqa: E731
    out = {k: alloc((S, nz, nw)) for k in main if want(k)}
    rows = nz - 1 if scheme == "n79" else nz
    for k in extra:
        if want(k):
            out[k] = alloc((S, rows, nw))
    if scheme == "bf" and want("rho_c"):
        out["rho_c"] = np.empty((S, nw))
    co = _abi.Out()
    for k in main:
        if k in out:
            setattr(co, k, _ptr(out[k]))
    for slot, k in zip(("x0", "x1", "x2"), extra):
        if k in out:
            setattr(co, slot, _ptr(out[k]))
    if "rho_c" in out:
        co.rho_c = _ptr(out["rho_c"])
    if band_w is not None:
        keep["band_w"] = np.ascontiguousarray(np.atleast_2d(np.asarray(band_w, dtype=np.float64)))
        co.band_w = _ptr(keep["band_w"])
        co.n_bw = keep["band_w"].shape[0]
        out["absorbed"] = np.empty((S, co.n_bw))
        co.absorbed = _ptr(out["absorbed"])
    if status:
        out["status"] = np.zeros(S, dtype=np.int32)
        co.status = _ptr(out["status"])
    rc = lib.crt1d_solve_host(_abi.SCHEME_IDS[scheme], ctypes.byref(cb), ctypes.byref(co),
                              _device_index() if device is None else int(device))
    _lib.check(rc)  # raises on errors; warns (RuntimeWarning) when some scenario is non-finite
    return out


def run_scheme(scheme, *, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, K_b_fn, soil_r=None, G_fn=None,
               mla=57.0, mu_s=0.501, tau_d_method="quad"):
    """One reference-style solver call on the GPU; returns the reference's dict of (n_z, n_wl) arrays."""
    lai = np.asarray(lai, dtype=np.float64)
    arrs = {k: np.asarray(v, dtype=np.float64) for k, v in
            dict(I_dr0_all=I_dr0_all, I_df0_all=I_df0_all, leaf_t=leaf_t, leaf_r=leaf_r).items()}
    if soil_r is not None:
        arrs["soil_r"] = np.asarray(soil_r, dtype=np.float64)
    nb = arrs["I_dr0_all"].size
    for k, v in arrs.items():
        if v.ndim != 1 or v.size != nb:
            raise ValueError(f"{k} must be 1-D with {nb} bands, got shape {v.shape}")
    if scheme in ("bf", "g77"):
        assert lai[0] == lai.max()  # ref _solve_bf.py:40, _solve_g77.py:35
    if G_fn is None:  # schemes that only take K_b_fn: G = K_b cos(psi)
        G_fn = lambda psi_: K_b_fn(psi_) * np.cos(psi_)  # noqa: E731
    batch = ScenarioBatch(
        psi=[psi], lai_lib=lai, leaf_r_lib=arrs["leaf_r"], leaf_t_lib=arrs["leaf_t"],
        soil_r_lib=arrs.get("soil_r", np.zeros(nb)), I_dr0_lib=arrs["I_dr0_all"], I_df0_lib=arrs["I_df0_all"],
        lai_idx=[0], leaf_idx=[0], soil_idx=[0], sky_idx=[0], mla=float(mla),
    )
    if scheme == "n79" and batch.n_z < 3:
        raise IndexError("index 1 is out of bounds for axis 0 with size 1")  # as ref _solve_n79.py:85 on td[1]
    pro = host_prologue(batch, scheme, K_b_fn=K_b_fn, G_fn=G_fn, mu_s=mu_s, tau_d_method=tau_d_method)
    res = solve_batch_host(batch, scheme, pro, mu_s=mu_s)
    sol = {k: v[0] for k, v in res.items()}
    if scheme == "bf":
        sol["rho_c"] = sol["rho_c"][-1]  # the reference returns the last band's scalar (ref _solve_bf.py:153)
    return sol

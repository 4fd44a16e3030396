"""TEST INFRASTRUCTURE: run the kernels' per-column arithmetic compiled for the host (tests/_hostcheck).

Lets the CPU test tier check the math the CUDA kernels execute against the oracle without a GPU.  The
product never imports this; it is not a fallback."""
import ctypes as C
import os
import subprocess

import numpy as np

from crt1d_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "_hostcheck", "hostcheck.cpp")
LIB = os.path.join(HERE, "_hostcheck", "libhostcheck.so")
CSRC = os.path.join(HERE, "..", "crt1d_b200", "csrc")
_lib = None


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("crt_core.cuh", "crt_scheme.cuh", "crt_leafangle.cuh", "crt_spectra.cuh")]
    deps.append(os.path.join(HERE, "..", "include", "crt1d_b200.h"))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
                               "-ffp-contract=off", "-o", LIB, SRC])
    return LIB


def build_variant(name, defines):
    """A second copy of the checker compiled with `-D` overrides (e.g. the 4s branch windows), for tests that compare
    two evaluation forms of the same kernel source."""
    path = os.path.join(HERE, "_hostcheck", f"libhostcheck_{name}.so")
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("crt_core.cuh", "crt_scheme.cuh", "crt_leafangle.cuh", "crt_spectra.cuh")]
    if not os.path.exists(path) or any(os.path.getmtime(d) > os.path.getmtime(path) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off"]
                              + [f"-D{d}" for d in defines] + ["-o", path, SRC])
    L = C.CDLL(path)
    L.hostcheck_solve.restype = C.c_int
    L.hostcheck_solve.argtypes = [C.c_int, C.POINTER(_abi.Batch), C.POINTER(_abi.Out), C.c_int]
    return L


class use_lib:
    """Context manager: route `solve()` through another build of the checker."""

    def __init__(self, L):
        self.L = L

    def __enter__(self):
        global _lib
        lib()
        self.saved, _lib = _lib, self.L
        return self

    def __exit__(self, *exc):
        global _lib
        _lib = self.saved


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.hostcheck_solve.restype = C.c_int
        L.hostcheck_solve.argtypes = [C.c_int, C.POINTER(_abi.Batch), C.POINTER(_abi.Out), C.c_int]
        L.hostcheck_leaf_G.restype = C.c_double
        L.hostcheck_leaf_G.argtypes = [C.c_int, C.c_double, C.c_double]
        L.hostcheck_tau_d.restype = C.c_double
        L.hostcheck_tau_d.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_double]
        L.hostcheck_leaf_integral.restype = C.c_double
        L.hostcheck_leaf_integral.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        L.hostcheck_smear_tuv.restype = None
        L.hostcheck_smear_tuv.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def solve(batch, scheme, prologue, *, vec=None, band_w=None, mu_s=0.501):
    """Host evaluation of `scheme` on a ScenarioBatch with a host prologue dict; returns numpy arrays."""
    from crt1d_b200.engine import EXTRA_NAMES

    S, nz, nw = batch.n_scen, batch.n_z, batch.n_wl
    keep = {}
    cb = _abi.Batch()
    cb.n_scen, cb.n_z, cb.n_wl = S, nz, nw
    cb.n_lai, cb.n_leaf = batch.lai_lib.shape[0], batch.leaf_r_lib.shape[0]
    cb.n_soil, cb.n_sky = batch.soil_r_lib.shape[0], batch.I_dr0_lib.shape[0]
    for k in ("psi", "lai_lib", "leaf_r_lib", "leaf_t_lib", "soil_r_lib", "I_dr0_lib", "I_df0_lib",
              "lai_idx", "leaf_idx", "soil_idx", "sky_idx"):
        keep[k] = np.ascontiguousarray(getattr(batch, k))
        setattr(cb, k, _ptr(keep[k]))
    for k, v in prologue.items():
        keep[k] = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
        setattr(cb, k, _ptr(keep[k]))
    cb.mla_deg = float(batch.mla)
    cb.mu_s = float(mu_s)
    out = {k: np.full((S, nz, nw), np.nan) for k in ("I_dr", "I_df_d", "I_df_u", "F")}
    rows = nz - 1 if scheme == "n79" else nz
    for k in EXTRA_NAMES.get(scheme, ()):
        out[k] = np.full((S, rows, nw), np.nan)
    if scheme == "bf":
        out["rho_c"] = np.full((S, nw), np.nan)
    co = _abi.Out()
    for k in ("I_dr", "I_df_d", "I_df_u", "F"):
        setattr(co, k, _ptr(out[k]))
    for slot, k in zip(("x0", "x1", "x2"), EXTRA_NAMES.get(scheme, ())):
        setattr(co, slot, _ptr(out[k]))
    if scheme == "bf":
        co.rho_c = _ptr(out["rho_c"])
    if band_w is not None:
        keep["bw"] = np.ascontiguousarray(np.atleast_2d(np.asarray(band_w, dtype=np.float64)))
        co.band_w = _ptr(keep["bw"])
        co.n_bw = keep["bw"].shape[0]
        out["absorbed"] = np.full((S, co.n_bw), np.nan)
        co.absorbed = _ptr(out["absorbed"])
    if vec is None:
        vec = 2 if nw % 2 == 0 else 1
    rc = lib().hostcheck_solve(_abi.SCHEME_IDS[scheme], C.byref(cb), C.byref(co), vec)
    assert rc == 0
    return out


def smear_tuv(x, y, bins):
    """The device function `smear_tuv_bin` compiled for the host, over all (row, bin) pairs."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y2 = np.ascontiguousarray(np.atleast_2d(np.asarray(y, dtype=np.float64)))
    bins = np.ascontiguousarray(bins, dtype=np.float64)
    out = np.empty((y2.shape[0], bins.size - 1))
    lib().hostcheck_smear_tuv(y2.shape[0], x.size, _ptr(x), _ptr(y2), bins.size - 1, _ptr(bins), _ptr(out))
    return out

"""Device-side execution of a `ScenarioBatch` through the C ABI.

torch is plumbing here: it owns HBM allocations (tensors), streams and (in `distributed.py`) NCCL.  All
arithmetic of the hot path happens in the kernels of libcrt1d_b200.so, which receive raw device
pointers (`tensor.data_ptr()`) and the current CUDA stream handle.
"""
import ctypes

import numpy as np

from . import _abi
from . import _lib
from .scenarios import ScenarioBatch
from .solvers import common as _common

SCHEMES = tuple(_abi.SCHEME_IDS)
EXTRA_NAMES = {
    "zq": ("I_df_d_ss", "I_df_u_ss", "F_ss"),
    "bf": ("aI_lsl", "aI_lsh", "aI_l"),
    "g77": ("aI_lsl", "aI_lsh", "aI_l"),
    "n79": ("aI_lsl", "aI_lsh"),
}
MAIN_NAMES = ("I_dr", "I_df_d", "I_df_u", "F")
DEFAULT_N_QUAD = 32  # Gauss-Legendre nodes per panel (12 graded panels) for the device-side prologue integrals


BATCH_ARRAYS = ("psi", "lai_lib", "leaf_r_lib", "leaf_t_lib", "soil_r_lib", "I_dr0_lib", "I_df0_lib",
                "lai_idx", "leaf_idx", "soil_idx", "sky_idx")


_INDEX_OF = {"lai_idx": "lai_lib", "leaf_idx": "leaf_r_lib", "soil_idx": "soil_r_lib", "sky_idx": "I_dr0_lib"}


def _abi_array(batch, k):
    """Array `k` of a batch in the dtype / layout the C ABI reads (float64, int32 row indices; C order), with the
    index ranges re-checked: attributes assigned after construction bypass `ScenarioBatch.__post_init__`, and an
    int64 or out-of-range index array would be read as garbage rows by the kernels."""
    a = getattr(batch, k)
    if k in _INDEX_OF:
        v = np.ascontiguousarray(a, dtype=np.int32)
        if v.shape != (batch.n_scen,):
            raise ValueError(f"{k} must have one entry per scenario ({batch.n_scen}), got shape {v.shape}")
        rows = getattr(batch, _INDEX_OF[k]).shape[0]
        if v.size and (v.min() < 0 or v.max() >= rows):
            raise IndexError(f"{k} out of range for a library of {rows} rows")
        return v
    return np.ascontiguousarray(a, dtype=np.float64)


def pin_batch(batch):
    """Page-locked host copies of a batch's arrays (made once; H2D from them is asynchronous DMA)."""
    torch = _torch()
    return {k: torch.as_tensor(_abi_array(batch, k)).pin_memory() for k in BATCH_ARRAYS}


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("crt1d_b200 needs a CUDA device (no CPU fallback): torch.cuda.is_available() is False")
    return torch


def _stream_ptr(torch):
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def host_prologue(batch: ScenarioBatch, scheme, *, K_b_fn=None, G_fn=None, mu_s=0.501, tau_d_method="quad"):
    """Band-independent scalars of every scenario evaluated ON THE HOST from Python callables, with the
    same scipy calls as the reference solver's prologue (plugin path).  Returns a dict of numpy arrays."""
    K_b_fn = K_b_fn or batch.leaf_angle.K_b_fn
    G_fn = G_fn or batch.leaf_angle.G_fn
    S = batch.n_scen
    pro = {"K_b": np.array([K_b_fn(p) for p in batch.psi], dtype=np.float64)}
    if scheme == "2s":
        pro["mu_bar"] = np.full(S, _common.mu_bar_fn(G_fn))  # depends on G_fn only (ref _solve_2s.py:32)
    elif scheme == "4s":
        pro["G"] = np.array([G_fn(p) for p in batch.psi], dtype=np.float64)
        pro["G_int"] = np.tile(np.array(_common.G_sector_integrals(G_fn, mu_s)), (S, 1))
    elif scheme == "zq":
        dm = np.array([_common.mean_dlai(row) for row in batch.lai_lib])
        ti = np.array([_common.tau_df_fn(K_b_fn, d) for d in dm])
        pro["tau_i"] = ti[batch.lai_idx]
        pro["tau_psi"] = np.array(
            [_common.tau_b_fn(K_b_fn, p, dm[i]) for p, i in zip(batch.psi, batch.lai_idx)], dtype=np.float64
        )
    elif scheme == "zq_pa":
        M = min(100, batch.n_z)  # ref _solve_zq_pa.py:95, :175: one tau_d for every equal-LAI layer
        td = np.array([_common.tau_df_fn(K_b_fn, row[0] / M) for row in batch.lai_lib])
        pro["tau_i"] = td[batch.lai_idx]
    elif scheme == "bl":
        pro["tau_d_lev"] = np.array([_common.tau_df_fn(K_b_fn, row) for row in batch.lai_lib])  # ref _solve_bl.py:35-37
    elif scheme == "n79":
        td = np.zeros_like(batch.lai_lib)
        for i, row in enumerate(batch.lai_lib):
            td[i, :-1] = _common.tau_df_fn(K_b_fn, row[:-1] - row[1:], method=tau_d_method)
        pro["tau_d_lev"] = td
    return pro


class DeviceBatch:
    """A `ScenarioBatch` resident in HBM plus the scheme prologue; owns the ctypes `crt1d_batch`."""

    def __init__(self, batch: ScenarioBatch, scheme, *, device=None, prologue="device", mu_s=0.501,
                 tau_d_method="quad", n_quad=DEFAULT_N_QUAD, pinned=None):
        if scheme not in _abi.SCHEME_IDS:
            raise KeyError(f"{scheme!r} is not a CUDA scheme; valid: {', '.join(SCHEMES)}")
        if scheme == "n79" and batch.n_z < 3:
            raise IndexError("n79 needs n_z >= 3 (the reference indexes td[1], _solve_n79.py:85)")
        torch = _torch()
        self.lib = _lib.load()
        self.scheme = scheme
        self.batch = batch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.mu_s = float(mu_s)
        self._t = {}
        with torch.cuda.device(self.device):
            put = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(self.device, non_blocking=False)  # noqa: E731
            for k in BATCH_ARRAYS:
                if pinned is not None:  # page-locked staging copies made once by the caller (pin_batch)
                    self._t[k] = pinned[k].to(self.device, non_blocking=True)
                else:
                    self._t[k] = put(_abi_array(batch, k))
            if isinstance(prologue, dict):
                for k, v in prologue.items():
                    self._t[k] = put(np.asarray(v, dtype=np.float64))
            elif prologue == "device":
                self._device_prologue(torch, tau_d_method, n_quad)
            else:
                raise ValueError("prologue must be 'device' or a dict from host_prologue()")
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self._t.values())
        self.cbatch = self._make_cbatch()

    # -- prologue on the device, parametric leaf-angle family -------------------------------------
    def _buf(self, torch, name, shape, like=None):
        """Reuse the device tensor `name` if it exists with the right shape (stable pointers across
        `reload()`), else allocate it."""
        t = self._t.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=torch.float64, device=self.device)
            self._t[name] = t
        return t

    def _device_prologue(self, torch, tau_d_method, n_quad):
        b, lib, la = self.batch, self.lib, self.batch.leaf_angle
        S = b.n_scen
        st = _stream_ptr(torch)
        if tau_d_method not in ("quad", "9sky"):
            raise ValueError("invalid `method`. Valid options are 'quad' and '9sky'.")
        K_b = self._buf(torch, "K_b", (S,))
        G = self._buf(torch, "G", (S,))
        _lib.check(lib.crt1d_leaf_G(la.family_id, la.param, S, self._t["psi"].data_ptr(), G.data_ptr(), K_b.data_ptr(), st))
        nq = 0 if tau_d_method == "9sky" else int(n_quad)
        if self.scheme in ("2s", "4s"):
            tri = self._buf(torch, "_tri", (3,))
            _lib.check(lib.crt1d_leaf_integrals(la.family_id, la.param, self.mu_s, int(n_quad), tri.data_ptr(), st))
            if self.scheme == "2s":
                self._buf(torch, "mu_bar", (S,)).copy_(tri[0].expand(S))
            else:
                self._buf(torch, "G_int", (S, 2)).copy_(tri[1:3].expand(S, 2))
        elif self.scheme == "bl":  # the reference's solve_bl has no `tau_d_method`: always the quadrature (_solve_bl.py:35-37)
            td = self._buf(torch, "tau_d_lev", tuple(self._t["lai_lib"].shape))
            _lib.check(lib.crt1d_tau_d(la.family_id, la.param, int(n_quad), td.numel(), self._t["lai_lib"].data_ptr(), td.data_ptr(), st))
        elif self.scheme == "n79":
            # layer thicknesses from the lai_lib that is ON THE DEVICE (after reload() the host mirror may lag)
            lai_d = self._t["lai_lib"]
            dl_d = torch.zeros_like(lai_d)
            dl_d[:, :-1] = lai_d[:, :-1] - lai_d[:, 1:]
            td = self._buf(torch, "tau_d_lev", tuple(dl_d.shape))
            _lib.check(lib.crt1d_tau_d(la.family_id, la.param, nq, td.numel(), dl_d.data_ptr(), td.data_ptr(), st))
        elif self.scheme == "zq_pa":
            M = min(100, b.n_z)
            dl_d = (self._t["lai_lib"][:, 0] / M).contiguous()
            td = torch.empty_like(dl_d)
            _lib.check(lib.crt1d_tau_d(la.family_id, la.param, int(n_quad), td.numel(), dl_d.data_ptr(), td.data_ptr(), st))
            self._buf(torch, "tau_i", (S,)).copy_(td[self._t["lai_idx"].long()])
        elif self.scheme == "zq":
            # |mean of the non-zero level differences| per profile (ref _solve_zq.py:50), from the device copy
            d = self._t["lai_lib"][:, 1:] - self._t["lai_lib"][:, :-1]
            nz_ = d != 0
            dm_d = ((d * nz_).sum(dim=1) / nz_.sum(dim=1)).abs().contiguous()
            ti = torch.empty_like(dm_d)
            _lib.check(lib.crt1d_tau_d(la.family_id, la.param, int(n_quad), ti.numel(), dm_d.data_ptr(), ti.data_ptr(), st))
            idx = self._t["lai_idx"].long()
            self._buf(torch, "tau_i", (S,)).copy_(ti[idx])
            self._buf(torch, "tau_psi", (S,)).copy_(torch.exp(-K_b * dm_d[idx]))  # tau_b_fn at the mean dlai (prologue scalar)

    def reload(self, pinned, batch=None, tau_d_method="quad", n_quad=DEFAULT_N_QUAD):
        """Copy a new batch of the SAME shape from page-locked host tensors into the existing device
        tensors (asynchronous DMA on the current stream) and redo the device prologue.  Device pointers --
        and therefore every ctypes call struct built from them -- stay valid.

        `batch`: the `ScenarioBatch` the pinned copies were made from (`pin_batch(batch)`); it becomes
        `self.batch`, the host mirror that `narrow()` slices.  If omitted, the mirror is rebuilt from the pinned
        tensors.  The prologue itself reads only device tensors, so the kernels never see a mix of old and new
        tables."""
        import copy

        torch = _torch()
        for k in BATCH_ARRAYS:
            if tuple(pinned[k].shape) != tuple(self._t[k].shape):
                raise ValueError(f"reload(): {k} has shape {tuple(pinned[k].shape)}, the resident batch {tuple(self._t[k].shape)}")
        if batch is None:
            batch = copy.copy(self.batch)
            for k in BATCH_ARRAYS:
                setattr(batch, k, pinned[k].numpy())
        elif (batch.n_scen, batch.n_z, batch.n_wl) != (self.batch.n_scen, self.batch.n_z, self.batch.n_wl):
            raise ValueError("reload(): the new batch must have the shape of the resident one")
        self.batch = batch
        with torch.cuda.device(self.device):
            for k in BATCH_ARRAYS:
                self._t[k].copy_(pinned[k], non_blocking=True)
            self._device_prologue(torch, tau_d_method, n_quad)
        self.cbatch.mla_deg = float(batch.mla)
        return self

    def _make_cbatch(self):
        b = self.batch
        cb = _abi.Batch()
        cb.n_scen, cb.n_z, cb.n_wl = b.n_scen, b.n_z, b.n_wl
        cb.n_lai, cb.n_leaf = b.lai_lib.shape[0], b.leaf_r_lib.shape[0]
        cb.n_soil, cb.n_sky = b.soil_r_lib.shape[0], b.I_dr0_lib.shape[0]
        for name, _ in _abi.Batch._fields_:
            if name in self._t:
                setattr(cb, name, self._t[name].data_ptr())
        cb.mla_deg = float(b.mla)
        cb.mu_s = self.mu_s
        return cb

    def tensor(self, name):
        return self._t[name]

    def narrow(self, lo, hi):
        """A view of scenarios [lo, hi) sharing every device tensor (no copies): chunked sweeps."""
        import copy

        v = copy.copy(self)
        v.batch = self.batch.slice(lo, hi)
        v._t = dict(self._t)
        for k in ("psi", "K_b", "G", "mu_bar", "G_int", "tau_i", "tau_psi", "lai_idx", "leaf_idx", "soil_idx", "sky_idx"):
            if k in v._t:
                v._t[k] = self._t[k][lo:hi]
        v.cbatch = v._make_cbatch()
        return v


class OutputBuffers:
    """Preallocated HBM outputs for up to `capacity` scenarios; reusable across chunks of a sweep."""

    def __init__(self, scheme, capacity, n_z, n_wl, *, device, fields=MAIN_NAMES, extras=True, band_w=None,
                 profile_dtype=None):
        torch = _torch()
        self.scheme, self.capacity, self.n_z, self.n_wl = scheme, int(capacity), int(n_z), int(n_wl)
        f64 = dict(dtype=torch.float64, device=device)
        pdt = torch.float64 if profile_dtype in (None, "f64", torch.float64) else profile_dtype
        if pdt in ("f32", torch.float32):
            pdt = torch.float32
            if scheme in ("n79", "zq"):
                raise ValueError(f"{scheme}: float32 profile storage is not available (n79 and zq park float64 "
                                 "elimination checkpoints in the profile arrays)")
        elif pdt is not torch.float64:
            raise ValueError("profile_dtype must be float64 (default) or float32")
        self.profile_f32 = pdt is torch.float32
        prof = dict(dtype=pdt, device=device)
        self.t = {}
        for k in fields:
            self.t[k] = torch.empty((self.capacity, n_z, n_wl), **prof)
        self.extra_names = EXTRA_NAMES.get(scheme, ()) if extras else ()
        rows = n_z - 1 if scheme == "n79" else n_z
        for k in self.extra_names:
            self.t[k] = torch.empty((self.capacity, rows, n_wl), **prof)
        if scheme == "bf" and extras:
            self.t["rho_c"] = torch.empty((self.capacity, n_wl), **f64)
        self.band_w = None
        if band_w is not None:
            bw = np.ascontiguousarray(np.atleast_2d(np.asarray(band_w, dtype=np.float64)))
            if bw.shape[1] != n_wl or not 1 <= bw.shape[0] <= 4:
                raise ValueError("band_w must be (1..4, n_wl)")
            self.band_w = torch.as_tensor(bw).to(device)
            self.t["absorbed"] = torch.empty((self.capacity, bw.shape[0]), **f64)
        # per-scenario status words (CRT1D_STATUS_NONFINITE): zeroed by the library at every launch
        self.t["status"] = torch.zeros((self.capacity,), dtype=torch.int32, device=device)

    def cout(self, n_scen):
        if n_scen > self.capacity:
            raise ValueError(f"{n_scen} scenarios exceed the buffer capacity {self.capacity}")
        co = _abi.Out()
        for k in MAIN_NAMES:
            if k in self.t:
                setattr(co, k, self.t[k].data_ptr())
        for slot, k in zip(("x0", "x1", "x2"), self.extra_names):
            setattr(co, slot, self.t[k].data_ptr())
        if "rho_c" in self.t:
            co.rho_c = self.t["rho_c"].data_ptr()
        if self.band_w is not None:
            co.band_w = self.band_w.data_ptr()
            co.n_bw = self.band_w.shape[0]
            co.absorbed = self.t["absorbed"].data_ptr()
        co.profile_f32 = 1 if self.profile_f32 else 0
        co.status = self.t["status"].data_ptr()
        return co

    def bytes_written_per_scenario(self):
        return sum(t[0].numel() * t.element_size() for k, t in self.t.items())


def solve_into(dbatch: DeviceBatch, out: OutputBuffers):
    """Enqueue one solve of every scenario of `dbatch` on the current stream (asynchronous)."""
    torch = _torch()
    with torch.cuda.device(dbatch.device):
        co = out.cout(dbatch.batch.n_scen)
        rc = dbatch.lib.crt1d_solve(_abi.SCHEME_IDS[dbatch.scheme], ctypes.byref(dbatch.cbatch), ctypes.byref(co),
                                    _stream_ptr(torch))
    _lib.check(rc)
    return out


def solve(batch, scheme, *, device=None, prologue="device", band_w=None, extras=True, mu_s=0.501,
          tau_d_method="quad", n_quad=DEFAULT_N_QUAD, profile_dtype=None):
    """Solve every scenario of `batch`; returns a dict of torch tensors in HBM, profiles (S, n_z, n_wl).
    `profile_dtype=torch.float32` stores the profiles as float32 (arithmetic stays float64; closed-form
    schemes only) and halves the HBM traffic."""
    db = batch if isinstance(batch, DeviceBatch) else DeviceBatch(
        batch, scheme, device=device, prologue=prologue, mu_s=mu_s, tau_d_method=tau_d_method, n_quad=n_quad)
    ob = OutputBuffers(scheme, db.batch.n_scen, db.batch.n_z, db.batch.n_wl, device=db.device, extras=extras, band_w=band_w,
                       profile_dtype=profile_dtype)
    solve_into(db, ob)
    return ob.t


def _profile_f64(torch, t, shape, device, name):
    """The kernels read raw pointers as contiguous float64 (S, n_z, n_wl) on `device`: check, or convert where
    that is exact (float32 storage -> float64, non-contiguous views -> packed copies)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch tensor in HBM, got {type(t).__name__}")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    if t.device != torch.device(device):
        raise ValueError(f"{name} lives on {t.device}, the batch on {device}")
    if t.dtype not in (torch.float64, torch.float32):
        raise TypeError(f"{name} must be float64 (or float32 storage), got {t.dtype}")
    return t.to(torch.float64).contiguous()


def calc_absorption(dbatch: DeviceBatch, I_dr, I_df_d, I_df_u):
    """Layerwise absorption of every scenario (replaces `_calc_absorption`, ref model.py:573-647).
    Inputs/outputs are torch tensors in HBM; outputs (S, n_z-1, n_wl)."""
    torch = _torch()
    S, nz, nw = dbatch.batch.n_scen, dbatch.batch.n_z, dbatch.batch.n_wl
    I_dr, I_df_d, I_df_u = (_profile_f64(torch, t, (S, nz, nw), dbatch.device, k)
                            for t, k in ((I_dr, "I_dr"), (I_df_d, "I_df_d"), (I_df_u, "I_df_u")))
    names = ("aI", "aI_df", "aI_dr", "aI_sh", "aI_sl", "aI_df_sl", "aI_df_sh")
    outs = {k: torch.empty((S, nz - 1, nw), dtype=torch.float64, device=dbatch.device) for k in names}
    ao = _abi.AbsorptionOut()
    for k in names:
        setattr(ao, k, outs[k].data_ptr())
    with torch.cuda.device(dbatch.device):
        rc = dbatch.lib.crt1d_calc_absorption(ctypes.byref(dbatch.cbatch), I_dr.data_ptr(), I_df_d.data_ptr(),
                                              I_df_u.data_ptr(), ctypes.byref(ao), _stream_ptr(torch))
    _lib.check(rc)
    return outs


EBAL_COLUMNS = ("incoming", "outgoing (reflected)", "soil absorbed", "canopy abs")


def energy_balance(I_dr, I_df_d, I_df_u, band_w):
    """Canopy energy balance per scenario and band group on the GPU (the reference's `compare_ebal`,
    ref diagnostics.py:476-530, batched): profiles `(S, n_z, n_wl)` torch tensors in HBM, `band_w`
    `(n_bw, n_wl)` weights (e.g. `spectra.band_weights`, or photon-flux weights).  Returns a tensor
    `(S, n_bw, 4)` with columns `EBAL_COLUMNS`; `incoming - outgoing - soil` closes against `canopy abs`."""
    torch = _torch()
    if not isinstance(I_dr, torch.Tensor) or I_dr.dim() != 3:
        raise ValueError("I_dr must be a (S, n_z, n_wl) torch tensor")
    S, nz, nw = I_dr.shape
    I_dr, I_df_d, I_df_u = (_profile_f64(torch, t, (S, nz, nw), I_dr.device, k)
                            for t, k in ((I_dr, "I_dr"), (I_df_d, "I_df_d"), (I_df_u, "I_df_u")))
    bw = np.ascontiguousarray(np.atleast_2d(np.asarray(band_w, dtype=np.float64)))
    if bw.shape[1] != nw or not 1 <= bw.shape[0] <= 4:
        raise ValueError("band_w must be (1..4, n_wl)")
    bw_d = torch.as_tensor(bw).to(I_dr.device)
    out = torch.empty((S, bw.shape[0], 4), dtype=torch.float64, device=I_dr.device)
    lib = _lib.load()
    with torch.cuda.device(I_dr.device):
        rc = lib.crt1d_energy_balance(S, nz, nw, I_dr.data_ptr(), I_df_d.data_ptr(), I_df_u.data_ptr(), bw_d.data_ptr(),
                                      bw.shape[0], out.data_ptr(), _stream_ptr(torch))
    _lib.check(rc)
    return out

"""The synthetic scenario sweep of BASELINE.json configs[2..3] ("cfg 3").

S = 10^6 = 100 solar zeniths x 100 total-LAI values x 100 spectra sets, scenario index
s = (i_sza * 100 + i_lai) * 100 + i_spec; 2100 one-nanometre bands (midpoints of the 400..2500 nm
PROSPECT sample grid), n_z = 60 interface levels.  "PROSPECT-style" spectra are seeded perturbations of
the bundled samples because prosail is not installed anywhere in this environment; the sky diffuse
fraction f_d is tied to the spectrum index.  Everything here is host-side numpy and deterministic.
"""
import numpy as np

from .cases import spectra_lib
from .leaf_angle import LeafAngle
from .scenarios import ScenarioBatch

N_SZA = 100
N_LAI = 100
N_SPEC = 100


def sweep_band_grid():
    """2100 band centres / widths (micrometres): midpoints of the 1-nm grid 400..2500 nm."""
    wl_nm = spectra_lib()["ps5_wl_nm"]
    wl = 0.5 * (wl_nm[:-1] + wl_nm[1:]) / 1000.0
    dwl = np.diff(wl_nm) / 1000.0
    return wl, dwl


def synthetic_spectra(n_spec=N_SPEC, seed=0):
    """(leaf_r, leaf_t, soil_r, I_dr0, I_df0), each (n_spec, 2100)."""
    lib = spectra_lib()
    rng = np.random.default_rng(seed)
    mid = lambda v: 0.5 * (v[:-1] + v[1:])  # noqa: E731  (midpoint averaging as ref data/__init__.py:99-102)
    r0, t0 = mid(lib["ps5_r"]), mid(lib["ps5_t"])
    dry, wet = mid(lib["soil_dry"]), mid(lib["soil_wet"])
    wl, dwl = sweep_band_grid()

    a_s = rng.uniform(0.7, 1.3, n_spec)[:, None]
    b_s = rng.uniform(0.7, 1.3, n_spec)[:, None]
    f_wet = rng.uniform(0.0, 1.0, n_spec)[:, None]
    f_d = rng.uniform(0.1, 0.9, n_spec)[:, None]

    leaf_r = np.clip(r0 * a_s, 1e-4, 0.6)
    leaf_t = np.clip(t0 * b_s, 1e-4, 0.6)
    over = np.maximum((leaf_r + leaf_t) / 0.98, 1.0)  # enforce r + t <= 0.98 (conservative-scattering guard)
    leaf_r = leaf_r / over
    leaf_t = leaf_t / over
    soil_r = f_wet * wet + (1 - f_wet) * dry

    si_tot = np.interp(wl, lib["sp2_wl_um"], lib["sp2_SI_dr"] + lib["sp2_SI_df"])
    I_tot = si_tot * dwl  # spectral -> in-band W m-2
    I_df0 = f_d * I_tot
    I_dr0 = (1 - f_d) * I_tot
    return leaf_r, leaf_t, soil_r, I_dr0, I_df0


def synthetic_sweep_spec(seed=0, n_z=60, n_sza=N_SZA, n_lai=N_LAI, n_spec=N_SPEC, sza_max_deg=85.0):
    """The full cross-product sweep as a `ScenarioBatch` (index arrays only: ~20 MB for 10^6 scenarios)."""
    leaf_r, leaf_t, soil_r, I_dr0, I_df0 = synthetic_spectra(n_spec, seed)
    wl, dwl = sweep_band_grid()
    sza = np.radians(np.linspace(0.0, sza_max_deg, n_sza))
    lai_tot = np.linspace(0.5, 8.0, n_lai)
    lai_lib = np.linspace(1.0, 0.0, n_z)[None, :] * lai_tot[:, None]  # what distribute_lai_beta yields (ref leaf_area.py:82-88)
    i_sza, i_lai, i_spec = np.meshgrid(
        np.arange(n_sza, dtype=np.int32), np.arange(n_lai, dtype=np.int32), np.arange(n_spec, dtype=np.int32),
        indexing="ij",
    )
    i_spec = i_spec.ravel()
    return ScenarioBatch(
        psi=sza[i_sza.ravel()], lai_lib=lai_lib, leaf_r_lib=leaf_r, leaf_t_lib=leaf_t, soil_r_lib=soil_r,
        I_dr0_lib=I_dr0, I_df0_lib=I_df0, lai_idx=i_lai.ravel(), leaf_idx=i_spec, soil_idx=i_spec,
        sky_idx=i_spec, leaf_angle=LeafAngle.from_mla(57), mla=57.0, wl=wl, dwl=dwl,
    )


def nonuniform_lai_library(lai_tot, n_z, h_c=20.0):
    """One cumulative-LAI profile per total in `lai_tot` from the generators whose axes are NOT equally spaced
    (`leaf_area.distribute_lai_weibull_z` for pine / spruce / birch crowns, `distribute_lai_gamma`), cycling through
    the four.  Rows obey the `ScenarioBatch` contract (`lai[0]` = total = max, `lai[-1] == 0`)."""
    from . import leaf_area

    z = np.linspace(0.0, h_c + 0.5, n_z)
    rows = []
    for i, tot in enumerate(np.asarray(lai_tot, dtype=np.float64)):
        kind = i % 4
        if kind == 3:
            lai = leaf_area.distribute_lai_gamma(h_c, float(tot), n_z).lai
        else:
            lai = leaf_area.distribute_lai_weibull_z(z, float(tot), h_c, hb=0.5, species=("pine", "spruce", "birch")[kind]).lai
        lai = np.array(lai, dtype=np.float64)
        lai[-1] = 0.0  # weibull_z ends in -0.0
        lai[0] = lai.max()
        rows.append(lai)
    return np.stack(rows)


def nonuniform_lai_spec(spec):
    """The same sweep as `spec` (indices, spectra, angles) with its LAI library replaced by non-uniform profiles of
    the same totals -- the level-recurrence fast paths of the kernels do not apply to any level group of these."""
    import copy

    out = copy.copy(spec)
    out.lai_lib = nonuniform_lai_library(spec.lai_lib[:, 0], spec.n_z)
    return out


class SweepRunner:
    """Chunked execution of a large `ScenarioBatch` on one GPU.

    The full profile output of the 10^6-scenario sweep is 4.03 TB, far beyond HBM,
    so scenarios are processed in chunks whose profile arrays cycle through a small ring of HBM
    buffers (each far larger than the 126 MB L2, so nothing is served from cache between chunks),
    while the fused per-scenario diagnostics (canopy-integrated absorbed PAR / NIR) of ALL scenarios
    stay resident and are what leaves the GPU.  One kernel launch per chunk; chunks are independent.

    `chunk` = scenarios per launch; a negative value means "at most |chunk|, rounded down to whole waves of resident
    CTAs for this scheme's kernel" (`crt1d_preferred_batch`).
    """

    def __init__(self, spec, scheme="2s", *, chunk=4096, device=None, n_buffers=2, bands=("PAR", "NIR"),
                 profiles=True, n_quad=32, profile_dtype=None, n_diag_buffers=1, order=None):
        import warnings

        from . import engine
        from .spectra import BAND_DEFNS_UM
        from .spectra import x_frac_in_bounds

        self.engine = engine
        # execution order: None = as given, or an explicit permutation (`ScenarioBatch.permuted`).  Profiles and
        # `absorbed` come out in EXECUTION order; `absorbed_in_spec_order()` undoes it.
        self.order = None
        if order is not None:
            self.order = np.asarray(order)
            spec = spec.permuted(self.order)
        self.spec, self.scheme = spec, scheme
        self.chunk = int(min(abs(chunk), spec.n_scen))
        if chunk < 0:  # "at most |chunk| scenarios per launch, whole waves of resident CTAs": the library knows its kernels
            from . import _abi
            from . import _lib

            idx = device if isinstance(device, int) else getattr(device, "index", None)
            idx = -1 if idx is None else int(idx)
            n = _lib.load().crt1d_preferred_batch(_abi.SCHEME_IDS[scheme], spec.n_z, spec.n_wl, self.chunk, idx)
            if n < 0:
                _lib.check(int(n))
            self.chunk = int(n)
        self.n_buffers = int(n_buffers)
        self.profiles = profiles
        self.profile_dtype = profile_dtype  # None / torch.float64, or torch.float32 storage (closed-form schemes)
        self.n_quad = n_quad
        self.device = device
        self.band_w = None
        if bands:
            wle = np.r_[spec.wl[0] - 0.5 * spec.dwl[0], spec.wl + 0.5 * spec.dwl]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.band_w = np.stack([x_frac_in_bounds(wle, BAND_DEFNS_UM[b]) for b in bands])
        self.db = None
        self.pinned = None  # set by pin_host(): page-locked staging copies of the scenario tables
        self.ring = None
        # per-scenario diagnostics of the whole sweep stay resident; with n_diag_buffers = 2 consecutive steps
        # alternate between two buffers, so a consumer on another stream (the NCCL all-gather of bench.py)
        # can read step k's diagnostics while step k+1 is being computed
        self.n_diag_buffers = int(n_diag_buffers)
        self.absorbed_bufs = None
        self._last = 0
        self._next = 0
        self._calls = None

    @property
    def absorbed(self):
        """Diagnostics `(S, n_bands)` of the most recent step."""
        return None if self.absorbed_bufs is None else self.absorbed_bufs[self._last]

    def absorbed_in_spec_order(self):
        """`absorbed` re-ordered to the scenario order of the batch this runner was built from."""
        if self.order is None:
            return self.absorbed
        import torch

        out = torch.empty_like(self.absorbed)
        out[torch.as_tensor(self.order, device=self.absorbed.device)] = self.absorbed
        return out

    @property
    def n_chunks(self):
        return (self.spec.n_scen + self.chunk - 1) // self.chunk

    @property
    def units_per_step(self):
        """layer.band solves per step = S * n_z * n_wl."""
        return self.spec.n_scen * self.spec.n_z * self.spec.n_wl

    def pin_host(self):
        """Stage the scenario tables in page-locked host memory once, so every `upload()` is pure DMA."""
        self.pinned = self.engine.pin_batch(self.spec)
        return self

    def upload(self):
        """H2D of the scenario tables and libraries + the device-side prologue; (re)builds call structs."""
        import ctypes

        import torch

        from . import _abi

        eng = self.engine
        if self.db is not None and self.pinned is not None and self._calls is not None:
            self.db.reload(self.pinned, self.spec, n_quad=self.n_quad)  # same shapes: DMA into place, call structs stay valid
            return self
        self.db = eng.DeviceBatch(self.spec, self.scheme, device=self.device, prologue="device", n_quad=self.n_quad,
                                  pinned=self.pinned)
        dev = self.db.device
        if self.ring is None:
            fields = eng.MAIN_NAMES if self.profiles else ()
            self.ring = [
                eng.OutputBuffers(self.scheme, self.chunk, self.spec.n_z, self.spec.n_wl, device=dev, fields=fields,
                                  extras=self.profiles, profile_dtype=self.profile_dtype)
                for _ in range(self.n_buffers if self.profiles else 1)
            ]
        if self.band_w is not None and getattr(self, "band_w_d", None) is None:
            self.band_w_d = torch.as_tensor(self.band_w).to(dev)
        if self.band_w is not None and self.absorbed_bufs is None:
            self.absorbed_bufs = [torch.empty((self.spec.n_scen, self.band_w.shape[0]), dtype=torch.float64, device=dev)
                                  for _ in range(max(1, self.n_diag_buffers))]
        self._calls = []  # [diagnostic buffer][chunk] -> (device view, crt1d_out)
        views = [self.db.narrow(c * self.chunk, min((c + 1) * self.chunk, self.spec.n_scen)) for c in range(self.n_chunks)]
        for b in range(len(self.absorbed_bufs) if self.absorbed_bufs else 1):
            calls = []
            for c, view in enumerate(views):
                lo, hi = c * self.chunk, min((c + 1) * self.chunk, self.spec.n_scen)
                co = self.ring[c % len(self.ring)].cout(hi - lo)
                if self.band_w is not None:
                    co.band_w = self.band_w_d.data_ptr()
                    co.n_bw = self.band_w.shape[0]
                    co.absorbed = self.absorbed_bufs[b][lo:hi].data_ptr()
                calls.append((view, co))
            self._calls.append(calls)
        self._sid = _abi.SCHEME_IDS[self.scheme]
        self._byref = ctypes.byref
        return self

    def step(self, events=None):
        """Enqueue one pass over all chunks on the current stream; returns the number of kernel launches.
        `events`: optional list that receives a (start, end, n_scenarios) CUDA-event record per launch."""
        import ctypes

        import torch

        from . import _lib

        lib = self.db.lib
        stream = torch.cuda.current_stream()
        sp = ctypes.c_void_p(stream.cuda_stream)
        calls = self._calls[self._next]
        self._last = self._next
        self._next = (self._next + 1) % len(self._calls)
        for view, co in calls:
            if events is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            rc = lib.crt1d_solve(self._sid, self._byref(view.cbatch), self._byref(co), sp)
            if rc != 0:
                _lib.check(rc)
            if events is not None:
                e1.record(stream)
                events.append((e0, e1, view.batch.n_scen))
        return len(calls)

    def algorithmic_bytes_per_unit(self):
        """bytes written per layer.band for this scheme + amortised input reads."""
        nz, nw = self.spec.n_z, self.spec.n_wl
        n_fields = 4 + len(self.engine.EXTRA_NAMES.get(self.scheme, ()))
        if self.scheme == "n79":
            n_fields = 4 + 2 * (nz - 1) / nz
        esz = 4.0 if self.ring and self.ring[0].profile_f32 else 8.0
        w = esz * n_fields if self.profiles else 0.0
        r = 5 * 8.0 / nz + 8.0 / nw
        return w + r

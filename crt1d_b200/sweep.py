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

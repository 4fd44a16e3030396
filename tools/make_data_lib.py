"""Pack the reference's bundled sample spectra into one small .npz the product can load anywhere.

Run HERE (build container, /root/reference present):  python tools/make_data_lib.py
Output: crt1d_b200/data/spectra_lib.npz  (committed; ~90 KB).  /root/reference does not exist on the
GPU box, so everything that needs these public sample spectra (default case, synthetic sweep) reads
the packed file instead.  Sources (read with numpy only, never copied as files):
  crt1d/data/PROSPECT_sample.txt                          (2101 x [wl_nm, r, t], 400-2500 nm)
  crt1d/data/PROSAIL_sample-soil.txt                      (2101 x [dry, wet])
  crt1d/data/SPCTRAL2_xls_default-spectrum.csv            (122 x [wl_um, direct, diffuse] W m-2 um-1)
  crt1d/data/ideal-green-leaf_SPCTRAL2-wavelengths.csv    (122 x [wl_um, t, r]; 14 trailing NaN rows)
"""
import os
import sys

import numpy as np

REF = os.environ.get("CRT1D_REFERENCE", "/root/reference")
D = os.path.join(REF, "crt1d", "data")
OUT = os.path.join(os.path.dirname(__file__), "..", "crt1d_b200", "data", "spectra_lib.npz")


def main():
    wl_nm, r, t = np.loadtxt(os.path.join(D, "PROSPECT_sample.txt"), unpack=True)
    dry, wet = np.loadtxt(os.path.join(D, "PROSAIL_sample-soil.txt"), unpack=True)
    sp2 = np.loadtxt(os.path.join(D, "SPCTRAL2_xls_default-spectrum.csv"), delimiter=",", skiprows=1)
    ideal = np.loadtxt(
        os.path.join(D, "ideal-green-leaf_SPCTRAL2-wavelengths.csv"), delimiter=",", skiprows=1
    )
    assert wl_nm.size == 2101 and dry.size == 2101 and sp2.shape == (122, 3) and ideal.shape == (122, 3)
    np.savez_compressed(
        OUT,
        ps5_wl_nm=wl_nm, ps5_r=r, ps5_t=t,
        soil_dry=dry, soil_wet=wet,
        sp2_wl_um=sp2[:, 0], sp2_SI_dr=sp2[:, 1], sp2_SI_df=sp2[:, 2],
        ideal_wl_um=ideal[:, 0], ideal_t=ideal[:, 1], ideal_r=ideal[:, 2],
    )
    print("wrote", os.path.abspath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())

"""Golden vectors for the spectral binning (`spectra.smear_tuv`, ref crt1d/spectra.py:221-300), generated HERE by
importing the unmodified reference (see _refimport.py).  Run:  python tests/golden/make_golden_smear.py

Cases: (a) the default case's leaf / soil spectra and spectral irradiances (107 bands, irregular grid) onto
regular 10-band and 37-band grids and onto bins that do not line up with, start below and end inside the source
grid; (b) a 1-nm sample of the synthetic sweep library onto 10-nm bins; (c) single-trapezoid and degenerate
placements (a bin inside one source interval, a bin edge exactly on a source point)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refimport import import_reference_spectra  # noqa: E402


def main():
    ref = import_reference_spectra()
    d = np.load(os.path.join(HERE, "default_inputs.npz"))
    wl, dwl = d["wl"], d["dwl"]
    rows = np.stack([d["leaf_r"], d["leaf_t"], d["soil_r"], d["I_dr0_all"] / dwl, d["I_df0_all"] / dwl])
    out = {"a_x": wl, "a_y": rows}
    bins_sets = {
        "a_bins10": np.linspace(0.35, 2.5, 11),
        "a_bins37": np.linspace(wl[0], wl[-1], 38),
        "a_bins_off": np.array([0.29, 0.3025, 0.31, 0.4, 0.40001, 0.7, 1.3333, 2.0, 2.549, 2.55]),
    }
    for name, bins in bins_sets.items():
        out[name] = bins
        out[name + "_res"] = np.stack([ref.smear_tuv(wl, r, bins) for r in rows])
    from crt1d_b200 import sweep

    spec = sweep.synthetic_sweep_spec(seed=0)
    x = spec.wl[:600]
    y = np.stack([spec.leaf_r_lib[3, :600], spec.soil_r_lib[17, :600], spec.I_dr0_lib[5, :600] / spec.dwl[:600]])
    bins = np.arange(0.4, 1.0 + 1e-9, 0.01)
    out.update(b_x=x, b_y=y, b_bins=bins, b_res=np.stack([ref.smear_tuv(x, r, bins) for r in y]))
    x = np.array([0.0, 1.0, 2.0, 4.0])
    y = np.array([[1.0, 3.0, 2.0, -1.0]])
    bins = np.array([0.25, 0.5, 1.0, 1.0 + 2**-30, 3.0, 4.0])
    out.update(c_x=x, c_y=y, c_bins=bins, c_res=np.stack([ref.smear_tuv(x, r, bins) for r in y]))
    np.savez_compressed(os.path.join(HERE, "ref_smear_tuv.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

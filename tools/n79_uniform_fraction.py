"""How many LAI rows of the synthetic sweep take n79's per-column layer path (|td[j] - td[0]| <= 16 eps td[0]) with the
DEVICE prologue's tau_d values.  Run on the GPU box: python tools/n79_uniform_fraction.py"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from crt1d_b200 import engine, sweep  # noqa: E402

spec = sweep.synthetic_sweep_spec(seed=0).slice(0, 1000)
db = engine.DeviceBatch(spec, "n79")
td = db._t["tau_d_lev"].cpu().numpy()[:, :-1]
dev = np.abs(td - td[:, :1]) / td[:, :1]
ok = dev.max(axis=1) <= 3.6e-15
print("rows:", td.shape[0], "uniform within 16 eps:", int(ok.sum()), "max deviation:", float(dev.max()), "eps units:", float(dev.max() / 2.22e-16))

"""Ad-hoc: deep-canopy (n_z = 1000) tile-kernel timing with and without the fused absorbed reduction
(with it: one CTA per scenario walks all band tiles; without: one CTA per (scenario, tile))."""
import sys

import torch

from crt1d_b200 import sweep

scheme = sys.argv[1] if len(sys.argv) > 1 else "zq"
spec = sweep.synthetic_sweep_spec(seed=0, n_z=1000).slice(0, 1184)
for bands in (("PAR", "NIR"), ()):
    r = sweep.SweepRunner(spec, scheme, chunk=296, bands=bands).pin_host().upload()
    for _ in range(2):
        r.step()
    torch.cuda.synchronize()
    ev = []
    r.step(events=ev)
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    units = 296 * 1000 * 2100
    print(scheme, "bands" if bands else "no-absorbed", "kernel ms %.3f" % (sum(ms) / len(ms)), "units/s %.3e" % (units / (sum(ms) / len(ms) * 1e-3)))
    del r
    torch.cuda.empty_cache()

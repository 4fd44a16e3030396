"""`zq_pa` on the GPU: drop-in for the reference's `solve_zq_pa` (ref crt1d/solvers/_solve_zq_pa.py:24-418)."""
from ._plugin import run_scheme

short_name = "ZQ-pA"
long_name = "Zhao & Qualls multi-scattering (pyAPES)"


def solve_zq_pa(*, psi, I_dr0_all, I_df0_all, lai, clump, leaf_t, leaf_r, soil_r, K_b_fn):
    """Zhao & Qualls (2005) as ported from pyAPES: the zq tridiagonal on min(100, n_z) equal-LAI layers,
    linearly interpolated back to the input levels.  `clump` only enters the reference's absorption
    block, which does not reach its return value; it is accepted and ignored here too."""
    del clump
    return run_scheme("zq_pa", psi=psi, I_dr0_all=I_dr0_all, I_df0_all=I_df0_all, lai=lai, leaf_t=leaf_t,
                      leaf_r=leaf_r, soil_r=soil_r, K_b_fn=K_b_fn)

"""`bf` on the GPU: drop-in for the reference's `solve_bf` (ref crt1d/solvers/_solve_bf.py:7-154)."""
from ._plugin import run_scheme

short_name = "BF"
long_name = "Bodin & Franklin improved Goudriaan"


def solve_bf(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn):
    """Bodin & Franklin (2012) ~1.5-stream scheme.  Returns the four standard profiles plus
    `aI_lsl, aI_lsh, aI_l` `(n_z, n_wl)` and the scalar `rho_c` of the last band, as the reference."""
    return run_scheme("bf", psi=psi, I_dr0_all=I_dr0_all, I_df0_all=I_df0_all, lai=lai, leaf_t=leaf_t,
                      leaf_r=leaf_r, soil_r=soil_r, K_b_fn=K_b_fn)

"""`g77` on the GPU: drop-in for the reference's `solve_g77` (ref crt1d/solvers/_solve_g77.py:7-135)."""
from ._plugin import run_scheme

short_name = "G77"
long_name = "Goudriaan (1977)"


def solve_g77(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn):
    """Goudriaan (1977) as formulated by Bodin & Franklin (2012); extras `aI_lsl, aI_lsh, aI_l`."""
    return run_scheme("g77", psi=psi, I_dr0_all=I_dr0_all, I_df0_all=I_df0_all, lai=lai, leaf_t=leaf_t,
                      leaf_r=leaf_r, soil_r=soil_r, K_b_fn=K_b_fn)

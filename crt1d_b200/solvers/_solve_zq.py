"""`zq` on the GPU: drop-in for the reference's `solve_zq` (ref crt1d/solvers/_solve_zq.py:13-229)."""
from ._plugin import run_scheme

short_name = "ZQ"
long_name = "Zhao & Qualls multi-scattering"


def solve_zq(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn, G_fn):
    """Zhao & Qualls (2005): (2 n_z + 2) tridiagonal system per band + multiple-scattering correction;
    extras `I_df_d_ss, I_df_u_ss, F_ss` (single-scattering results)."""
    return run_scheme("zq", psi=psi, I_dr0_all=I_dr0_all, I_df0_all=I_df0_all, lai=lai, leaf_t=leaf_t,
                      leaf_r=leaf_r, soil_r=soil_r, K_b_fn=K_b_fn, G_fn=G_fn)

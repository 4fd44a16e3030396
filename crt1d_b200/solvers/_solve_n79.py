"""`n79` on the GPU: drop-in for the reference's `solve_n79` (ref crt1d/solvers/_solve_n79.py:11-200)."""
from ._plugin import run_scheme

short_name = "N79"
long_name = "Norman (1979)"


def solve_n79(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn, tau_d_method="quad"):
    """Norman (1979) after Bonan SP 14.3: a 2 n_z tridiagonal system per band solved with the reference's
    Thomas recurrences; extras `aI_lsl, aI_lsh` are `(n_z-1, n_wl)`.  `tau_d_method` in {'quad', '9sky'}."""
    return run_scheme("n79", psi=psi, I_dr0_all=I_dr0_all, I_df0_all=I_df0_all, lai=lai, leaf_t=leaf_t,
                      leaf_r=leaf_r, soil_r=soil_r, K_b_fn=K_b_fn, tau_d_method=tau_d_method)

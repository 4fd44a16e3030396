"""`2s` on the GPU: drop-in for the reference's `solve_2s` (ref crt1d/solvers/_solve_2s.py:11-163)."""
from ._plugin import run_scheme

short_name = "2s"
long_name = "Dickinson–Sellers two-stream"


def solve_2s(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn, G_fn, mla):
    """Dickinson-Sellers two-stream (Sellers 1985, with the 1996 correction), sm_100a kernel.

    Same keyword arguments and return dict (`I_dr, I_df_d, I_df_u, F`, each `(n_z, n_wl)` float64) as
    the reference function; `K_b_fn`/`G_fn` are evaluated on the host exactly where the reference
    evaluates them (K_b, and the quad integral for mu_bar)."""
    return run_scheme("2s", psi=psi, I_dr0_all=I_dr0_all, I_df0_all=I_df0_all, lai=lai, leaf_t=leaf_t,
                      leaf_r=leaf_r, soil_r=soil_r, K_b_fn=K_b_fn, G_fn=G_fn, mla=mla)

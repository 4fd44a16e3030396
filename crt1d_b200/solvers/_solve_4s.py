"""`4s` on the GPU: drop-in for the reference's `solve_4s` (ref crt1d/solvers/_solve_4s.py:8-293)."""
from ._plugin import run_scheme

short_name = "4s"
long_name = "Tian et al. four-stream"


def solve_4s(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn, G_fn, mu_s=0.501):
    """Tian et al. (2007) four-stream, closed-form eigen-solution of the reference's linear BVP.

    The reference integrates the 4-ODE system with `scipy.integrate.solve_bvp(tol=1e-6)`; this kernel
    evaluates the exact solution of the same equations and boundary conditions, so it agrees with the
    reference run at tight tolerance to ~1e-10 and with the as-shipped `tol=1e-6` run to the latter's
    own accuracy (~1e-4 relative), see DESIGN.md.  `mu_s` as in the reference."""
    return run_scheme("4s", psi=psi, I_dr0_all=I_dr0_all, I_df0_all=I_df0_all, lai=lai, leaf_t=leaf_t,
                      leaf_r=leaf_r, soil_r=soil_r, K_b_fn=K_b_fn, G_fn=G_fn, mu_s=mu_s)

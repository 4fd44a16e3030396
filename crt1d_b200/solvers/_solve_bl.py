"""`bl` on the GPU: drop-in for the reference's `solve_bl` (ref crt1d/solvers/_solve_bl.py:9-93)."""
from ._plugin import run_scheme

short_name = "B–L"
long_name = "Beer–Lambert"


def solve_bl(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, K_b_fn):
    """Beer-Lambert attenuation with grey-leaf scattering of the direct beam; `I_df_u` is zero, as in
    the reference."""
    return run_scheme("bl", psi=psi, I_dr0_all=I_dr0_all, I_df0_all=I_df0_all, lai=lai, leaf_t=leaf_t,
                      leaf_r=leaf_r, K_b_fn=K_b_fn)

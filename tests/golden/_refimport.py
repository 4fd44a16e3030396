"""Import the *unmodified* reference (zmoon/crt1d at /root/reference) for golden-vector generation.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (`crt1d_b200/`), in `bench.py`, in `smoke()` or in
the `-m gpu` tests may import this module: /root/reference does not exist on the GPU box.  It is used
by `make_golden.py` (here, in the build container) and by the optional `-m "not gpu"` tests that
re-validate the oracle against the live reference when the reference tree happens to be present.

`import crt1d` fails in this image (xarray / matplotlib / crt1d._version are absent), so the solvers are
imported through a stub parent package that skips `crt1d/__init__.py` (recipe A).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CRT1D_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "crt1d", "solvers"))


def import_reference_solvers():
    """Recipe A: returns (solvers, leaf_angle, leaf_area) modules of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "crt1d" not in sys.modules or not getattr(sys.modules["crt1d"], "_b200_stub", False):
        pkg = types.ModuleType("crt1d")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "crt1d")]
        pkg._b200_stub = True
        sys.modules["crt1d"] = pkg
    import crt1d.leaf_angle as leaf_angle  # noqa: E402
    import crt1d.leaf_area as leaf_area  # noqa: E402
    import crt1d.solvers as solvers  # noqa: E402

    return solvers, leaf_angle, leaf_area


def import_reference_spectra():
    """`crt1d.spectra` needs xarray at import; give it an empty stub (only numpy paths are exercised)."""
    import_reference_solvers()
    if "xarray" not in sys.modules:
        sys.modules["xarray"] = types.ModuleType("xarray")
    import crt1d.spectra as spectra  # noqa: E402

    return spectra


def tight_4s_solver(tol=1e-11, max_nodes=200000):
    """The reference's solve_4s with scipy.integrate.solve_bvp forced to a tight tolerance.

    Same `eqns` / `dfdr_bcs` closures, same quad calls; only `tol`/`max_nodes` change.
    """
    import scipy.integrate as _integ

    solvers, _, _ = import_reference_solvers()
    import crt1d.solvers._solve_4s as m4

    class _Shim:
        quad = staticmethod(_integ.quad)

        @staticmethod
        def solve_bvp(fun, bc, x, y, tol=None, **kw):  # noqa: ARG004 - tol deliberately overridden
            return _integ.solve_bvp(fun, bc, x, y, tol=tol_, max_nodes=max_nodes, **kw)

    tol_ = tol

    def solve_4s_tight(**kwargs):
        saved = m4.integrate
        m4.integrate = _Shim
        try:
            return m4.solve_4s(**kwargs)
        finally:
            m4.integrate = saved

    return solve_4s_tight

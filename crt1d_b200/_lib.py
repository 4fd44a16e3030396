"""Loader of the CUDA shared library.  There is NO fallback: if libcrt1d_b200.so is missing, stale or
unloadable, or no GPU is visible when a solver is called, the call raises."""
import ctypes
import os

from . import _abi

# CRT1D_B200_LIB selects a tuning variant built by `build.build(defines=..., out=...)` (benchmarks only)
LIB_PATH = os.environ.get("CRT1D_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcrt1d_b200.so")
_lib = None


class Crt1dB200Error(RuntimeError):
    """A C-ABI call returned a negative error code."""

    def __init__(self, code, message):
        super().__init__(f"crt1d_b200 error {code}: {message}")
        self.code = code


def load():
    """Load (once) and return the ctypes handle of libcrt1d_b200.so."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m crt1d_b200.build` (needs nvcc). "
                "crt1d_b200 has no CPU fallback."
            )
        lib = _abi.declare(ctypes.CDLL(LIB_PATH))
        v = lib.crt1d_abi_version()
        if v != _abi.ABI_VERSION:
            raise ImportError(f"{LIB_PATH} has ABI version {v}, Python layer expects {_abi.ABI_VERSION}: rebuild")
        _lib = lib
    return _lib


def check(code):
    """Raise Crt1dB200Error for a negative return code.  The positive finding CRT1D_NONFINITE (results delivered,
    some scenario holds NaN/Inf) becomes a RuntimeWarning: the reference returns such arrays with at most a
    numpy RuntimeWarning, and a drop-in must not turn that into an exception."""
    if code == _abi.OK:
        return code
    lib = load()
    detail = lib.crt1d_last_error().decode() or lib.crt1d_strerror(code).decode()
    if code == _abi.NONFINITE:
        import warnings

        warnings.warn(f"crt1d_b200: {detail}", RuntimeWarning, stacklevel=3)
        return code
    raise Crt1dB200Error(code, detail)

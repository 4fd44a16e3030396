"""The C-ABI library loads and exports exactly what include/crt1d_b200.h declares; argument validation
answers with error codes (never a crash, never a CPU fallback).  No compute calls -- runs without a GPU."""
import ctypes
import os
import re

import pytest

from crt1d_b200 import _abi
from crt1d_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "crt1d_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"CRT1D_API\s+[\w\s\*]+?\b(crt1d_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_abi.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert lib.crt1d_abi_version() == _abi.ABI_VERSION == int(re.search(r"#define CRT1D_ABI_VERSION (\d+)", open(HEADER).read()).group(1))


def test_header_constants_match_python_mirror():
    text = open(HEADER).read()
    for name, val in re.findall(r"#define CRT1D_SCHEME_(\w+) (\d+)", text):
        assert _abi.SCHEME_IDS[name.lower()] == int(val)
    from crt1d_b200.leaf_angle import G_FAMILY_IDS

    for name, val in re.findall(r"#define CRT1D_G_(\w+) (\d+)", text):
        assert G_FAMILY_IDS[name.lower()] == int(val)
    for name, val in re.findall(r"#define CRT1D_ERR_(\w+) \((-\d+)\)", text):
        assert getattr(_abi, f"ERR_{name}") == int(val)


def test_struct_layout_matches_c(tmp_path):
    """sizeof/offsetof of the ctypes mirrors equal what the C compiler lays out for the header."""
    import subprocess

    fields_b = [f for f, _ in _abi.Batch._fields_]
    fields_o = [f for f, _ in _abi.Out._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
             'printf("%zu %zu %zu\\n", sizeof(crt1d_batch), sizeof(crt1d_out), sizeof(crt1d_absorption_out));']
    lines += [f'printf("%zu\\n", offsetof(crt1d_batch, {f}));' for f in fields_b]
    lines += [f'printf("%zu\\n", offsetof(crt1d_out, {f}));' for f in fields_o]
    lines += ["return 0;}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert [int(x) for x in out[:3]] == [ctypes.sizeof(_abi.Batch), ctypes.sizeof(_abi.Out), ctypes.sizeof(_abi.AbsorptionOut)]
    offs = [int(x) for x in out[3:]]
    assert offs[: len(fields_b)] == [getattr(_abi.Batch, f).offset for f in fields_b]
    assert offs[len(fields_b):] == [getattr(_abi.Out, f).offset for f in fields_o]


def test_validation_error_codes_without_gpu():
    lib = _lib.load()
    cb, co = _abi.Batch(), _abi.Out()
    assert lib.crt1d_solve(-1, ctypes.byref(cb), ctypes.byref(co), None) == _abi.ERR_INVALID_ARG
    assert lib.crt1d_solve(7, ctypes.byref(cb), ctypes.byref(co), None) == _abi.ERR_INVALID_ARG
    assert lib.crt1d_solve(0, None, ctypes.byref(co), None) == _abi.ERR_NULL_POINTER
    cb.n_scen, cb.n_z, cb.n_wl, cb.n_lai, cb.n_leaf, cb.n_soil, cb.n_sky = 4, 1, 8, 1, 1, 1, 1
    assert lib.crt1d_solve(0, ctypes.byref(cb), ctypes.byref(co), None) == _abi.ERR_INVALID_ARG  # n_z < 2
    cb.n_z = 2
    assert lib.crt1d_solve(_abi.SCHEME_IDS["n79"], ctypes.byref(cb), ctypes.byref(co), None) == _abi.ERR_INVALID_ARG
    cb.n_z = 10
    assert lib.crt1d_solve(0, ctypes.byref(cb), ctypes.byref(co), None) == _abi.ERR_NULL_POINTER
    assert b"psi" in lib.crt1d_last_error()
    cb.n_scen = 0  # empty batch is a no-op, not an error
    assert lib.crt1d_solve(0, ctypes.byref(cb), ctypes.byref(co), None) == _abi.OK
    co.n_bw = 9
    cb.n_scen = 1
    assert lib.crt1d_leaf_G(42, 0.0, 1, None, None, None, None) == _abi.ERR_INVALID_ARG
    assert lib.crt1d_tau_d(0, 0.0, 999, 1, None, None, None) in (_abi.ERR_INVALID_ARG, _abi.ERR_NULL_POINTER)
    assert lib.crt1d_leaf_integrals(0, 0.0, 1.5, 32, None, None) == _abi.ERR_NULL_POINTER
    assert lib.crt1d_strerror(_abi.ERR_CUDA) == b"CUDA runtime error"
    with pytest.raises(_lib.Crt1dB200Error) as ei:
        _lib.check(lib.crt1d_solve(-1, ctypes.byref(cb), ctypes.byref(co), None))
    assert ei.value.code == _abi.ERR_INVALID_ARG


def test_no_cpu_fallback_when_gpu_missing(default_p):
    """On a machine without a GPU the product must fail loudly, not silently compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import crt1d_b200 as crt

    m = crt.Model("2s", nlayers=10)
    with pytest.raises(_lib.Crt1dB200Error) as ei:
        m.run()
    assert ei.value.code == _abi.ERR_NO_DEVICE
    with pytest.raises(RuntimeError):
        from crt1d_b200 import engine

        engine.solve(m.scenario_batch(), "2s")


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py's CPU-baseline legs may touch oracle/ (checked textually)."""
    pkg = os.path.join(ROOT, "crt1d_b200")
    bad_py = re.compile(r"^\s*(from|import)\s+(crt_oracle|hostcheck|oracle|tests)\b", re.M)
    bad_c = re.compile(r"#include\s+\"[^\"]*(oracle|tests)/", re.M)
    n = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            text = None
            if f.endswith(".py"):
                text, pat = open(os.path.join(dirpath, f)).read(), bad_py
            elif f.endswith((".cu", ".cuh", ".h")):
                text, pat = open(os.path.join(dirpath, f)).read(), bad_c
            if text is not None:
                n += 1
                assert not pat.search(text), f"{f} reaches into test infrastructure"
                assert "sys.path" not in text or f == "build.py", f"{f} manipulates sys.path"
    assert n > 20


def test_graft_entry_build_runs():
    """The driver's build check: `__graft_entry__.build()` compiles (or finds up to date) the CUDA library and the
    host math checker, loads the library and checks its ABI version against the Python layer's."""
    import importlib
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    ge = importlib.import_module("__graft_entry__")
    ge.build()

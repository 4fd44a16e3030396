"""CUDA-backed canopy RT solvers behind the reference's plugin interface.

`AVAILABLE_SCHEMES` has the same structure the reference builds at import
(ref crt1d/solvers/__init__.py:43-109): scheme id -> dict(name, module_name, short_name, long_name,
solver, args, options), discovered from `_solve_<id>.py` modules and the keyword-only signature of
`solve_<id>`; every `args` name must be a canopy-radiation-state input key.  `Model.run` consumes it as
`scheme["solver"](**{k: p[k] for k in scheme["args"]}, **extra)` (ref crt1d/model.py:305-310).

`register_cuda_schemes(target)` adds these entries to another registry -- e.g. the reference's own
`crt1d.solvers.AVAILABLE_SCHEMES` -- which is the external-registration hook the reference lists as a
TODO (ref crt1d/solvers/__init__.py:50).
"""
import inspect
import warnings
from importlib import import_module
from pathlib import Path

from ..variables import VMD as _vmd

__all__ = ["AVAILABLE_SCHEMES", "RET_KEYS_ALL_SCHEMES", "CANOPY_RAD_STATE_INPUT_KEYS", "register_cuda_schemes"]

CANOPY_RAD_STATE_INPUT_KEYS = list(_vmd.intent("in"))
RET_KEYS_ALL_SCHEMES = ["I_dr", "I_df_d", "I_df_u", "F"]
assert all(k in _vmd.intent("out") for k in RET_KEYS_ALL_SCHEMES)


def _scheme_dict(module_name):
    """Build one registry entry from a `_solve_<id>.py` module; None if its arguments are not valid inputs."""
    name = module_name[len("_solve_"):]
    module = import_module(f".{module_name}", package=__name__)
    solver = getattr(module, f"solve_{name}")
    long_name = getattr(module, "long_name", "")
    if not long_name:
        warnings.warn(f"`long_name` not defined for solver module {module_name!r}")
    spec = inspect.getfullargspec(solver)
    defaults = spec.kwonlydefaults or {}
    args = [k for k in spec.kwonlyargs if k not in defaults]
    invalid = [k for k in args if k not in CANOPY_RAD_STATE_INPUT_KEYS]
    if invalid:
        warnings.warn(
            f"Some arguments for scheme {name!r} not compatible with the expected:\n"
            f"  {', '.join(CANOPY_RAD_STATE_INPUT_KEYS)}\n"
            f"As a result, {name!r} will not be loaded.\nInvalid keys:\n  {', '.join(invalid)}"
        )
        return None
    return dict(
        module_name=module_name, name=name, short_name=getattr(module, "short_name", name), long_name=long_name,
        solver=solver, args=args, options=list(defaults),
    )


def _discover():
    found = {}
    for path in sorted(Path(__file__).parent.glob("_solve_*.py")):
        entry = _scheme_dict(path.stem)
        if entry is not None:
            found[entry["name"]] = entry
    return found


AVAILABLE_SCHEMES = _discover()
"""Scheme id -> info dict, same keys as the reference's registry."""

for _entry in AVAILABLE_SCHEMES.values():
    globals()[_entry["solver"].__name__] = _entry["solver"]
    __all__.append(_entry["solver"].__name__)


def register_cuda_schemes(target, *, suffix="_cuda", overwrite=False):
    """Add the CUDA schemes to another `AVAILABLE_SCHEMES`-style dict (e.g. the reference's).

    With the default suffix the reference's own `2s` stays and `2s_cuda` appears next to it;
    `suffix="", overwrite=True` swaps the implementations in place.  Returns the ids added."""
    added = []
    for name, entry in AVAILABLE_SCHEMES.items():
        key = name + suffix
        if key in target and not overwrite:
            raise KeyError(f"scheme {key!r} already registered (pass overwrite=True to replace)")
        e = dict(entry)
        e["name"] = key
        target[key] = e
        added.append(key)
    return added

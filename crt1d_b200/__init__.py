"""crt1d_b200 -- B200-native (sm_100a CUDA) solvers for the canopy radiative-transfer hot path of crt1d.

Public surface mirrors the reference for this path only: `Model`, `solvers.AVAILABLE_SCHEMES`,
`run_sensitivity`; plus the batched engine (`ScenarioBatch`, `engine`, `sweep.SweepRunner`).  All
numerics run in hand-written CUDA kernels behind the C ABI in `include/crt1d_b200.h`; there is no CPU
fallback -- a missing or unloadable `libcrt1d_b200.so`, or no GPU, raises at first use.
"""
__version__ = "0.1.0"

from . import cases  # noqa: F401
from . import leaf_angle  # noqa: F401
from . import leaf_area  # noqa: F401
from . import solvers  # noqa: F401
from . import spectra  # noqa: F401
from .leaf_angle import LeafAngle  # noqa: F401
from .model import Model  # noqa: F401
from .model import run_sensitivity  # noqa: F401
from .model import sensitivity_dataset  # noqa: F401
from .model import sensitivity_to_xr  # noqa: F401
from .scenarios import ScenarioBatch  # noqa: F401

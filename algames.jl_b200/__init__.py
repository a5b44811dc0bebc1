"""algames_b200 — B200-native batched Newton/KKT + augmented-Lagrangian solve behind Algames.jl's API.

The directory is named `algames.jl_b200` (not importable by that name); `import algames_b200` works through the
shim `algames_b200.py` at the repo root.
"""
from . import _capi
from .problem import *          # noqa: F401,F403
from .problem import AlgamesError, init_traj
from . import workloads
from . import distributed
from . import mpc

__version__ = "0.1.0"

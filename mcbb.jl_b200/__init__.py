"""mcbb.jl_b200 — B200-native (sm_100a) back end for the MCBB.jl hot path.

The directory name follows the reference's name; Python cannot import a dotted directory name directly, so
load it with `__graft_entry__.load_package()` (registers it as module `mcbb_jl_b200`).
"""
from . import _abi as abi  # noqa: F401
from ._lib import EXPORTS, LIB_PATH, MCBBError, check, kernel_launch_count, last_error, lib  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import eval_ode_run, mean, std, sort_  # noqa: F401
from . import api, dist  # noqa: F401

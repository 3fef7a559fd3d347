"""pomcp.jl_b200 — B200-native hot path of POMCP.jl behind the reference's
POMCPSolver / solve / action / update API (host side of the C ABI in
include/pomcp_b200.h).  Import as `pomcp_jl_b200`."""
from . import _abi  # noqa: F401
from .problems import (BabyPOMDP, LightDark1D, RockSamplePOMDP, TigerPOMDP, obs_bits,  # noqa: F401
                       obs_from_bits, philox4x32_10)
from .api import *  # noqa: F401,F403
from .api import __all__ as _api_all

__all__ = ["BabyPOMDP", "LightDark1D", "RockSamplePOMDP", "TigerPOMDP", "obs_bits", "obs_from_bits",
           "philox4x32_10"] + list(_api_all)

"""space_filling_forest_star_b200 -- B200-native collision-and-neighbour engine for the Space-Filling Forest* planner.

Only the planner's data-parallel hot path lives here (pose / edge collision verdicts, exact k-NN / radius search),
as hand-written sm_100a CUDA behind the C ABI of ``include/sffg.h``; this package is the thin host-side mirror of the
reference's interfaces for that path.  There is no CPU implementation in the package.
"""
from ._lib import SffgError, lib_path, load  # noqa: F401
from .engine import (BUILD_AUTO, BUILD_DEVICE, BUILD_HOST, COLLISION_SAMPLE_SIZE, ROT_INTERPOLATE, ROT_REFERENCE, Environment, Index, device_count,  # noqa: F401
                     gen_poses_device, init, load_mesh)

__all__ = ["Environment", "Index", "load_mesh", "init", "device_count", "gen_poses_device", "SffgError", "load", "lib_path",
           "ROT_REFERENCE", "ROT_INTERPOLATE", "COLLISION_SAMPLE_SIZE", "BUILD_AUTO", "BUILD_HOST", "BUILD_DEVICE"]

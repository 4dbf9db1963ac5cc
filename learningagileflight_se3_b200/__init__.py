"""B200-native batched SE(3) MPC engine: the gate-traversal MPC solve + reward hot path of
yanrui89/LearningAgileFlight_SE3 as hand-written sm_100a kernels behind the reference's call signatures.

    from learningagileflight_se3_b200 import ops                 # batched device / host operators
    from learningagileflight_se3_b200.quad_policy import run_quad  # drop-in for the reference class
"""
from . import _lib, ops  # noqa: F401
from ._lib import Params, Se3MpcError, default_params, launch_count  # noqa: F401

__all__ = ["ops", "Params", "Se3MpcError", "default_params", "launch_count"]

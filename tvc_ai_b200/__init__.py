"""tvc_ai_b200 -- B200-native (sm_100a) batched implementation of TVC-AI's EnhancedRocketTVCEnv
step/reset hot path behind the reference's Gymnasium env API.

Layout: csrc/ (CUDA kernels + C ABI -> libtvc_b200.so), _abi.py (ctypes stub), engine.py (handle +
torch buffers), env.py (single-env facade with the reference's names), vector_env.py (batched
VectorEnv), curriculum.py (stage index logic), dist.py (episode-statistics all-reduce), replay.py + sac.py (on-device replay ring
fed by the fused rollout kernel and the graph-captured batched SAC update: BASELINE config 5).
Importing the package does not need a GPU; constructing an env does (no CPU fallback).
"""
from . import _abi  # noqa: F401
from ._abi import CONTRACT_R, CONTRACT_X  # noqa: F401

__all__ = ["EnhancedRocketTVCEnv", "MissionPhase", "RocketTVCVectorEnv", "RocketTVCHostPipelineEnv", "BatchedEngine", "CurriculumManager",
           "make_training_env", "make_evaluation_env", "make_debug_env", "CONTRACT_R", "CONTRACT_X", "DeviceReplay", "SACLearner", "SACConfig",
           "train_sac"]


def __getattr__(name):
    if name in ("EnhancedRocketTVCEnv", "MissionPhase", "make_training_env", "make_evaluation_env", "make_debug_env",
                "make_enhanced_tvc_env", "install_as_reference_env"):
        from . import env
        return getattr(env, name)
    if name == "RocketTVCVectorEnv":
        from .vector_env import RocketTVCVectorEnv
        return RocketTVCVectorEnv
    if name == "RocketTVCHostPipelineEnv":
        from .vector_env import RocketTVCHostPipelineEnv
        return RocketTVCHostPipelineEnv
    if name == "BatchedEngine":
        from .engine import BatchedEngine
        return BatchedEngine
    if name == "CurriculumManager":
        from .curriculum import CurriculumManager
        return CurriculumManager
    if name == "DeviceReplay":
        from .replay import DeviceReplay
        return DeviceReplay
    if name in ("SACLearner", "SACConfig", "train_sac"):
        from . import sac
        return getattr(sac, name)
    raise AttributeError(name)

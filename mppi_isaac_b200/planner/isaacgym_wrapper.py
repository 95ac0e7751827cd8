"""Name-compatible import path: the reference keeps these symbols in
``mppiisaac/planner/isaacgym_wrapper.py`` (``IsaacGymConfig`` :10-18, ``ActorWrapper`` :49-77,
``IsaacGymWrapper`` :83).  Here ``IsaacGymWrapper`` is the CUDA-backed ``RolloutSim``."""
from ..utils.config_store import ActorWrapper, IsaacGymConfig  # noqa: F401
from .rollout_sim import RolloutSim as IsaacGymWrapper  # noqa: F401

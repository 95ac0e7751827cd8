from mppi_isaac_b200.planner.isaacgym_wrapper import ActorWrapper, IsaacGymConfig, IsaacGymWrapper  # noqa: F401

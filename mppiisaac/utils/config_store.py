from mppi_isaac_b200.utils.config_store import (ExampleConfig, IsaacGymConfig, MPPIConfig, load_actor_cfgs,  # noqa: F401
                                                load_config, load_isaacgym_config)

from mppi_isaac_b200.planner.mppi_isaac import MPPIisaacPlanner  # noqa: F401
from mppi_isaac_b200.planner.mppi import MPPIPlanner  # noqa: F401

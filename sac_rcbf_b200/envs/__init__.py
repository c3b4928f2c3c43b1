from .unicycle_env import UnicycleEnv  # noqa: F401
from .simulated_cars_env import SimulatedCarsEnv  # noqa: F401

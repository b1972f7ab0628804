"""microhh_b200 -- B200-native (sm_100a) dynamical core for MicroHH's RK3 step, behind a C ABI."""
from .grid import GridData  # noqa: F401

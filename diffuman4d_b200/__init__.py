"""diffuman4d_b200 -- Blackwell-native (sm_100a) denoise-step hot path for Diffuman4D."""
from .config import UNetConfig, SchedulerConfig  # noqa: F401

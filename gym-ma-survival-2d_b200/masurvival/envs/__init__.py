from .masurvival_env import MaSurvival, MaSurvivalVec  # noqa: F401

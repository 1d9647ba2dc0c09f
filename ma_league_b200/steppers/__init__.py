from .batched_stepper import BatchedEpisodeStepper
from .synthetic_env import SyntheticVecEnv

__all__ = ["BatchedEpisodeStepper", "SyntheticVecEnv"]

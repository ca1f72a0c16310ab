from .pendulum import PendulumVecEnv

__all__ = ["PendulumVecEnv"]

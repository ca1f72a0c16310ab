from .pendulum import PendulumEnv, PendulumVecEnv
from .cartpole import CartPoleVecEnv

__all__ = ["PendulumEnv", "PendulumVecEnv", "CartPoleVecEnv"]

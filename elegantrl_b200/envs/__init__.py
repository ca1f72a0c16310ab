from .pendulum import PendulumVecEnv
from .cartpole import CartPoleVecEnv

__all__ = ["PendulumVecEnv", "CartPoleVecEnv"]

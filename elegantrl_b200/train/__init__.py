from .replay_buffer import ReplayBuffer
from .run import train_agent

__all__ = ["ReplayBuffer", "train_agent"]

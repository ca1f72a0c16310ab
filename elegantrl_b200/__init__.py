"""B200-native on-policy actor-learner behind ElegantRL's Agent / vec-env API (rollout -> GAE -> PPO update)."""
from .config import Config

__all__ = ["Config"]
__version__ = "0.1.0"

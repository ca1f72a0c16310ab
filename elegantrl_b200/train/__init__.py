from .run import train_agent

__all__ = ["train_agent"]

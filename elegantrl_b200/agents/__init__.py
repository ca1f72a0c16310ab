from .AgentPPO import AgentPPO
from .nets import ActorPPO, CriticPPO

__all__ = ["AgentPPO", "ActorPPO", "CriticPPO"]

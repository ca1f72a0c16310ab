from .AgentPPO import AgentPPO
from .nets import ActorPPO, CriticPPO
from . import helloworld

__all__ = ["AgentPPO", "ActorPPO", "CriticPPO", "helloworld"]

from .AgentPPO import AgentDiscretePPO, AgentPPO
from .nets import ActorDiscretePPO, ActorPPO, CriticPPO
from . import helloworld

__all__ = ["AgentPPO", "AgentDiscretePPO", "ActorPPO", "ActorDiscretePPO", "CriticPPO", "helloworld"]

from .AgentPPO import AgentA2C, AgentDiscreteA2C, AgentDiscretePPO, AgentPPO
from .AgentSAC import AgentSAC
from .nets import ActorDiscretePPO, ActorPPO, ActorSAC, CriticEnsemble, CriticPPO
from . import helloworld

__all__ = ["AgentPPO", "AgentDiscretePPO", "AgentA2C", "AgentDiscreteA2C", "AgentSAC", "ActorPPO", "ActorDiscretePPO", "CriticPPO",
           "ActorSAC", "CriticEnsemble", "helloworld"]

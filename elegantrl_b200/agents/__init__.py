from .AgentPPO import AgentA2C, AgentDiscreteA2C, AgentDiscretePPO, AgentPPO
from .nets import ActorDiscretePPO, ActorPPO, CriticPPO
from . import helloworld

__all__ = ["AgentPPO", "AgentDiscretePPO", "AgentA2C", "AgentDiscreteA2C", "ActorPPO", "ActorDiscretePPO", "CriticPPO",
           "helloworld"]

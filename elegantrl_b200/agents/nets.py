"""Actor / critic ``nn.Module`` containers for the on-policy path.

They hold the parameters the CUDA engine works on *in place* (the engine receives raw pointers into these
storages, see ``AgentPPO._net_desc``) and provide the plain-PyTorch forward the callers of the hot path need:
the Evaluator calls ``actor(state)`` and pickles the module (reference ``elegantrl/train/evaluator.py:123,
176-183``), ``run.py:288-293`` ships ``agent.act`` through a Pipe.  Attribute names / state_dict keys follow the
reference nets (``elegantrl/agents/AgentPPO.py:348-390`` ActorPPO, ``:425-441`` CriticPPO; MLP builder
``elegantrl/agents/AgentBase.py:345-365``) so checkpoints' ``state_dict`` s are interchangeable.

The torch methods here are NOT the product path: ``AgentPPO`` routes rollout / GAE / update through
``libb200rl.so`` and raises if it is missing.
"""
import math

import torch as th
from torch import nn

TEN = th.Tensor
ACTIVATIONS = {"gelu": nn.GELU, "relu": nn.ReLU}


def make_mlp(dims, activation: str = "gelu") -> nn.Sequential:
    """Linear -> act -> ... -> Linear (no activation on the output), as the reference ``build_mlp``."""
    act_class = ACTIVATIONS[activation]
    layers = []
    for d_in, d_out in zip(dims[:-1], dims[1:]):
        layers += [nn.Linear(d_in, d_out), act_class()]
    return nn.Sequential(*layers[:-1])


def _init_output_layer(layer: nn.Linear, std: float, bias_const: float = 1e-6):
    nn.init.orthogonal_(layer.weight, std)
    nn.init.constant_(layer.bias, bias_const)


class _StateNormMixin:
    state_avg: nn.Parameter
    state_std: nn.Parameter

    def state_norm(self, state: TEN) -> TEN:
        if self.state_avg is None:  # helloworld nets have no input normalisation
            return state
        return (state - self.state_avg) / (self.state_std + 1e-4)


class ActorPPO(nn.Module, _StateNormMixin):
    """Gaussian policy: mean = MLP(state_norm(s)), std = exp(action_std_log) (state independent)."""

    def __init__(self, net_dims, state_dim: int, action_dim: int, activation: str = "gelu", state_norm: bool = True):
        super().__init__()
        self.net = make_mlp([state_dim, *net_dims, action_dim], activation)
        self.action_std_log = nn.Parameter(th.zeros((1, action_dim)), requires_grad=True)
        if state_norm:  # the elegantrl nets (AgentPPO.py:348-361); helloworld nets have neither (its :172-178)
            _init_output_layer(self.net[-1], std=0.1)
            self.state_avg = nn.Parameter(th.zeros((state_dim,)), requires_grad=False)
            self.state_std = nn.Parameter(th.ones((state_dim,)), requires_grad=False)
        else:
            self.state_avg = self.state_std = None
        self.activation = activation

    def forward(self, state: TEN) -> TEN:  # deterministic action for evaluation
        return self.net(self.state_norm(state)).tanh()

    def get_action(self, state: TEN):  # torch statement of what csrc/policy_step does
        mean = self.net(self.state_norm(state))
        std = self.action_std_log.exp()
        action = mean + std * th.randn_like(mean)
        return action, self._logprob(mean, std, action)

    def get_logprob_entropy(self, state: TEN, action: TEN):
        mean = self.net(self.state_norm(state))
        std = self.action_std_log.exp()
        entropy = (0.5 + 0.5 * math.log(2 * math.pi) + std.log()).expand_as(mean).sum(1)
        return self._logprob(mean, std, action), entropy

    @staticmethod
    def _logprob(mean: TEN, std: TEN, action: TEN) -> TEN:
        var = std * std
        return (-((action - mean) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(1)

    @staticmethod
    def convert_action_for_env(action: TEN) -> TEN:
        return action.tanh()


class CriticPPO(nn.Module, _StateNormMixin):
    """State-value net: V = MLP(state_norm(s))."""

    def __init__(self, net_dims, state_dim: int, action_dim: int, activation: str = "gelu", state_norm: bool = True):
        super().__init__()
        self.net = make_mlp([state_dim, *net_dims, 1], activation)
        if state_norm:
            _init_output_layer(self.net[-1], std=0.5)
            self.state_avg = nn.Parameter(th.zeros((state_dim,)), requires_grad=False)
            self.state_std = nn.Parameter(th.ones((state_dim,)), requires_grad=False)
        else:
            self.state_avg = self.state_std = None
        self.activation = activation

    def forward(self, state: TEN) -> TEN:
        return self.net(self.state_norm(state))


class ActorDiscretePPO(ActorPPO):
    """Categorical policy over ``action_dim`` actions: probs = softmax(MLP(state_norm(s))) -- reference
    ``elegantrl/agents/AgentPPO.py:393-425``.  Like the reference it inherits ``action_std_log`` (unused: it never
    receives a gradient, so Adam never touches it) which keeps the state_dict keys interchangeable."""

    def forward(self, state: TEN) -> TEN:  # greedy action index for evaluation
        return self.net(self.state_norm(state)).argmax(dim=1)

    def get_action(self, state: TEN):  # torch statement of what b200rl_policy_step_discrete does
        logits = self.net(self.state_norm(state)).log_softmax(dim=-1)
        race = logits.exp() / th.empty_like(logits).exponential_(1)  # torch.multinomial's one-draw fast path
        action = race.argmax(dim=-1)
        return action, logits.gather(1, action[:, None])[:, 0]

    def get_logprob_entropy(self, state: TEN, action: TEN):
        logits = self.net(self.state_norm(state)).log_softmax(dim=-1)
        return logits.gather(1, action.long().reshape(-1, 1))[:, 0], -(logits.exp() * logits).sum(1)

    @staticmethod
    def convert_action_for_env(action: TEN) -> TEN:
        return action.long()


# ------------------------------------------------------------------------------------------ off-policy (SAC) containers
class ActorSAC(nn.Module):
    """Container with the reference's attribute names / state_dict keys (``elegantrl/agents/AgentSAC.py:167-198``):
    ``net_s`` = build_mlp([S, *net_dims], if_raw_out=False) -- an activation after EVERY Linear -- and ``net_a`` = one Linear
    producing [mean | log_std].  ``forward`` (what the Evaluator calls) is the tanh of the mean."""

    def __init__(self, net_dims, state_dim: int, action_dim: int):
        super().__init__()
        self.state_dim, self.action_dim = state_dim, action_dim
        layers = []
        for d_in, d_out in zip([state_dim, *net_dims][:-1], [state_dim, *net_dims][1:]):
            layers += [nn.Linear(d_in, d_out), nn.GELU()]
        self.net_s = nn.Sequential(*layers)
        self.net_a = nn.Sequential(nn.Linear(net_dims[-1], action_dim * 2))
        _init_output_layer(self.net_a[-1], std=0.1)

    def forward(self, state: TEN) -> TEN:
        return self.net_a(self.net_s(state))[:, :self.action_dim].tanh()


class CriticEnsemble(nn.Module):
    """``elegantrl/agents/AgentSAC.py:244-259``: ``encoder_sa`` (one raw Linear of (state, action)) feeding
    ``decoder_q00`` ... (``build_mlp([*net_dims, 1])`` each).  ``forward`` = mean over the ensemble, as ``CriticBase.forward``."""

    def __init__(self, net_dims, state_dim: int, action_dim: int, num_ensembles: int = 4):
        super().__init__()
        self.state_dim, self.action_dim, self.num_ensembles = state_dim, action_dim, num_ensembles
        self.encoder_sa = nn.Sequential(nn.Linear(state_dim + action_dim, net_dims[0]))
        for i in range(num_ensembles):
            decoder = make_mlp([*net_dims, 1])
            _init_output_layer(decoder[-1], std=0.5)
            setattr(self, f"decoder_q{i:02}", decoder)

    @property
    def decoder_qs(self):
        return [getattr(self, f"decoder_q{i:02}") for i in range(self.num_ensembles)]

    def get_q_values(self, state: TEN, action: TEN) -> TEN:
        enc = self.encoder_sa(th.cat((state, action), dim=1))
        return th.cat([dec(enc) for dec in self.decoder_qs], dim=-1)

    def forward(self, state: TEN, action: TEN) -> TEN:
        return self.get_q_values(state, action).mean(dim=-1, keepdim=True)

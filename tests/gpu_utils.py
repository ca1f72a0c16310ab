"""Helpers shared by the GPU parity tests (build agents / tensors from the committed goldens)."""
import numpy as np
import torch as th

from elegantrl_b200 import Config
from elegantrl_b200.agents import AgentPPO
from tests import golden_utils as gu


def load_module(module, net):
    """Copy an oracle net dict (numpy) into a torch ActorPPO / CriticPPO."""
    linears = [m for m in module.net if isinstance(m, th.nn.Linear)]
    with th.no_grad():
        for layer, w, b in zip(linears, net["W"], net["b"]):
            layer.weight.copy_(th.from_numpy(np.ascontiguousarray(w)))
            layer.bias.copy_(th.from_numpy(np.ascontiguousarray(b)))
        module.state_avg.copy_(th.from_numpy(net["state_avg"]))
        module.state_std.copy_(th.from_numpy(net["state_std"]))
        if "action_std_log" in net:
            module.action_std_log.copy_(th.from_numpy(net["action_std_log"]))


def module_to_net(module):
    from oracle import ppo_oracle as po
    return po.net_from_torch(module)


def agent_from_golden(g, gpu_id=0, **overrides):
    dims = [int(x) for x in g["dims"]]
    state_dim, action_dim, num_envs, horizon_len = dims[:4]
    net_dims = dims[4:]
    hp = gu.hyper_of(g)
    env_args = {'env_name': 'golden', 'num_envs': num_envs, 'max_step': 200, 'state_dim': state_dim,
                'action_dim': action_dim, 'if_discrete': False}
    args = Config(agent_class=AgentPPO, env_class=None, env_args=env_args)
    args.net_dims = net_dims
    for k in ("gamma", "ratio_clip", "lambda_entropy", "clip_grad_norm", "learning_rate", "reward_scale",
              "batch_size", "repeat_times", "lambda_gae_adv", "if_use_v_trace"):
        setattr(args, k, hp[k])
    for k, v in overrides.items():
        setattr(args, k, v)
    agent = AgentPPO(net_dims, state_dim, action_dim, gpu_id=gpu_id, args=args)
    load_module(agent.act, gu.net_of(g, "actor"))
    load_module(agent.cri, gu.net_of(g, "critic"))
    return agent


def cuda(x, device="cuda:0"):
    return th.from_numpy(np.ascontiguousarray(x)).to(device)


def assert_close(actual, expected, rtol=1e-4, atol=1e-5, msg=""):
    if th.is_tensor(actual):
        actual = actual.detach().cpu().numpy()
    np.testing.assert_allclose(actual, expected, rtol=rtol, atol=atol, err_msg=msg)


def discrete_agent_from_golden(g, gpu_id=0, **overrides):
    """AgentDiscretePPO with the golden's nets (its inherited, unused action_std_log is loaded too)."""
    from elegantrl_b200.agents import AgentDiscretePPO
    dims = [int(x) for x in g["dims"]]
    state_dim, action_dim, num_envs, horizon_len = dims[:4]
    net_dims = dims[4:]
    hp = gu.hyper_of(g)
    env_args = {'env_name': 'golden', 'num_envs': num_envs, 'max_step': 200, 'state_dim': state_dim,
                'action_dim': action_dim, 'if_discrete': True}
    args = Config(agent_class=AgentDiscretePPO, env_class=None, env_args=env_args)
    args.net_dims = net_dims
    for k in ("gamma", "ratio_clip", "lambda_entropy", "clip_grad_norm", "learning_rate", "reward_scale",
              "batch_size", "repeat_times", "lambda_gae_adv", "if_use_v_trace"):
        setattr(args, k, hp[k])
    for k, v in overrides.items():
        setattr(args, k, v)
    agent = AgentDiscretePPO(net_dims, state_dim, action_dim, gpu_id=gpu_id, args=args)
    load_module(agent.act, gu.net_of(g, "actor"))
    load_module(agent.cri, gu.net_of(g, "critic"))
    return agent


def discrete_module_to_net(module):
    net = module_to_net(module)
    net.pop("action_std_log", None)
    return net

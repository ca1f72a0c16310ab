"""GPU parity of the off-policy path: ``elegantrl_b200.train.ReplayBuffer`` + ``elegantrl_b200.agents.
AgentSAC`` through the C-ABI (``b200rl_replay_append`` / ``b200rl_sac_policy_step`` / ``b200rl_sac_update``) against the goldens
minted from the reference's ``ReplayBuffer`` / ``AgentSAC`` (``oracle/make_golden.py::main_sac``, ``main_sac_cycle``; the
numpy oracle is pinned to the same files by tests/test_sac_oracle_golden.py).  rtol 1e-4 (north_star); ring contents,
pointer arithmetic and masks bit-exact."""
import glob
import os

import numpy as np
import pytest
import torch as th

from tests import golden_utils as gu
from tests import gpu_utils as G

pytestmark = pytest.mark.gpu
RTOL = 1e-4
SAC_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(gu.GOLDEN_DIR, "sac_*.npz")))
SACCYCLE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(gu.GOLDEN_DIR, "saccycle_*.npz")))
FIELDS = ("states", "actions", "rewards", "undones", "unmasks")


def _load_seq(seq, g, prefix):
    linears = [m for m in seq if isinstance(m, th.nn.Linear)]
    assert len(linears) == int(g[f"{prefix}.n_layers"])
    with th.no_grad():
        for i, layer in enumerate(linears):
            layer.weight.copy_(th.from_numpy(g[f"{prefix}.W{i}"]))
            layer.bias.copy_(th.from_numpy(g[f"{prefix}.b{i}"]))


def agent_from_golden(g, num_envs, **overrides):
    from elegantrl_b200 import Config
    from elegantrl_b200.agents import AgentSAC
    dims = [int(x) for x in g["dims"]]
    s_dim, a_dim = dims[0], dims[1]
    net_dims = dims[4:] if len(dims) > 6 or "append0.states" in g else dims[4:]
    args = Config()
    args.num_envs = num_envs
    args.num_ensembles = int(g["hp.num_ensembles"])
    for k in ("gamma", "clip_grad_norm", "learning_rate", "soft_update_tau", "repeat_times", "reward_scale"):
        setattr(args, k, float(g[f"hp.{k}"]))
    args.batch_size = int(g["hp.batch_size"])
    for k, v in overrides.items():
        setattr(args, k, v)
    agent = AgentSAC(net_dims, s_dim, a_dim, gpu_id=0, args=args)
    assert abs(agent.target_entropy - float(g["hp.target_entropy"])) < 1e-6
    _load_seq(agent.act.net_s, g, "init.actor.net_s")
    _load_seq(agent.act.net_a, g, "init.actor.net_a")
    for name, cri in (("critic", agent.cri), ("critic_target", agent.cri_target)):
        _load_seq(cri.encoder_sa, g, f"init.{name}.encoder")
        for e, dec in enumerate(cri.decoder_qs):
            _load_seq(dec, g, f"init.{name}.decoder{e}")
    with th.no_grad():
        agent.alpha_log.copy_(th.from_numpy(g["init.alpha_log"]))
    return agent


def _seq_params(seq):
    return [p for m in seq if isinstance(m, th.nn.Linear) for p in (m.weight, m.bias)]


def check_params(agent, g, prefix, atol):
    def want_seq(name):
        n = int(g[f"{name}.n_layers"])
        return [g[f"{name}.{t}{i}"] for i in range(n) for t in ("W", "b")]
    pairs = [(_seq_params(agent.act.net_s), want_seq(f"{prefix}.actor.net_s")), (_seq_params(agent.act.net_a), want_seq(f"{prefix}.actor.net_a"))]
    for name, cri in (("critic", agent.cri), ("critic_target", agent.cri_target)):
        pairs.append((_seq_params(cri.encoder_sa), want_seq(f"{prefix}.{name}.encoder")))
        for e, dec in enumerate(cri.decoder_qs):
            pairs.append((_seq_params(dec), want_seq(f"{prefix}.{name}.decoder{e}")))
    for got, want in pairs:
        for a, b in zip(got, want):
            G.assert_close(a, b, RTOL, atol)
    G.assert_close(agent.alpha_log, g[f"{prefix}.alpha_log"], RTOL, atol)


def build_buffer(g):
    from elegantrl_b200.train import ReplayBuffer
    s_dim, a_dim, num_seqs, max_size = (int(x) for x in g["dims"][:4])
    buf = ReplayBuffer(max_size, s_dim, a_dim, gpu_id=0, num_seqs=num_seqs)
    for i in range(3):
        buf.update(tuple(G.cuda(g[f"append{i}.{k}"]) for k in FIELDS))
        assert buf.p == int(g[f"append{i}.p"]) and buf.cur_size == int(g[f"append{i}.cur_size"])   # pointer arithmetic: exact
    return buf


@pytest.mark.parametrize("case", SAC_CASES)
def test_replay_buffer_ring(case):
    """ReplayBuffer.update incl. the wrap-around branch: ring contents bit-exact (reference replay_buffer.py:78-118)."""
    g = gu.load(case)
    buf = build_buffer(g)
    assert buf.if_full
    for k in FIELDS:
        assert np.array_equal(getattr(buf, k).cpu().numpy(), g[f"buffer.{k}"]), k
    state, action, reward, undone, unmask, next_state = buf.sample(16)   # API parity of sample()
    assert state.shape == (16, buf.state_dim) and next_state.shape == state.shape and reward.shape == (16,)
    assert int(buf.ids0.max()) < buf.cur_size - 1 and int(buf.ids1.max()) < buf.num_seqs


@pytest.mark.parametrize("case", SAC_CASES)
def test_sac_update_objectives_against_reference(case):
    """AgentSAC.update_objectives x1 and x3 with the reference's sampled ids and both rsample draws replayed: the two logged
    scalars, every parameter of actor / critic ensemble / target ensemble, alpha_log."""
    g = gu.load(case)
    n_updates = len(g["update.ids"])
    for k in (1, n_updates):
        buf = build_buffer(g)
        agent = agent_from_golden(g, num_envs=int(g["dims"][2]))
        agent.repeat_times = (k + 0.5) * agent.batch_size / buf.cur_size      # update_times = int(cur_size * repeat / batch) = k
        agent._inject_ids = G.cuda(g["update.ids"][:k])
        agent._inject_eps_next = G.cuda(g["update.eps_next"][:k])
        agent._inject_eps_pg = G.cuda(g["update.eps_pg"][:k])
        result = agent.update_net(buf)
        assert agent.last_update_info["update_times"] == k
        G.assert_close(np.array(result), g["update.scalars"][:k].mean(axis=0), RTOL, 2e-6)
        check_params(agent, g, "after1" if k == 1 else "after", 3e-6)
        assert float(agent.act_optimizer.state[agent.act.net_a[0].weight]["step"]) == k


def test_sac_policy_step_statistics_and_determinism():
    g = gu.load(SAC_CASES[0])
    agent = agent_from_golden(g, num_envs=4)
    state = th.zeros((100_000, agent.state_dim), device="cuda:0")
    a1 = agent.explore_action(state)
    a2 = agent.explore_action(state)
    assert not th.equal(a1, a2) and bool((a1.abs() <= 1).all())
    with th.no_grad():
        out = agent.act.net_a(agent.act.net_s(state[:1]))[0]
    avg, std = out[:agent.action_dim], out[agent.action_dim:].clamp(-16, 2).exp()
    z = (th.atanh(a1.clamp(-0.999999, 0.999999)) - avg) / std
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
    eps = th.randn((64, agent.action_dim), device="cuda:0")
    got = agent.explore_action(state[:64], eps)
    G.assert_close(got, th.tanh(avg + std * eps).cpu().numpy(), RTOL, 2e-6)


@pytest.mark.parametrize("case", SACCYCLE_CASES)
def test_sac_full_cycle_against_reference(case):
    """explore_env (tanh'ed actions stored) on the Pendulum vec env -> ReplayBuffer.update -> update_net."""
    from elegantrl_b200.envs import PendulumVecEnv
    from elegantrl_b200.train import ReplayBuffer
    g = gu.load(case)
    h, max_step, n, max_size = int(g["horizon_len"]), int(g["max_step"]), int(g["dims"][2]), int(g["dims"][3])
    agent = agent_from_golden(g, num_envs=n)
    env = PendulumVecEnv(num_envs=n, gpu_id=0, max_step=max_step)
    env.theta, env.theta_dot, env.cur_step = G.cuda(g["env.theta0"]), G.cuda(g["env.theta_dot0"]), G.cuda(g["env.cur_step0"])
    env.inject_reset_noise(G.cuda(g["env.reset_noise"]))
    agent.last_state = env.get_state()
    agent._inject_eps = G.cuda(g["explore.eps"])
    items = agent.explore_env(env, h)
    for name, got in zip(FIELDS[:3], items[:3]):
        G.assert_close(got, g[f"explore.{name}"], RTOL, 1e-5, name)
    assert np.array_equal(items[3].cpu().numpy(), g["explore.undones"]) and np.array_equal(items[4].cpu().numpy(), g["explore.unmasks"])
    G.assert_close(agent.last_state, g["explore.last_state"], RTOL, 1e-5)
    buf = ReplayBuffer(max_size, 3, 1, gpu_id=0, num_seqs=n)
    buf.update(tuple(G.cuda(g[f"explore.{k}"]) for k in FIELDS))   # the golden's own rollout: the update is checked in isolation
    assert buf.cur_size == h and not buf.if_full
    assert len(g["update.ids"]) == int(buf.cur_size * agent.repeat_times / agent.batch_size)
    agent._inject_ids, agent._inject_eps_next, agent._inject_eps_pg = (G.cuda(g[f"update.{k}"]) for k in ("ids", "eps_next", "eps_pg"))
    result = agent.update_net(buf)
    G.assert_close(np.array(result), g["update_net.result"], RTOL, 2e-6)
    check_params(agent, g, "after", 3e-6)

"""GPU parity of the discrete-action PPO path: ``AgentDiscretePPO`` / ``ActorDiscretePPO`` of the
reference (``elegantrl/agents/AgentPPO.py:252-270, 393-425``) through the C-ABI, against goldens minted from the reference
(``oracle/make_golden.py::main_discrete``).  Integer actions, masks and indices bit-exact; fp32 results rtol 1e-4."""
import ctypes as C

import numpy as np
import pytest
import torch as th

from elegantrl_b200 import _lib
from oracle import ppo_oracle as po
from tests import golden_utils as gu
from tests import gpu_utils as G

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(params=["tc", "cluster", "multilaunch"])
def update_impl(request, monkeypatch):
    monkeypatch.setenv("B200RL_UPDATE", request.param)
    return request.param


@pytest.fixture(params=["tc", "ffma"])
def forward_impl(request, monkeypatch):
    """tcgen05 forward kernel (64 x 64 GELU nets, csrc/forward_tc.cu) / CUDA-core kernel for every shape (csrc/forward.cu)"""
    if request.param == "ffma":
        monkeypatch.setenv("B200RL_FORWARD", "ffma")
    else:
        monkeypatch.delenv("B200RL_FORWARD", raising=False)
    return request.param


@pytest.mark.parametrize("case", gu.DISCRETE_CASES)
def test_policy_step_discrete_injected_noise(case, forward_impl):
    """b200rl_policy_step_discrete with torch.multinomial's Exp(1) noise replayed: sampled indices bit-exact."""
    g = gu.load(case)
    agent = G.discrete_agent_from_golden(g)
    action, logprob, env_action = agent._policy_step(G.cuda(g["nets.state"]), G.cuda(g["sample.expo"]))
    assert action.dtype == th.int32 and env_action.dtype == th.int64
    assert np.array_equal(action.cpu().numpy(), g["sample.action"])
    assert th.equal(env_action, action.long())
    G.assert_close(logprob, g["sample.logprob"], RTOL, 1e-6)
    # the torch module statement of the same thing (what an Evaluator calls)
    assert np.array_equal(agent.act(G.cuda(g["nets.state"])).cpu().numpy(), g["nets.actor_forward"])
    lp, ent = agent.act.get_logprob_entropy(G.cuda(g["nets.state"]), G.cuda(g["nets.action"]))
    G.assert_close(lp, g["nets.logprob"], RTOL, 1e-6)
    G.assert_close(ent, g["nets.entropy"], RTOL, 1e-6)


def test_policy_step_discrete_philox_statistics(forward_impl):
    """Device RNG: the empirical action frequencies of 400 000 draws follow softmax(logits); log-probs match the drawn index."""
    g = gu.load(gu.DISCRETE_CASES[-1])
    agent = G.discrete_agent_from_golden(g)
    state = G.cuda(g["nets.state"][:4])
    rows = 100_000
    batch = state.repeat_interleave(rows, dim=0).contiguous()
    action, logprob, _ = agent._policy_step(batch)
    action2, _, _ = agent._policy_step(batch)
    assert not th.equal(action, action2)  # the step counter advances the Philox stream
    logits = po.actor_mean(gu.discrete_net_of(g, "actor"), g["nets.state"][:4])
    probs, logp = po.categorical_from_logits(logits)
    a = action.cpu().numpy().reshape(4, rows)
    for i in range(4):
        freq = np.bincount(a[i], minlength=probs.shape[1]) / rows
        sigma = np.sqrt(probs[i] * (1 - probs[i]) / rows)
        assert np.all(np.abs(freq - probs[i]) < 5 * sigma + 1e-4), (freq, probs[i])
        G.assert_close(logprob.cpu().numpy().reshape(4, rows)[i], logp[i][a[i]], RTOL, 1e-5)


def _run_ppo_update(agent, buffer, ids):
    lib = _lib.load()
    states, actions, unmasks, logprobs, advantages, reward_sums = buffer
    h, n = states.shape[:2]
    act_desc, cri_desc = agent._net_desc(agent.act), agent._net_desc(agent.cri)
    assert not act_desc.action_std_log
    act_adam, cri_adam = agent._adam_desc(agent.act_optimizer, agent.act), agent._adam_desc(agent.cri_optimizer, agent.cri)
    ws = agent._get_workspace(act_desc, cri_desc)
    tb = _lib.TrainBuffer(states=states.data_ptr(), actions=actions.data_ptr(), unmasks=unmasks.data_ptr(),
                          logprobs=logprobs.data_ptr(), advantages=advantages.data_ptr(), reward_sums=reward_sums.data_ptr(),
                          adv_stats=None, horizon_len=h, num_envs=n, discrete_actions=1)
    hp = _lib.PPOHyper(ratio_clip=agent.ratio_clip, lambda_entropy=agent.lambda_entropy, clip_grad_norm=agent.clip_grad_norm)
    out = th.empty(3, device="cuda:0")
    ids = ids.contiguous()
    _lib.check(lib.b200rl_ppo_update(C.byref(act_desc), C.byref(cri_desc), C.byref(act_adam), C.byref(cri_adam), C.byref(tb),
                                     C.byref(hp), ids.shape[1], ids.shape[0], ids.data_ptr(), 0, 0, out.data_ptr(),
                                     ws.data_ptr(), ws.numel(), None))
    agent._set_adam_step(agent.act_optimizer, agent.act, act_adam.step)
    agent._set_adam_step(agent.cri_optimizer, agent.cri, cri_adam.step)
    return out.cpu().numpy()


def _check_params(agent, g, prefix, atol=2e-6):
    for which, module in (("actor", agent.act), ("critic", agent.cri)):
        got = gu.flat_params(G.discrete_module_to_net(module))
        ref = gu.flat_params(gu.discrete_net_of(g, f"{prefix}.{which}"))
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            G.assert_close(a, b, RTOL, atol, f"{which} {prefix}")
    # the inherited action_std_log is not a trainable of the categorical actor: untouched
    assert np.array_equal(agent.act.action_std_log.detach().cpu().numpy(), g["actor.action_std_log"])


@pytest.mark.parametrize("case", gu.DISCRETE_CASES)
def test_discrete_update_objectives_against_reference(case, update_impl):
    g = gu.load(case)
    buffer = [G.cuda(g[k]) for k in ("buf.states", "buf.actions", "buf.unmasks", "buf.logprobs", "gae.adv_norm", "gae.reward_sums")]
    assert buffer[1].dtype == th.int32
    ids = G.cuda(g["update.ids"])
    agent = G.discrete_agent_from_golden(g)
    scalars = _run_ppo_update(agent, buffer, ids[:1])
    G.assert_close(scalars, g["update.scalars"][0], RTOL, 1e-6)
    _check_params(agent, g, "update.after1", 1e-6)
    agent = G.discrete_agent_from_golden(g)
    scalars = _run_ppo_update(agent, buffer, ids)
    G.assert_close(scalars, g["update.scalars"].mean(axis=0), RTOL, 1e-6)
    _check_params(agent, g, "update.after")


@pytest.mark.parametrize("case", gu.DISCRETE_CASES + gu.CARTPOLE_CASES)
def test_discrete_update_net_against_reference(case, update_impl):
    g = gu.load(case)
    agent = G.discrete_agent_from_golden(g)
    src = "buf" if "buf.states" in g else "rollout"
    buffer = [G.cuda(g[f"{src}.{k}"]) for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")]
    agent.last_state = G.cuda(g[f"{src}.last_state"])
    agent._inject_ids = G.cuda(g["update_net.ids"])
    result = agent.update_net(buffer)
    G.assert_close(np.array(result), g["update_net.result"], RTOL, 1e-6)
    _check_params(agent, g, "update_net.after")
    assert np.array_equal(buffer[4].cpu().numpy(), g["gae.undones_after"])


@pytest.mark.parametrize("case", gu.DISCRETE_CASES)
def test_discrete_packed_minibatches(case, update_impl):
    """Env-sharded form: records carry the action index as a float in the first action slot."""
    g = gu.load(case)
    agent = G.discrete_agent_from_golden(g)
    lib = _lib.load()
    ids = G.cuda(g["update.ids"])
    updates, batch = ids.shape
    keys = ("buf.states", "buf.actions", "buf.unmasks", "buf.logprobs", "gae.adv_norm", "gae.reward_sums")
    states, actions, unmasks, logprobs, advantages, reward_sums = [G.cuda(g[k]) for k in keys]
    h, n = states.shape[:2]
    tb = _lib.TrainBuffer(states=states.data_ptr(), actions=actions.data_ptr(), unmasks=unmasks.data_ptr(),
                          logprobs=logprobs.data_ptr(), advantages=advantages.data_ptr(), reward_sums=reward_sums.data_ptr(),
                          adv_stats=None, horizon_len=h, num_envs=n, discrete_actions=1)
    rec = ((agent.state_dim + agent.action_dim + 3) & ~3) + 4
    records = th.full((updates * batch, rec), float("nan"), device="cuda:0")
    _lib.check(lib.b200rl_pack_minibatches(C.byref(tb), agent.state_dim, agent.action_dim, batch, updates, ids.data_ptr(), 0, 0,
                                           records.data_ptr(), None))
    rec_np = records.cpu().numpy()
    ids0, ids1 = po.split_ids(g["update.ids"].reshape(-1), h)
    np.testing.assert_array_equal(rec_np[:, agent.state_dim], g["buf.actions"][ids0, ids1].astype(np.float32))
    assert np.isfinite(rec_np).all()
    act_desc, cri_desc = agent._net_desc(agent.act), agent._net_desc(agent.cri)
    act_adam, cri_adam = agent._adam_desc(agent.act_optimizer, agent.act), agent._adam_desc(agent.cri_optimizer, agent.cri)
    ws = agent._get_workspace(act_desc, cri_desc)
    packed = _lib.TrainBuffer(states=records.data_ptr(), horizon_len=0, num_envs=records.shape[0], discrete_actions=1)
    hp = _lib.PPOHyper(ratio_clip=agent.ratio_clip, lambda_entropy=agent.lambda_entropy, clip_grad_norm=agent.clip_grad_norm)
    out = th.empty(3, device="cuda:0")
    _lib.check(lib.b200rl_ppo_update(C.byref(act_desc), C.byref(cri_desc), C.byref(act_adam), C.byref(cri_adam), C.byref(packed),
                                     C.byref(hp), batch, updates, None, 0, 0, out.data_ptr(), ws.data_ptr(), ws.numel(), None))
    G.assert_close(out, g["update.scalars"].mean(axis=0), RTOL, 1e-6)
    _check_params(agent, g, "update.after")


@pytest.mark.parametrize("case", gu.CARTPOLE_CASES)
def test_discrete_rollout_cartpole_against_reference(case):
    """explore_env on the torch CartPole vec env (external-env path: one policy-step kernel per step around env.step) with
    the sampler's and the env's noise replayed, then update_net on what it returned."""
    from elegantrl_b200.envs import CartPoleVecEnv
    g = gu.load(case)
    agent = G.discrete_agent_from_golden(g)
    n, h = int(g["dims"][2]), int(g["dims"][3])
    agent.if_vec_env = True
    env = CartPoleVecEnv(num_envs=n, gpu_id=0, max_step=int(g["max_step"]))
    env.state, env.cur_step = G.cuda(g["env.state0"]), G.cuda(g["env.cur_step0"])
    env.inject_reset_noise(G.cuda(g["env.reset_noise"]))
    agent.last_state = env.state.clone()
    agent._inject_eps = G.cuda(g["expo"])
    states, actions, logprobs, rewards, undones, unmasks = agent.explore_env(env, h)
    assert actions.dtype == th.int32 and tuple(actions.shape) == (h, n)
    assert np.array_equal(actions.cpu().numpy(), g["rollout.actions"])
    assert np.array_equal(undones.cpu().numpy(), g["rollout.undones"]) and np.array_equal(unmasks.cpu().numpy(), g["rollout.unmasks"])
    assert np.array_equal(env.cur_step.cpu().numpy(), g["rollout.cur_step"])
    for name, got in (("states", states), ("logprobs", logprobs), ("rewards", rewards)):
        G.assert_close(got, g[f"rollout.{name}"], RTOL, 1e-5, name)
    G.assert_close(agent.last_state, g["rollout.last_state"], RTOL, 1e-5)
    agent._inject_ids = G.cuda(g["update_net.ids"])
    result = agent.update_net([states, actions, logprobs, rewards, undones, unmasks])
    G.assert_close(np.array(result), g["update_net.result"], RTOL, 1e-6)
    _check_params(agent, g, "update_net.after")


def test_discrete_api_errors():
    lib = _lib.load()
    g = gu.load(gu.DISCRETE_CASES[0])
    cont = G.agent_from_golden(gu.load(gu.SYNTH_CASES[0]))
    # a Gaussian actor (with action_std_log) is refused by the categorical entry points
    act_desc = cont._net_desc(cont.act)
    state = th.zeros((4, cont.state_dim), device="cuda:0")
    action = th.zeros(4, dtype=th.int32, device="cuda:0")
    logprob = th.zeros(4, device="cuda:0")
    rc = lib.b200rl_policy_step_discrete(C.byref(act_desc), None, state.data_ptr(), 4, None, 0, 0, 0, action.data_ptr(),
                                         logprob.data_ptr(), None, None)
    assert rc != 0 and b"action_std_log" in lib.b200rl_last_error()
    with pytest.raises(AssertionError):
        G.discrete_agent_from_golden(g).update_net([th.zeros((4, 2, 4), device="cuda:0"), th.zeros((4, 2), device="cuda:0"),
                                                    th.zeros((4, 2), device="cuda:0"), th.zeros((4, 2), device="cuda:0"),
                                                    th.ones((4, 2), dtype=th.bool, device="cuda:0"),
                                                    th.ones((4, 2), dtype=th.bool, device="cuda:0")])  # float actions


# ------------------------------------------------------------------------ CUDA-graph captured external-env rollout
def _fresh_pair(agent_class, env_class, state_dim, action_dim, if_discrete, n, max_step, graph):
    from elegantrl_b200 import Config
    args = Config(agent_class, None, {'env_name': 'x', 'num_envs': n, 'max_step': max_step, 'state_dim': state_dim,
                                      'action_dim': action_dim, 'if_discrete': if_discrete})
    args.net_dims = [32, 16]  # no fused kernel for these dims -> external-env path also for Pendulum
    args.batch_size, args.repeat_times, args.random_seed = 64, 8, 5
    th.manual_seed(3)
    agent = agent_class(args.net_dims, state_dim, action_dim, gpu_id=0, args=args)
    agent.cuda_graph_rollout = graph
    env = env_class(num_envs=n, gpu_id=0, max_step=max_step, seed=11)
    agent.last_state = env.reset()[0]
    return agent, env


@pytest.mark.parametrize("kind", ["cartpole", "pendulum"])
def test_external_env_rollout_cuda_graph(kind):
    """The graph-captured H-step loop (policy-step kernels + the env's torch ops) must reproduce the eager loop on the
    first cycle and keep going on the next ones: env state and observation carried over, fresh policy and env noise."""
    from elegantrl_b200.agents import AgentDiscretePPO, AgentPPO
    from elegantrl_b200.envs import CartPoleVecEnv, PendulumVecEnv
    spec = dict(cartpole=(AgentDiscretePPO, CartPoleVecEnv, 4, 2, True, 12), pendulum=(AgentPPO, PendulumVecEnv, 3, 1, False, 9))[kind]
    n, h = 40, 24
    eager, env_e = _fresh_pair(*spec[:5], n, spec[5], graph=False)
    graphed, env_g = _fresh_pair(*spec[:5], n, spec[5], graph=True)
    buf_e = [t.clone() for t in eager.explore_env(env_e, h)]
    buf_g = [t.clone() for t in graphed.explore_env(env_g, h)]
    for name, a, b in zip(("states", "actions", "logprobs", "rewards", "undones", "unmasks"), buf_e, buf_g):
        if a.dtype in (th.bool, th.int32):
            assert th.equal(a, b), name
        else:
            G.assert_close(b, a.cpu().numpy(), 1e-5, 1e-6, name)
    assert (~buf_g[5]).any()  # truncations happened
    # second cycle: a replay of the same graph
    last = graphed.last_state.clone()
    buf_g2 = [t.clone() for t in graphed.explore_env(env_g, h)]
    assert len(graphed._rollout_graphs) == 1
    assert th.equal(buf_g2[0][0], last)                  # continues from where the first cycle stopped
    assert not th.equal(buf_g2[1], buf_g[1])             # fresh policy noise (device-resident Philox step counter)
    buf_e2 = [t.clone() for t in eager.explore_env(env_e, h)]
    if kind == "pendulum":                               # the eager path keeps its step counter on the host: same stream
        G.assert_close(buf_g2[0], buf_e2[0].cpu().numpy(), 1e-4, 1e-5, "second cycle states")
    result = graphed.update_net(list(graphed.explore_env(env_g, h)))
    assert all(np.isfinite(result))


# ------------------------------------------------------------------------------------------------ A2C
@pytest.mark.parametrize("case", gu.A2C_CASES)
def test_a2c_update_net_against_reference(case, update_impl):
    """AgentA2C (reference AgentPPO.py:252-311) on single-env buffers: the B200RL_PPO_A2C actor objective."""
    from elegantrl_b200 import Config
    from elegantrl_b200.agents import AgentA2C
    g = gu.load(case)
    dims = [int(x) for x in g["dims"]]
    hp = gu.hyper_of(g)
    args = Config(AgentA2C, None, {'env_name': 'golden', 'num_envs': 1, 'max_step': 200, 'state_dim': dims[0],
                                   'action_dim': dims[1], 'if_discrete': False})
    args.net_dims = dims[4:]
    for k in ("gamma", "ratio_clip", "lambda_entropy", "clip_grad_norm", "learning_rate", "reward_scale", "batch_size",
              "repeat_times", "lambda_gae_adv", "if_use_v_trace"):
        setattr(args, k, hp[k])
    agent = AgentA2C(args.net_dims, dims[0], dims[1], gpu_id=0, args=args)
    G.load_module(agent.act, gu.net_of(g, "actor"))
    G.load_module(agent.cri, gu.net_of(g, "critic"))
    buffer = [G.cuda(g[f"buf.{k}"]) for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")]
    agent.last_state = G.cuda(g["buf.last_state"])
    agent._inject_ids = G.cuda(g["update_net.ids"])
    result = agent.update_net(buffer)
    G.assert_close(np.array(result), g["update_net.result"], RTOL, 1e-6)
    assert result[2] == 0.0
    for which, module in (("actor", agent.act), ("critic", agent.cri)):
        got, ref = gu.flat_params(G.module_to_net(module)), gu.flat_params(gu.net_of(g, f"update_net.after.{which}"))
        for a, b in zip(got, ref):
            G.assert_close(a, b, RTOL, 2e-6, which)

"""GPU parity: the CUDA path (through the C-ABI, via the drop-in agent) against the goldens minted from the
Python reference and against the numpy oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): rtol 1e-4 for fp32 results (atol stated per check, for values that pass
through zero); masks, indices and in-place mask mutation bit-exact."""
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


# ------------------------------------------------------------------------------------------- nets
@pytest.fixture(params=["tc", "ffma"])
def forward_impl(request, monkeypatch):
    """S -> 64 -> 64 -> OUT GELU nets take the tcgen05 forward kernel (csrc/forward_tc.cu); B200RL_FORWARD=ffma forces the
    CUDA-core kernel (csrc/forward.cu), which every other shape uses anyway."""
    if request.param == "ffma":
        monkeypatch.setenv("B200RL_FORWARD", "ffma")
    else:
        monkeypatch.delenv("B200RL_FORWARD", raising=False)
    return request.param


@pytest.mark.parametrize("case", gu.SYNTH_CASES)
def test_mlp_forward(case, forward_impl):
    g = gu.load(case)
    agent = G.agent_from_golden(g)
    lib = _lib.load()
    state = G.cuda(g["nets.state"])
    rows = state.shape[0]
    act_desc, cri_desc = agent._net_desc(agent.act), agent._net_desc(agent.cri)
    mean = th.empty((rows, agent.action_dim), device="cuda:0")
    _lib.check(lib.b200rl_mlp_forward(C.byref(act_desc), state.data_ptr(), rows, mean.data_ptr(), 0, None))
    G.assert_close(mean, g["nets.actor_mean"], RTOL, 2e-6)
    tanh_a = th.empty_like(mean)
    _lib.check(lib.b200rl_mlp_forward(C.byref(act_desc), state.data_ptr(), rows, tanh_a.data_ptr(), 1, None))
    G.assert_close(tanh_a, g["nets.actor_forward"], RTOL, 2e-6)
    G.assert_close(agent.get_values(state), g["nets.value"], RTOL, 2e-6)
    # the torch modules the Evaluator uses must agree with the engine too
    with th.no_grad():
        G.assert_close(agent.act(state), g["nets.actor_forward"], RTOL, 2e-6)


@pytest.mark.parametrize("rows", [1, 31, 64, 65, 127, 128, 129, 1000, 40001])
@pytest.mark.parametrize("case", ["synth_s8_a2_128x64", "synth_s8_a2_64x64"])
def test_mlp_forward_ragged_rows(rows, case):
    g = gu.load(case)
    agent = G.agent_from_golden(g)
    rng = np.random.default_rng(rows)
    state = rng.standard_normal((rows, 8)).astype(np.float32)
    want = po.critic_value(gu.net_of(g, "critic"), state)
    G.assert_close(agent.get_values(G.cuda(state)), want, RTOL, 2e-6)


def test_mlp_forward_wide_net():
    """Widths > 256 take the 32-sample tile path; odd input width takes the scalar-K path."""
    th.manual_seed(3)
    from elegantrl_b200 import Config
    from elegantrl_b200.agents import AgentPPO
    agent = AgentPPO([320, 96, 40], 11, 3, gpu_id=0, args=Config())
    state = th.randn((77, 11), device="cuda:0")
    want = po.critic_value(G.module_to_net(agent.cri), state.cpu().numpy())
    G.assert_close(agent.get_values(state), want, RTOL, 2e-6)


def test_forward_tc_against_ffma_many_tiles(monkeypatch):
    """The persistent tcgen05 forward kernel (296 CTAs walking 128-row tiles) against the CUDA-core kernel on 100 003 rows
    (782 tiles, ragged tail), S = 11 / A = 3: values, means, and the injected-noise policy step."""
    g = gu.load("synth_s11_a3_64x64")
    agent = G.agent_from_golden(g)
    rng = np.random.default_rng(5)
    rows = 100_003
    state = G.cuda(rng.standard_normal((rows, 11)).astype(np.float32) * 2.0)
    eps = G.cuda(rng.standard_normal((rows, 3)).astype(np.float32))
    res = {}
    for impl in ("tc", "ffma"):
        if impl == "ffma":
            monkeypatch.setenv("B200RL_FORWARD", "ffma")
        res[impl] = (agent.get_values(state).clone(),) + tuple(t.clone() for t in agent._policy_step(state, eps=eps))
    for a, b in zip(res["tc"], res["ffma"]):
        G.assert_close(a, b.cpu().numpy(), RTOL, 1e-5)
    sub = rng.integers(0, rows, 200)
    want = po.critic_value(gu.net_of(g, "critic"), state[sub].cpu().numpy())
    G.assert_close(res["tc"][0][sub], want, RTOL, 2e-6)


@pytest.mark.parametrize("case", gu.SYNTH_CASES)
def test_policy_step_injected_noise(case, forward_impl):
    g = gu.load(case)
    agent = G.agent_from_golden(g)
    state = g["nets.state"]
    rng = np.random.default_rng(7)
    eps = rng.standard_normal((state.shape[0], agent.action_dim)).astype(np.float32)
    action, logprob, env_action = agent._policy_step(G.cuda(state), eps=G.cuda(eps))
    want_action, want_logprob = po.sample_action(gu.net_of(g, "actor"), state, eps)
    G.assert_close(action, want_action, RTOL, 2e-6)
    G.assert_close(logprob, want_logprob, RTOL, 2e-5)
    G.assert_close(env_action, np.tanh(want_action), RTOL, 2e-6)
    # golden: logprob of the golden (state, action) pairs == reference get_logprob_entropy
    mean = po.actor_mean(gu.net_of(g, "actor"), state)
    std = np.exp(g["actor.action_std_log"])
    eps_g = ((g["nets.action"] - mean) / std).astype(np.float32)
    action_g, logprob_g, _ = agent._policy_step(G.cuda(state), eps=G.cuda(eps_g))
    G.assert_close(action_g, g["nets.action"], RTOL, 1e-5)
    # logprob = -(a - mu)^2 / (2 sd^2) - log sd - log sqrt(2 pi): the engine sees the golden action only through the fp32
    # reconstruction eps_g, so (a - mu) carries an absolute error of ~ulp(a) <= 2.4e-7 |a| and the log-prob one of
    # |a - mu| ulp(a) / sd^2 -- an ABSOLUTE floor that does not scale with |logprob| (which crosses zero); rtol as north_star
    diff = np.abs(g["nets.action"] - mean)
    floor = float((diff * np.abs(g["nets.action"]) * 2.4e-7 / std ** 2).sum(axis=1).max()) + 2e-6
    G.assert_close(logprob_g, g["nets.logprob"], RTOL, floor)


@pytest.mark.parametrize("case", ["synth_s8_a2_128x64", "synth_s8_a2_64x64"])
def test_policy_step_philox_statistics(case):
    g = gu.load(case)
    agent = G.agent_from_golden(g)
    rows = 200_000
    state = th.zeros((rows, 8), device="cuda:0")
    a1, lp1, _ = agent._policy_step(state)
    a2, _, _ = agent._policy_step(state)
    actor = gu.net_of(g, "actor")
    mean = po.actor_mean(actor, np.zeros((1, 8), np.float32))[0]
    std = np.exp(actor["action_std_log"][0])
    z = (a1.cpu().numpy() - mean) / std
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.01          # action dims independent
    assert abs(np.corrcoef(z[:-1, 0], z[1:, 0])[0, 1]) < 0.01       # envs independent
    assert abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3.0) < 0.08
    assert not th.equal(a1, a2)                                      # step counter advances the stream
    want_lp = po.gaussian_logprob(np.broadcast_to(mean, a1.shape).astype(np.float32), actor["action_std_log"], a1.cpu().numpy())
    G.assert_close(lp1, want_lp, 1e-4, 1e-4)


# -------------------------------------------------------------------------------------------- GAE
@pytest.mark.parametrize("case", gu.SYNTH_CASES)
@pytest.mark.parametrize("tag", ["gae", "gae_alt"])
def test_gae_against_reference(case, tag):
    g = gu.load(case)
    hp = gu.hyper_of(g)
    v_trace = hp["if_use_v_trace"] if tag == "gae" else not hp["if_use_v_trace"]
    agent = G.agent_from_golden(g, if_use_v_trace=v_trace)
    rewards, undones = G.cuda(g["buf.rewards"]), G.cuda(g["buf.undones"])
    unmasks, values = G.cuda(g["buf.unmasks"]), G.cuda(g[f"{tag}.values"])
    adv, rsum, stat_sums = agent.get_advantages(None, rewards, undones, unmasks, values, G.cuda(g[f"{tag}.last_value"]))
    h, n = rewards.shape
    small = h < 16  # short horizons run the single-segment scan: the reference's op order, bit for bit
    if small:
        assert np.array_equal(adv.cpu().numpy(), g[f"{tag}.advantages"])
        assert np.array_equal(rsum.cpu().numpy(), g[f"{tag}.reward_sums"])
    G.assert_close(adv, g[f"{tag}.advantages"], RTOL, 1e-5)
    G.assert_close(rsum, g[f"{tag}.reward_sums"], RTOL, 1e-5)
    assert np.array_equal(rewards.cpu().numpy(), g[f"{tag}.rewards_after"])   # r[trunc] += V: one fp32 add
    assert np.array_equal(undones.cpu().numpy(), g[f"{tag}.undones_after"])   # mask mutation: bit-exact
    assert np.array_equal(unmasks.cpu().numpy(), g["buf.unmasks"])
    # normalisation statistics and the materialised normalised advantages
    lib = _lib.load()
    stats = th.empty(4, device="cuda:0")
    _lib.check(lib.b200rl_adv_stats(stat_sums.data_ptr(), h * n, ((h + 3) // 4) * ((n + 3) // 4), stats.data_ptr(), None))
    if ((h + 3) // 4) * ((n + 3) // 4) >= 2:
        G.assert_close(stats[:2], [g[f"{tag}.adv_mean"], g[f"{tag}.adv_std"]], RTOL, 1e-6)
        adv_n = adv.clone()
        _lib.check(lib.b200rl_normalize_adv(adv_n.data_ptr(), adv_n.numel(), stats.data_ptr(), None))
        G.assert_close(adv_n, g[f"{tag}.adv_norm"], RTOL, 2e-5)


@pytest.mark.parametrize("shape", [(128, 40000), (7, 40001), (128, 3000), (512, 33), (64, 1), (1, 5), (40, 37888), (128, 65536), (33, 38400)])
@pytest.mark.parametrize("v_trace", [True, False])
def test_gae_sizes_against_oracle(shape, v_trace):
    """Large env counts run the sequential scan (bit-exact vs the oracle); small ones the segmented scan."""
    h, n = shape
    rng = np.random.default_rng(h * 100003 + n)
    rewards = rng.standard_normal((h, n)).astype(np.float32)
    values = rng.standard_normal((h, n)).astype(np.float32)
    last_value = rng.standard_normal((n,)).astype(np.float32)
    terminals = rng.random((h, n)) < 0.03
    truncates = (rng.random((h, n)) < 0.03) & ~terminals
    undones, unmasks = ~terminals, ~truncates
    r_ref, u_ref = rewards.copy(), undones.copy()
    want = po.gae(r_ref, u_ref, unmasks, values, last_value, 0.99, 0.95, v_trace)
    g = gu.load("synth_s3_a1_64x64")
    agent = G.agent_from_golden(g, if_use_v_trace=v_trace, gamma=0.99, lambda_gae_adv=0.95)
    r_gpu, u_gpu = G.cuda(rewards), G.cuda(undones)
    adv, rsum, stat_sums = agent.get_advantages(None, r_gpu, u_gpu, G.cuda(unmasks), G.cuda(values), G.cuda(last_value))
    sequential = n >= 148 * 256 or h < 16  # n % 128 == 0 and h >= 32 additionally take the TMA-staged scan
    if sequential:
        assert np.array_equal(adv.cpu().numpy(), want)
    G.assert_close(adv, want, RTOL, 2e-5)
    G.assert_close(rsum, want + values, RTOL, 2e-5)
    assert np.array_equal(r_gpu.cpu().numpy(), r_ref) and np.array_equal(u_gpu.cpu().numpy(), u_ref)
    sums = stat_sums.cpu().numpy()
    lat = want[::4, ::4].astype(np.float64)
    np.testing.assert_allclose(sums[0], want.astype(np.float64).sum(), rtol=1e-6, atol=1e-3)
    np.testing.assert_allclose(sums[1], lat.sum(), rtol=1e-6, atol=1e-3)
    np.testing.assert_allclose(sums[2], (lat ** 2).sum(), rtol=1e-6, atol=1e-3)


# ----------------------------------------------------------------------------------------- update
@pytest.fixture(params=["tc", "cluster", "multilaunch"])
def update_impl(request, monkeypatch):
    """Both update drivers: one persistent thread-block-cluster launch for all minibatches, and one launch per
    minibatch (what large batches use)."""
    monkeypatch.setenv("B200RL_UPDATE", request.param)
    return request.param


def _adam_tensors(agent, which):
    module, opt = (agent.act, agent.act_optimizer) if which == "actor" else (agent.cri, agent.cri_optimizer)
    from elegantrl_b200.agents.AgentPPO import _trainable
    return [(opt.state[p]["exp_avg"], opt.state[p]["exp_avg_sq"], float(opt.state[p]["step"])) for p in _trainable(module)]


def _run_ppo_update(agent, buffer, ids, pre_normalised=True, stats=None):
    lib = _lib.load()
    states, actions, unmasks, logprobs, advantages, reward_sums = buffer
    h, n = states.shape[:2]
    act_desc, cri_desc = agent._net_desc(agent.act), agent._net_desc(agent.cri)
    act_adam, cri_adam = agent._adam_desc(agent.act_optimizer, agent.act), agent._adam_desc(agent.cri_optimizer, agent.cri)
    ws = agent._get_workspace(act_desc, cri_desc)
    tb = _lib.TrainBuffer(states=states.data_ptr(), actions=actions.data_ptr(), unmasks=unmasks.data_ptr(),
                          logprobs=logprobs.data_ptr(), advantages=advantages.data_ptr(), reward_sums=reward_sums.data_ptr(),
                          adv_stats=None if pre_normalised else stats.data_ptr(), horizon_len=h, num_envs=n)
    hp = _lib.PPOHyper(ratio_clip=agent.ratio_clip, lambda_entropy=agent.lambda_entropy, clip_grad_norm=agent.clip_grad_norm)
    out = th.empty(3, device="cuda:0")
    ids = ids.contiguous()
    _lib.check(lib.b200rl_ppo_update(C.byref(act_desc), C.byref(cri_desc), C.byref(act_adam), C.byref(cri_adam), C.byref(tb),
                                     C.byref(hp), ids.shape[1], ids.shape[0], ids.data_ptr(), 0, 0, out.data_ptr(),
                                     ws.data_ptr(), ws.numel(), None))
    agent._set_adam_step(agent.act_optimizer, agent.act, act_adam.step)
    agent._set_adam_step(agent.cri_optimizer, agent.cri, cri_adam.step)
    return out.cpu().numpy()


@pytest.mark.parametrize("case", gu.SYNTH_CASES)
def test_update_objectives_against_reference(case, update_impl):
    g = gu.load(case)
    buffer = [G.cuda(g[k]) for k in ("buf.states", "buf.actions", "buf.unmasks", "buf.logprobs", "gae.adv_norm", "gae.reward_sums")]
    ids = G.cuda(g["update.ids"])
    # one minibatch
    agent = G.agent_from_golden(g)
    scalars = _run_ppo_update(agent, buffer, ids[:1])
    G.assert_close(scalars, g["update.scalars"][0], RTOL, 1e-6)
    for which, module in (("actor", agent.act), ("critic", agent.cri)):
        got, ref = gu.flat_params(G.module_to_net(module)), gu.flat_params(gu.net_of(g, f"update.after1.{which}"))
        for a, b in zip(got, ref):
            G.assert_close(a, b, RTOL, 1e-6, f"{which} after 1 update")
    # k minibatches in one call
    agent = G.agent_from_golden(g)
    scalars = _run_ppo_update(agent, buffer, ids)
    G.assert_close(scalars, g["update.scalars"].mean(axis=0), RTOL, 1e-6)
    for which, module in (("actor", agent.act), ("critic", agent.cri)):
        got, ref = gu.flat_params(G.module_to_net(module)), gu.flat_params(gu.net_of(g, f"update.after.{which}"))
        for a, b in zip(got, ref):
            G.assert_close(a, b, RTOL, 2e-6, f"{which} after {len(ids)} updates")
        n_layers = int(g[f"{which}.n_layers"])
        names = [x for i in range(n_layers) for x in (f"W{i}", f"b{i}")] + (["action_std_log"] if which == "actor" else [])
        for (m, v, step), name in zip(_adam_tensors(agent, which), names):
            assert step == float(g[f"update.after.{which}_adam.step"])
            # the moments are linear (exp_avg) / quadratic (exp_avg_sq) in the minibatch gradients, each element a sum over
            # the minibatch with cancellation: rtol as north_star, atol = 1e-5 of the tensor's largest entry
            m_ref, v_ref = g[f"update.after.{which}_adam.m.{name}"], g[f"update.after.{which}_adam.v.{name}"]
            G.assert_close(m, m_ref, RTOL, 1e-5 * float(np.abs(m_ref).max()), f"{which} exp_avg {name}")
            G.assert_close(v, v_ref, 2 * RTOL, 1e-5 * float(np.abs(v_ref).max()), f"{which} exp_avg_sq {name}")


@pytest.mark.parametrize("case", gu.SYNTH_CASES)
def test_update_gather_time_normalisation_equals_prenormalised(case):
    g = gu.load(case)
    ids = G.cuda(g["update.ids"])
    stats = G.cuda(np.array([g["gae.adv_mean"], g["gae.adv_std"], 0, 0], dtype=np.float32))
    keys = ("buf.states", "buf.actions", "buf.unmasks", "buf.logprobs", "gae.advantages", "gae.reward_sums")
    agent = G.agent_from_golden(g)
    scalars = _run_ppo_update(agent, [G.cuda(g[k]) for k in keys], ids, pre_normalised=False, stats=stats)
    G.assert_close(scalars, g["update.scalars"].mean(axis=0), RTOL, 1e-6)


@pytest.mark.parametrize("case", gu.SYNTH_CASES + gu.ROLLOUT_CASES)
def test_update_net_against_reference(case, update_impl):
    g = gu.load(case)
    agent = G.agent_from_golden(g)
    src = "buf" if "buf.states" in g else "rollout"
    buffer = [G.cuda(g[f"{src}.{k}"]) for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")]
    agent.last_state = G.cuda(g[f"{src}.last_state"])
    agent._inject_ids = G.cuda(g["update_net.ids"])
    result = agent.update_net(buffer)
    G.assert_close(np.array(result), g["update_net.result"], RTOL, 1e-6)
    for which, module in (("actor", agent.act), ("critic", agent.cri)):
        got, ref = gu.flat_params(G.module_to_net(module)), gu.flat_params(gu.net_of(g, f"update_net.after.{which}"))
        for a, b in zip(got, ref):
            G.assert_close(a, b, RTOL, 2e-6, which)
    info = agent.last_update_info
    G.assert_close(info["values"], g["gae.values"], RTOL, 1e-5)
    G.assert_close(info["advantages"], g["gae.advantages"], RTOL, 2e-5)
    G.assert_close(info["adv_stats"][:2], [g["gae.adv_mean"], g["gae.adv_std"]], RTOL, 1e-6)
    assert np.array_equal(buffer[4].cpu().numpy(), g["gae.undones_after"])  # caller's buffer mutated as the reference does


def test_update_device_sampled_indices():
    """Without injected ids the kernel draws them (Philox): in range, deterministic per (seed, offset), and the
    update equals an injected-ids update with the same indices."""
    g = gu.load("synth_s3_a1_64x64")
    keys = ("buf.states", "buf.actions", "buf.unmasks", "buf.logprobs", "gae.adv_norm", "gae.reward_sums")
    buffer = [G.cuda(g[k]) for k in keys]
    a1, a2 = G.agent_from_golden(g), G.agent_from_golden(g)
    lib = _lib.load()

    def run(agent, seed):
        act_desc, cri_desc = agent._net_desc(agent.act), agent._net_desc(agent.cri)
        act_adam, cri_adam = agent._adam_desc(agent.act_optimizer, agent.act), agent._adam_desc(agent.cri_optimizer, agent.cri)
        ws = agent._get_workspace(act_desc, cri_desc)
        h, n = buffer[0].shape[:2]
        tb = _lib.TrainBuffer(states=buffer[0].data_ptr(), actions=buffer[1].data_ptr(), unmasks=buffer[2].data_ptr(),
                              logprobs=buffer[3].data_ptr(), advantages=buffer[4].data_ptr(), reward_sums=buffer[5].data_ptr(),
                              adv_stats=None, horizon_len=h, num_envs=n)
        hp = _lib.PPOHyper(ratio_clip=0.25, lambda_entropy=0.001, clip_grad_norm=3.0)
        out = th.empty(3, device="cuda:0")
        _lib.check(lib.b200rl_ppo_update(C.byref(act_desc), C.byref(cri_desc), C.byref(act_adam), C.byref(cri_adam), C.byref(tb),
                                         C.byref(hp), 64, 3, None, seed, 5, out.data_ptr(), ws.data_ptr(), ws.numel(), None))
        return out.cpu().numpy()

    s1, s2 = run(a1, 123), run(a2, 123)
    assert np.array_equal(s1, s2) or np.allclose(s1, s2, rtol=1e-5)
    assert np.isfinite(s1).all()
    a3 = G.agent_from_golden(g)
    s3 = run(a3, 124)
    assert not np.allclose(s1, s3)


# ---------------------------------------------------------------------------------------- rollout
@pytest.fixture(params=["ts", "tc", "ffma"])
def rollout_impl(request, monkeypatch):
    """All fused-rollout implementations (tcgen05 with the A operand in tensor memory, tcgen05 with a shared-memory
    ring, FP32 pipe) must agree with the reference."""
    monkeypatch.setenv("B200RL_ROLLOUT", request.param)
    return request.param


@pytest.mark.parametrize("case", gu.ROLLOUT_CASES)
def test_fused_rollout_against_reference(case, rollout_impl):
    from elegantrl_b200.envs import PendulumVecEnv
    g = gu.load(case)
    agent = G.agent_from_golden(g)
    n, h = int(g["dims"][2]), int(g["dims"][3])
    env = PendulumVecEnv(num_envs=n, gpu_id=0, max_step=int(g["max_step"]))
    env.theta, env.theta_dot, env.cur_step = G.cuda(g["env.theta0"]), G.cuda(g["env.theta_dot0"]), G.cuda(g["env.cur_step0"])
    agent._inject_eps = G.cuda(g["eps"])
    agent._inject_reset_noise = G.cuda(g["env.reset_noise"])
    agent.last_state = G.cuda(g["state0"])
    states, actions, logprobs, rewards, undones, unmasks = agent.explore_env(env, h)
    assert states.shape == (h, n, 3) and actions.shape == (h, n, 1) and undones.dtype == th.bool
    assert np.array_equal(undones.cpu().numpy(), g["rollout.undones"])      # masks: bit-exact
    assert np.array_equal(unmasks.cpu().numpy(), g["rollout.unmasks"])
    assert np.array_equal(env.cur_step.cpu().numpy(), g["rollout.cur_step"])
    for name, got in (("states", states), ("actions", actions), ("logprobs", logprobs), ("rewards", rewards)):
        G.assert_close(got, g[f"rollout.{name}"], RTOL, 1e-5, name)
    G.assert_close(agent.last_state, g["rollout.last_state"], RTOL, 1e-5)
    G.assert_close(env.theta, g["rollout.theta"], RTOL, 1e-5)
    G.assert_close(env.theta_dot, g["rollout.theta_dot"], RTOL, 1e-5)
    _, values, last_value = agent._value_cache
    G.assert_close(values, g["gae.values"], RTOL, 1e-5)
    G.assert_close(last_value, g["gae.last_value"], RTOL, 1e-5)
    # and straight into update_net (values come from the rollout kernel's cache)
    agent._inject_ids = G.cuda(g["update_net.ids"])
    result = agent.update_net([states, actions, logprobs, rewards, undones, unmasks])
    G.assert_close(np.array(result), g["update_net.result"], RTOL, 1e-6)


@pytest.mark.parametrize("n", [1, 31, 33, 70, 130, 515, 1100])
def test_fused_rollout_ragged_env_counts(n, rollout_impl):
    """Partial warps / unaligned rows take the scalar store path; must equal the oracle all the same."""
    from elegantrl_b200.envs import PendulumVecEnv
    g = gu.load("rollout_pendulum_n8_h16")
    h = 12
    agent = G.agent_from_golden(g, num_envs=n)
    agent.num_envs, agent.if_vec_env = n, True
    rng = np.random.default_rng(n)
    theta0 = rng.uniform(-3, 3, n).astype(np.float32)
    theta_dot0 = rng.uniform(-1, 1, n).astype(np.float32)
    cur0 = rng.integers(0, 9, n).astype(np.int32)
    eps = rng.standard_normal((h, n, 1)).astype(np.float32)
    reset_u = rng.random((h, n, 2)).astype(np.float32)
    env = PendulumVecEnv(num_envs=n, gpu_id=0, max_step=9)
    env.theta, env.theta_dot, env.cur_step = G.cuda(theta0), G.cuda(theta_dot0), G.cuda(cur0)
    agent._inject_eps, agent._inject_reset_noise = G.cuda(eps), G.cuda(reset_u)
    out = agent.explore_env(env, h)
    want = po.rollout_pendulum(gu.net_of(g, "actor"), gu.net_of(g, "critic"), theta0, theta_dot0, cur0, h, eps, reset_u,
                               float(g["hp.reward_scale"]), 9)
    for name, got in zip(("states", "actions", "logprobs", "rewards"), out[:4]):
        G.assert_close(got, want[name], RTOL, 1e-5, name)
    assert np.array_equal(out[4].cpu().numpy(), want["undones"]) and np.array_equal(out[5].cpu().numpy(), want["unmasks"])
    G.assert_close(agent._value_cache[1], want["values"], RTOL, 1e-5)


def test_fused_rollout_matches_torch_env_stepwise(rollout_impl):
    """The fused kernel and the per-step path (engine policy step + torch env.step) walk the same trajectory when
    fed the same noise -- the vec-env contract the reference's loop relies on (AgentPPO.py:112-123)."""
    from elegantrl_b200.envs import PendulumVecEnv
    g = gu.load("rollout_pendulum_n32_h40")
    n, h, max_step = 64, 24, 10
    rng = np.random.default_rng(5)
    eps = rng.standard_normal((h, n, 1)).astype(np.float32)
    reset_u = rng.random((h + 1, n, 2)).astype(np.float32)
    agent = G.agent_from_golden(g, num_envs=n)
    agent.num_envs, agent.if_vec_env = n, True
    env_a = PendulumVecEnv(num_envs=n, gpu_id=0, max_step=max_step)
    env_a.inject_reset_noise(G.cuda(reset_u))
    state0, _ = env_a.reset()
    env_b = PendulumVecEnv(num_envs=n, gpu_id=0, max_step=max_step)
    env_b.theta, env_b.theta_dot, env_b.cur_step = env_a.theta.clone(), env_a.theta_dot.clone(), env_a.cur_step.clone()
    # per-step path on env_a
    state = state0
    traj_states, traj_rewards, traj_trunc = [], [], []
    for t in range(h):
        action, logprob, env_action = agent._policy_step(state, eps=G.cuda(eps[t]))
        traj_states.append(state)
        state, reward, terminal, truncate, _ = env_a.step(env_action)
        traj_rewards.append(reward * agent.reward_scale)
        traj_trunc.append(truncate)
    # fused path on env_b
    agent._inject_eps, agent._inject_reset_noise = G.cuda(eps), G.cuda(reset_u[1:])
    states, actions, logprobs, rewards, undones, unmasks = agent.explore_env(env_b, h)
    G.assert_close(states, th.stack(traj_states).cpu().numpy(), RTOL, 1e-5)
    G.assert_close(rewards, th.stack(traj_rewards).cpu().numpy(), RTOL, 1e-5)
    assert th.equal(unmasks, ~th.stack(traj_trunc))
    G.assert_close(agent.last_state, state.cpu().numpy(), RTOL, 1e-5)
    assert th.equal(env_a.cur_step, env_b.cur_step)


def test_external_vec_env_path():
    """A vec env the engine has no fused kernel for: per-step policy kernel + env.step."""
    from elegantrl_b200 import Config
    from elegantrl_b200.agents import AgentPPO
    from elegantrl_b200.envs import PendulumVecEnv
    n, h = 48, 20
    args = Config(AgentPPO, None, {'env_name': 'x', 'num_envs': n, 'max_step': 7, 'state_dim': 3, 'action_dim': 1, 'if_discrete': False})
    args.net_dims = [32, 16]           # no fused kernel for these dims -> stepwise path
    args.batch_size, args.repeat_times, args.reward_scale = 64, 8, 0.5
    agent = AgentPPO(args.net_dims, 3, 1, gpu_id=0, args=args)
    env = PendulumVecEnv(num_envs=n, gpu_id=0, max_step=7)
    agent.last_state = env.reset()[0]
    buf = agent.explore_env(env, h)
    states, actions, logprobs, rewards, undones, unmasks = buf
    assert states.shape == (h, n, 3) and actions.shape == (h, n, 1) and logprobs.shape == rewards.shape == (h, n)
    assert undones.dtype == th.bool and unmasks.dtype == th.bool and undones.all() and (~unmasks).sum() > 0
    lp, _ = po.logprob_entropy(G.module_to_net(agent.act), states.reshape(-1, 3).cpu().numpy(), actions.reshape(-1, 1).cpu().numpy())
    G.assert_close(logprobs.reshape(-1), lp, 1e-4, 1e-4)
    before = [p.clone() for p in agent.act.parameters()]
    result = agent.update_net(list(buf))
    assert all(np.isfinite(result))
    assert any(not th.equal(a, b) for a, b in zip(before, agent.act.parameters()))


# --------------------------------------------------------------------------- checkpoint / pickling
def test_checkpoint_roundtrip_and_actor_pickle(tmp_path):
    import pickle
    g = gu.load("synth_s3_a1_64x64")
    agent = G.agent_from_golden(g)
    buffer = [G.cuda(g[f"buf.{k}"]) for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")]
    agent.last_state = G.cuda(g["buf.last_state"])
    agent._inject_ids = G.cuda(g["update_net.ids"])
    agent.update_net(buffer)
    agent.save_or_load_agent(str(tmp_path), if_save=True)
    other = G.agent_from_golden(g)
    other.save_or_load_agent(str(tmp_path), if_save=False)
    for a, b in zip(agent.act.parameters(), other.act.parameters()):
        assert th.equal(a, b)
    p0 = list(other.act.parameters())[0]
    assert float(other.act_optimizer.state[p0]["step"]) == float(agent.act_optimizer.state[list(agent.act.parameters())[0]]["step"])
    actor2 = pickle.loads(pickle.dumps(agent.act))                  # what run.py:288-293 sends through a Pipe
    x = th.randn(5, 3, device="cuda:0")
    with th.no_grad():
        assert th.equal(actor2(x), agent.act(x))
    # training continues after reload and keeps moving the loaded modules
    buffer = [G.cuda(g[f"buf.{k}"]) for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")]
    other.last_state = G.cuda(g["buf.last_state"])
    before = other.act.net[0].weight.clone()
    other.update_net(buffer)
    assert not th.equal(before, other.act.net[0].weight)


@pytest.mark.parametrize("case", gu.SYNTH_CASES)
def test_packed_minibatches_equal_direct_gather(case, update_impl):
    """The env-sharded path packs the sampled transitions into records (b200rl_pack_minibatches), all-gathers them and
    runs the update on the packed buffer: on one rank that must reproduce the reference exactly like the direct path."""
    g = gu.load(case)
    agent = G.agent_from_golden(g)
    lib = _lib.load()
    ids = G.cuda(g["update.ids"])
    updates, batch = ids.shape
    keys = ("buf.states", "buf.actions", "buf.unmasks", "buf.logprobs", "gae.advantages", "gae.reward_sums")
    states, actions, unmasks, logprobs, advantages, reward_sums = [G.cuda(g[k]) for k in keys]
    stats = G.cuda(np.array([g["gae.adv_mean"], g["gae.adv_std"], 0, 0], dtype=np.float32))
    h, n = states.shape[:2]
    tb = _lib.TrainBuffer(states=states.data_ptr(), actions=actions.data_ptr(), unmasks=unmasks.data_ptr(),
                          logprobs=logprobs.data_ptr(), advantages=advantages.data_ptr(), reward_sums=reward_sums.data_ptr(),
                          adv_stats=stats.data_ptr(), horizon_len=h, num_envs=n)
    rec = ((agent.state_dim + agent.action_dim + 3) & ~3) + 4
    records = th.full((updates * batch, rec), float("nan"), device="cuda:0")
    _lib.check(lib.b200rl_pack_minibatches(C.byref(tb), agent.state_dim, agent.action_dim, batch, updates, ids.data_ptr(), 0, 0,
                                           records.data_ptr(), None))
    rec_np = records.cpu().numpy()
    ids0, ids1 = po.split_ids(g["update.ids"].reshape(-1), h)
    np.testing.assert_array_equal(rec_np[:, :agent.state_dim], g["buf.states"][ids0, ids1])       # gathers: bit-exact
    np.testing.assert_array_equal(rec_np[:, rec - 4], g["buf.unmasks"][ids0, ids1].astype(np.float32))
    np.testing.assert_allclose(rec_np[:, rec - 2], g["gae.adv_norm"][ids0, ids1], rtol=1e-5, atol=1e-6)

    act_desc, cri_desc = agent._net_desc(agent.act), agent._net_desc(agent.cri)
    act_adam, cri_adam = agent._adam_desc(agent.act_optimizer, agent.act), agent._adam_desc(agent.cri_optimizer, agent.cri)
    ws = agent._get_workspace(act_desc, cri_desc)
    packed = _lib.TrainBuffer(states=records.data_ptr(), horizon_len=0, num_envs=records.shape[0])
    hp = _lib.PPOHyper(ratio_clip=agent.ratio_clip, lambda_entropy=agent.lambda_entropy, clip_grad_norm=agent.clip_grad_norm)
    out = th.empty(3, device="cuda:0")
    _lib.check(lib.b200rl_ppo_update(C.byref(act_desc), C.byref(cri_desc), C.byref(act_adam), C.byref(cri_adam), C.byref(packed),
                                     C.byref(hp), batch, updates, None, 0, 0, out.data_ptr(), ws.data_ptr(), ws.numel(), None))
    G.assert_close(out, g["update.scalars"].mean(axis=0), RTOL, 1e-6)
    for which, module in (("actor", agent.act), ("critic", agent.cri)):
        got, ref = gu.flat_params(G.module_to_net(module)), gu.flat_params(gu.net_of(g, f"update.after.{which}"))
        for a, b in zip(got, ref):
            G.assert_close(a, b, RTOL, 2e-6, which)


# ------------------------------------------------------------------------------- helloworld variant
@pytest.mark.parametrize("case", gu.HELLOWORLD_CASES)
def test_helloworld_agent_against_reference(case, update_impl):
    """elegantrl_b200.agents.helloworld.AgentPPO (same kernels, variant flags) against the tutorial's
    helloworld_PPO_single_file.AgentPPO.update_net: ReLU nets without state_norm, full-buffer std, SmoothL1 critic with
    the mean(unmask) weight, min/clamp clip, entropy bonus, no grad clipping, 2-D single-env buffer."""
    from elegantrl_b200 import Config
    from elegantrl_b200.agents import helloworld
    g = gu.load(case)
    hp = gu.helloworld_hyper_of(g)
    dims = [int(x) for x in g["dims"]]
    args = Config()
    for k in ("gamma", "lambda_gae_adv", "ratio_clip", "lambda_entropy", "learning_rate", "batch_size", "repeat_times"):
        setattr(args, k, hp[k])
    agent = helloworld.AgentPPO(dims[4:], dims[0], dims[1], gpu_id=0, args=args)
    for module, prefix in ((agent.act, "actor"), (agent.cri, "critic")):
        net = gu.plain_net_of(g, prefix)
        linears = [m for m in module.net if isinstance(m, th.nn.Linear)]
        with th.no_grad():
            for layer, w, b in zip(linears, net["W"], net["b"]):
                layer.weight.copy_(th.from_numpy(w)); layer.bias.copy_(th.from_numpy(b))
            if "action_std_log" in net:
                module.action_std_log.copy_(th.from_numpy(net["action_std_log"]))
    buffer = [G.cuda(g[f"buf.{k}"]) for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")]
    agent.last_state = g["buf.last_state"]          # numpy, as the tutorial keeps it
    agent._inject_ids = G.cuda(g["update_net.ids"])
    result = agent.update_net(buffer)
    G.assert_close(np.array(result), g["update_net.result"], RTOL, 1e-6)
    G.assert_close(buffer[3], g["update_net.rewards_after"], RTOL, 2e-6)
    assert np.array_equal(buffer[4].cpu().numpy(), g["update_net.undones_after"])
    G.assert_close(agent.last_update_info["values"].reshape(-1), g["values"], RTOL, 2e-6)
    for which, module in (("actor", agent.act), ("critic", agent.cri)):
        got = gu.flat_params(po.net_from_torch(module))
        ref = gu.flat_params(gu.plain_net_of(g, f"update_net.after.{which}"))
        for a, b in zip(got, ref):
            G.assert_close(a, b, RTOL, 2e-6, which)
    assert isinstance(agent.last_state, np.ndarray)


def test_helloworld_agent_rollout_plumbing():
    """Tutorial loop shape contract (helloworld_PPO_single_file.py:248-277, 490-517) with a numpy gym-style env."""
    from elegantrl_b200 import Config
    from elegantrl_b200.agents import helloworld

    class NumpyPendulum:  # gym-style single env: numpy in / out, (state, reward, terminal, truncate, info)
        def __init__(self):
            from elegantrl_b200.envs import PendulumVecEnv
            self.inner = PendulumVecEnv(num_envs=1, gpu_id=-1, max_step=30, seed=1)
        def reset(self):
            return self.inner.reset()[0][0].numpy(), {}
        def step(self, action):
            s, r, te, tr, _ = self.inner.step(th.as_tensor(action, dtype=th.float32).reshape(1, 1))
            return s[0].numpy(), float(r[0]), bool(te[0]), bool(tr[0]), {}

    args = Config()
    args.batch_size, args.repeat_times, args.gamma = 32, 4, 0.97
    agent = helloworld.AgentPPO([64, 32], 3, 1, gpu_id=0, args=args)
    env = NumpyPendulum()
    agent.last_state = env.reset()[0]
    buf = agent.explore_env(env, 64)
    shapes = [tuple(t.shape) for t in buf]
    assert shapes == [(64, 3), (64, 1), (64,), (64, 1), (64, 1), (64, 1)]
    assert float(buf[2].abs().max()) == 0.0 and buf[4].dtype == th.bool and int((~buf[5]).sum()) == 2
    res = agent.update_net(buf)
    assert len(res) == 3 and np.isfinite(res[:2]).all() and res[2] == 0.0


def test_explore_one_env_against_reference():
    """``AgentPPO._explore_one_env`` (single gym-style env, numpy in / out per step, the agent resets the env
    after a truncated step; reference ``elegantrl/agents/AgentPPO.py:34-85``) against a golden minted from the reference
    on the same numpy env with the policy noise and the env's reset noise replayed, then ``update_net`` on that buffer."""
    from elegantrl_b200.envs import PendulumEnv
    g = gu.load("oneenv_pendulum_h48")
    agent = G.agent_from_golden(g)
    assert not agent.if_vec_env
    h = int(g["dims"][3])
    env = PendulumEnv(max_step=int(g["max_step"]))
    env.inner.inject_reset_noise(th.from_numpy(g["env.reset_noise"]))
    state, _ = env.reset()
    assert np.array_equal(state, g["state0"])
    agent.last_state = th.as_tensor(state, dtype=th.float32, device=agent.device).unsqueeze(0)
    agent._inject_eps = G.cuda(g["eps"])
    states, actions, logprobs, rewards, undones, unmasks = agent.explore_env(env, h)
    assert states.shape == (h, 1, 3) and actions.shape == (h, 1, 1) and logprobs.shape == (h, 1) and rewards.shape == (h, 1)
    assert undones.dtype == th.bool and unmasks.dtype == th.bool and all(t.is_cuda for t in (states, rewards, undones))
    assert np.array_equal(undones.cpu().numpy(), g["rollout.undones"])      # masks: bit-exact
    assert np.array_equal(unmasks.cpu().numpy(), g["rollout.unmasks"]) and int((~unmasks).sum()) >= 2
    for name, got in (("states", states), ("actions", actions), ("logprobs", logprobs), ("rewards", rewards)):
        G.assert_close(got, g[f"rollout.{name}"], RTOL, 1e-5, name)
    G.assert_close(agent.last_state, g["rollout.last_state"], RTOL, 1e-5)
    agent._inject_ids = G.cuda(g["update_net.ids"])
    result = agent.update_net([states, actions, logprobs, rewards, undones, unmasks])
    G.assert_close(np.array(result), g["update_net.result"], RTOL, 1e-6)
    for which, module in (("actor", agent.act), ("critic", agent.cri)):
        got, ref = gu.flat_params(G.module_to_net(module)), gu.flat_params(gu.net_of(g, f"update_net.after.{which}"))
        for a, b in zip(got, ref):
            G.assert_close(a, b, RTOL, 2e-6, which)


def test_fused_rollout_long_horizon(rollout_impl):
    """Many steps through the mbarrier / ring phase logic of the persistent kernels: masks and step counters exact for
    all 300 steps, values against the oracle for the first 48 (before chaotic error growth matters)."""
    from elegantrl_b200.envs import PendulumVecEnv
    g = gu.load("rollout_pendulum_n8_h16")
    n, h, max_step = 130, 300, 37
    agent = G.agent_from_golden(g, num_envs=n)
    agent.num_envs, agent.if_vec_env = n, True
    rng = np.random.default_rng(99)
    theta0 = rng.uniform(-3, 3, n).astype(np.float32)
    theta_dot0 = rng.uniform(-1, 1, n).astype(np.float32)
    cur0 = rng.integers(0, max_step, n).astype(np.int32)
    eps = rng.standard_normal((h, n, 1)).astype(np.float32)
    reset_u = rng.random((h, n, 2)).astype(np.float32)
    env = PendulumVecEnv(num_envs=n, gpu_id=0, max_step=max_step)
    env.theta, env.theta_dot, env.cur_step = G.cuda(theta0), G.cuda(theta_dot0), G.cuda(cur0)
    agent._inject_eps, agent._inject_reset_noise = G.cuda(eps), G.cuda(reset_u)
    states, actions, logprobs, rewards, undones, unmasks = agent.explore_env(env, h)
    t_idx = np.arange(1, h + 1, dtype=np.int64)[:, None]
    want_trunc = ((cur0[None, :].astype(np.int64) + t_idx) % max_step) == 0
    assert np.array_equal(unmasks.cpu().numpy(), ~want_trunc) and bool(undones.all())
    assert np.array_equal(env.cur_step.cpu().numpy(), (cur0 + h) % max_step)
    want = po.rollout_pendulum(gu.net_of(g, "actor"), gu.net_of(g, "critic"), theta0, theta_dot0, cur0, 48, eps[:48], reset_u[:48],
                               float(g["hp.reward_scale"]), max_step)
    for name, got in zip(("states", "actions", "logprobs", "rewards"), (states, actions, logprobs, rewards)):
        G.assert_close(got[:48], want[name], RTOL, 2e-5, name)
    assert th.isfinite(states).all() and th.isfinite(agent._value_cache[1]).all()
    # every reset row must hold exactly the injected reset noise (bit-exact data path through the kernel)
    tr = np.argwhere(want_trunc[:-1])
    got_thd = states[1:, :, 2].cpu().numpy()[tr[:, 0], tr[:, 1]]
    np.testing.assert_array_equal(got_thd, (reset_u[tr[:, 0], tr[:, 1], 1] * np.float32(2.0) - np.float32(1.0)))


@pytest.mark.parametrize("batch", [148 * 32, 9472 + 32, 20000])
def test_large_minibatch_against_oracle(batch):
    """Minibatches above 296 tiles: every CTA of b200rl_ppo_update's per-minibatch kernel walks several 32-sample tiles and
    sums its weight gradients in shared memory before one RED.ADD per element.  Two updates against the numpy oracle."""
    g = gu.load("synth_s3_a1_64x64")
    agent = G.agent_from_golden(g, batch_size=batch)
    rng = np.random.default_rng(batch)
    h, n = 64, 512
    states = rng.standard_normal((h, n, 3)).astype(np.float32)
    actions = (0.7 * rng.standard_normal((h, n, 1))).astype(np.float32)
    actor, critic = gu.net_of(g, "actor"), gu.net_of(g, "critic")
    logprobs = (po.logprob_entropy(actor, states.reshape(-1, 3), actions.reshape(-1, 1))[0].reshape(h, n)
                + 0.05 * rng.standard_normal((h, n))).astype(np.float32)
    unmasks = rng.random((h, n)) > 0.05
    adv = rng.standard_normal((h, n)).astype(np.float32)
    rsum = rng.standard_normal((h, n)).astype(np.float32)
    ids = rng.integers(0, h * n, size=(2, batch)).astype(np.int64)
    buffer = [G.cuda(x) for x in (states, actions, unmasks, logprobs, adv, rsum)]
    scalars = _run_ppo_update(agent, buffer, G.cuda(ids))
    hp = gu.hyper_of(g)
    opt_a, opt_c = po.new_adam_state(actor, True), po.new_adam_state(critic, False)
    buf = dict(states=states, actions=actions, unmasks=unmasks, logprobs=logprobs, advantages=adv, reward_sums=rsum)
    want = [po.ppo_minibatch(actor, critic, opt_a, opt_c, po.gather_minibatch(buf, ids[u]), hp)[0] for u in range(2)]
    G.assert_close(scalars, np.mean(np.array(want), axis=0), RTOL, 2e-6)
    for which, module, net in (("actor", agent.act, actor), ("critic", agent.cri, critic)):
        for a, b in zip(gu.flat_params(G.module_to_net(module)), gu.flat_params(net)):
            G.assert_close(a, b, RTOL, 2e-6, which)


def test_net_descriptor_follows_parameter_storage():
    """The cached C descriptor of a net is revalidated by the parameters' current addresses: an in-place change (what
    ``load_state_dict`` and the engine's own Adam do) keeps it, re-pointing a parameter at new storage (``module.to()``,
    ``p.data = ...``) rebuilds it -- the engine must never read a stale storage."""
    g = gu.load("synth_s8_a2_64x64")
    agent = G.agent_from_golden(g)
    state = G.cuda(g["nets.state"])
    v0 = agent.get_values(state).clone()
    desc0 = agent._net_desc(agent.cri)
    with th.no_grad():
        agent.cri.net[0].weight.mul_(1.5)                      # in place: same storage, same descriptor
    assert agent._net_desc(agent.cri) is desc0
    with th.no_grad():
        want = agent.cri(state).reshape(-1)
    G.assert_close(agent.get_values(state), want.cpu().numpy(), RTOL, 2e-6)
    assert not th.allclose(agent.get_values(state), v0)
    agent.cri.net[2].bias.data = agent.cri.net[2].bias.data.clone() + 0.25   # new storage: descriptor rebuilt
    assert agent._net_desc(agent.cri) is not desc0
    with th.no_grad():
        want = agent.cri(state).reshape(-1)
    G.assert_close(agent.get_values(state), want.cpu().numpy(), RTOL, 2e-6)

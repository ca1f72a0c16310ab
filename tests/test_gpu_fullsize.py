"""BASELINE.json full size (65 536 envs x 128 steps, 2x64 nets): size-independent properties of the CUDA path.

The oracle is too slow for 8.4 M env-steps, so the full-size run is checked through
  * self-consistency between kernels (fused-rollout values == generic critic kernel on the stored states;
    stored logprobs == Gaussian log-density of the stored actions under the generic actor kernel),
  * the oracle on a random subset of env columns (envs are independent, so a column subset is a complete problem),
  * invariants (masks are exactly the complement of the truncation schedule, determinism under a fixed seed).
"""
import numpy as np
import pytest
import torch as th

from elegantrl_b200 import Config
from elegantrl_b200.agents import AgentPPO
from elegantrl_b200.envs import PendulumVecEnv
from oracle import ppo_oracle as po
from tests import gpu_utils as G

pytestmark = pytest.mark.gpu
N, H, MAX_STEP = 65536, 128, 200


def make(seed=0):
    env_args = {'env_name': 'Pendulum-v1', 'num_envs': N, 'max_step': MAX_STEP, 'state_dim': 3, 'action_dim': 1, 'if_discrete': False}
    cfg = Config(AgentPPO, PendulumVecEnv, env_args)
    cfg.net_dims, cfg.batch_size, cfg.repeat_times, cfg.random_seed = [64, 64], 128, 8.0, seed
    th.manual_seed(seed)
    agent = AgentPPO([64, 64], 3, 1, gpu_id=0, args=cfg)
    env = PendulumVecEnv(num_envs=N, gpu_id=0, max_step=MAX_STEP, seed=seed)
    agent.last_state = env.reset()[0]
    env.cur_step[:] = (th.arange(N, device="cuda:0", dtype=th.int32) * 7) % MAX_STEP
    return agent, env


@pytest.fixture(params=["ts", "tc", "ffma"])
def rollout_impl(request, monkeypatch):
    monkeypatch.setenv("B200RL_ROLLOUT", request.param)
    return request.param


def test_full_size_cycle_properties(rollout_impl):
    agent, env = make()
    theta0, theta_dot0, cur0 = env.theta.clone(), env.theta_dot.clone(), env.cur_step.clone()
    states, actions, logprobs, rewards, undones, unmasks = agent.explore_env(env, H)
    _, values, last_value = agent._value_cache
    assert states.shape == (H, N, 3) and values.shape == (H, N)
    for t in (states, actions, logprobs, rewards, values, last_value, agent.last_state):
        assert th.isfinite(t).all()
    # masks: bit-exact complement of the truncation schedule; Pendulum never terminates
    t_idx = th.arange(1, H + 1, device="cuda:0", dtype=th.int32)[:, None]
    want_trunc = ((cur0[None, :] + t_idx) % MAX_STEP) == 0
    assert th.equal(unmasks, ~want_trunc) and bool(undones.all())
    assert th.equal(env.cur_step, (cur0 + H) % MAX_STEP)
    # observations lie on the unit circle; |theta_dot| <= 8
    assert float((states[..., 0] ** 2 + states[..., 1] ** 2 - 1).abs().max()) < 1e-5 and float(states[..., 2].abs().max()) <= 8.0
    # fused critic == generic critic kernel on the stored states; logprob == log-density of the stored action
    G.assert_close(values, agent.get_values(states).cpu().numpy(), 1e-4, 2e-5)
    G.assert_close(last_value, agent.get_values(agent.last_state).cpu().numpy(), 1e-4, 2e-5)
    cols = th.randperm(N, device="cuda:0")[:96]
    sub_s, sub_a = states[:, cols].reshape(-1, 3).cpu().numpy(), actions[:, cols].reshape(-1, 1).cpu().numpy()
    lp, _ = po.logprob_entropy(G.module_to_net(agent.act), sub_s, sub_a)
    G.assert_close(logprobs[:, cols].reshape(-1), lp, 1e-4, 1e-4)
    # env dynamics on the subset: stored state[t+1] == oracle step from stored state[t], action[t] (where not reset)
    sub = states[:, cols].cpu().numpy(); act = actions[:, cols, 0].cpu().numpy(); rew = rewards[:, cols].cpu().numpy()
    theta = np.arctan2(sub[:, :, 1], sub[:, :, 0]).astype(np.float32)
    for t in (0, 17, 63, 126):
        th_t, thd_t = theta[t], sub[t, :, 2]
        _, nthd, _, r, _, _ = po.pendulum_step(th_t, thd_t, np.zeros(96, np.int32), np.tanh(act[t]), np.zeros((96, 2), np.float32), 10 ** 9)
        keep = unmasks[t, cols].cpu().numpy()
        np.testing.assert_allclose(sub[t + 1, keep, 2], nthd[keep], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(rew[t], r * float(agent.reward_scale), rtol=1e-4, atol=2e-5)

    # GAE at full size: bit-exact vs the oracle on a column subset (sequential scan), statistics vs float64 numpy
    r_np, u_np, m_np, v_np = (x[:, cols].cpu().numpy().copy() for x in (rewards, undones, unmasks, values))
    want_adv = po.gae(r_np, u_np, m_np, v_np, last_value[cols].cpu().numpy(), agent.gamma, agent.lambda_gae_adv, True)
    result = agent.update_net([states, actions, logprobs, rewards, undones, unmasks])
    info = agent.last_update_info
    assert np.array_equal(info["advantages"][:, cols].cpu().numpy(), want_adv)
    adv = info["advantages"].cpu().numpy().astype(np.float64)
    G.assert_close(info["adv_stats"][:2], [adv.mean(), adv[::4, ::4].std(ddof=1)], 1e-5, 1e-6)
    assert all(np.isfinite(result)) and info["update_times"] == 8
    assert abs(result[2] - (0.5 + 0.5 * np.log(2 * np.pi))) < 0.2      # entropy of sigma ~ 1 policy


def test_full_size_rollout_is_deterministic(rollout_impl):
    out = []
    for _ in range(2):
        agent, env = make(seed=3)
        buf = agent.explore_env(env, H)
        out.append([t.clone() for t in buf] + [agent.last_state.clone(), agent._value_cache[1].clone()])
    for a, b in zip(*out):
        assert th.equal(a, b)
    agent, env = make(seed=4)
    assert not th.equal(agent.explore_env(env, H)[1], out[0][1])   # a different seed gives different noise


def test_both_rollout_kernels_agree_at_full_size(monkeypatch):
    res = {}
    for mode in ("tc", "ts", "ffma"):
        monkeypatch.setenv("B200RL_ROLLOUT", mode)
        agent, env = make(seed=5)
        buf = agent.explore_env(env, 16)
        res[mode] = [t.float().cpu().numpy() for t in buf] + [agent._value_cache[1].cpu().numpy()]
    for other in ("tc", "ts"):
        for a, b in zip(res[other], res["ffma"]):
            np.testing.assert_allclose(a, b, rtol=1e-4, atol=2e-5)


def test_fused_rollout_native_noise_statistics(rollout_impl):
    """The fused kernels draw their N(0,1) / U(0,1) numbers from the same ``rollout_noise()`` as the per-step kernel but with
    the counter laid out as (env, step_offset + t): recover the noise from the stored trajectory (z = (a - mu(s)) / sigma, mu
    from torch's fp32 MLP) and check its moments, its independence across envs, across consecutive time steps and
    across consecutive rollouts (the step offset advances by H), and the reset draws (uniform in [-pi, pi) x [-1, 1))."""
    agent, env = make(seed=11)
    sd = float(th.exp(agent.act.action_std_log[0]))
    zs = []
    for _ in range(2):
        states, actions, logprobs, _, _, unmasks = agent.explore_env(env, H)
        with th.no_grad():   # torch fp32 mean on the stored states (TF32 off by default for matmul)
            mu = agent.act.net(agent.act.state_norm(states.reshape(-1, 3))).reshape(H, N)
        zs.append(((actions[..., 0] - mu) / sd).double())
        trunc = ~unmasks
    z = zs[0]
    assert abs(float(z.mean())) < 2e-3 and abs(float(z.std()) - 1.0) < 2e-3
    assert abs(float((z ** 3).mean())) < 6e-3 and abs(float((z ** 4).mean()) - 3.0) < 2e-2
    corr = lambda a, b: float(((a - a.mean()) * (b - b.mean())).mean() / (a.std() * b.std()))
    assert abs(corr(z[:-1], z[1:])) < 2e-3            # consecutive time steps of the same env
    assert abs(corr(z[:, :-1], z[:, 1:])) < 2e-3      # neighbouring envs at the same step
    assert abs(corr(z[:-7, :-5], z[7:, 5:])) < 2e-3   # a diagonal of the (step, env) counter lattice
    assert abs(corr(zs[0], zs[1])) < 2e-3             # the next rollout continues the stream (step_offset += H)
    assert not th.equal(zs[0], zs[1])
    # tail mass: P(|z| > 3) = 2.6998e-3
    assert abs(float((z.abs() > 3).double().mean()) - 2.6998e-3) < 2e-4
    # reset draws of the second rollout: the state after a truncation is (cos, sin)(theta0), theta_dot0 with theta0 ~ U(-pi, pi), theta_dot0 ~ U(-1, 1)
    t_i, n_i = th.nonzero(trunc[:-1], as_tuple=True)
    nxt = states[t_i + 1, n_i]
    theta0, thd0 = th.atan2(nxt[:, 1], nxt[:, 0]).double(), nxt[:, 2].double()
    assert len(theta0) > 20000
    assert abs(float(theta0.mean())) < 0.05 and abs(float(theta0.std()) - np.pi / np.sqrt(3)) < 0.03
    assert abs(float(thd0.mean())) < 0.02 and abs(float(thd0.std()) - 1 / np.sqrt(3)) < 0.01 and float(thd0.abs().max()) <= 1.0
    assert abs(corr(theta0, thd0)) < 0.02

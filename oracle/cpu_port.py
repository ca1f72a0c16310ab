"""PyTorch-CPU port of the reference's on-policy cycle, used as the CPU BASELINE that bench.py times.
TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by the product package.

The reference itself is Python + PyTorch (no native code), so its "own CPU implementation of the path" is the
sequence of ATen CPU kernels its agent launches.  This module restates that sequence op for op -- same
launch granularity, same multi-threaded ATen kernels, same autograd / clip_grad_norm_ / torch.optim.Adam calls
-- so that its wall clock is the reference's wall clock on the same host (the real reference cannot travel to
the GPU box).  Cited lines are in the reference's ``elegantrl/agents/AgentPPO.py`` unless noted.
``tests/test_cpu_port_golden.py`` pins it to the goldens minted from the real reference.
"""
import time

import torch as th
from torch.nn.utils import clip_grad_norm_

from elegantrl_b200.agents.nets import ActorPPO, CriticPPO
from elegantrl_b200.envs import PendulumVecEnv


class CpuPPO:
    def __init__(self, net_dims, state_dim, action_dim, num_envs, *, gamma=0.99, lambda_gae_adv=0.95, ratio_clip=0.25,
                 lambda_entropy=0.001, clip_grad_norm=3.0, learning_rate=6e-5, reward_scale=1.0, batch_size=128,
                 repeat_times=8.0, if_use_v_trace=True, device="cpu"):
        self.device = th.device(device)  # "cuda:0" turns the same op sequence into the eager PyTorch-CUDA baseline
        self.state_dim, self.action_dim, self.num_envs = state_dim, action_dim, num_envs
        self.gamma, self.lambda_gae_adv, self.ratio_clip = gamma, lambda_gae_adv, ratio_clip
        self.lambda_entropy = th.tensor(lambda_entropy, dtype=th.float32, device=self.device)
        self.clip_grad_norm, self.reward_scale = clip_grad_norm, reward_scale
        self.batch_size, self.repeat_times, self.if_use_v_trace = batch_size, repeat_times, if_use_v_trace
        self.act = ActorPPO(net_dims, state_dim, action_dim).to(self.device)
        self.cri = CriticPPO(net_dims, state_dim, action_dim).to(self.device)
        self.act_optimizer = th.optim.Adam(self.act.parameters(), learning_rate)
        self.cri_optimizer = th.optim.Adam(self.cri.parameters(), learning_rate)
        self.criterion = th.nn.MSELoss(reduction="none")
        self.last_state = None

    # ---- rollout: Python loop of H steps, ~45 ATen launches per step (:87-129, get_action :368-376)
    def explore_env(self, env, horizon_len):
        n = self.num_envs
        dev = self.device
        states = th.zeros((horizon_len, n, self.state_dim), dtype=th.float32).to(dev)  # (:101-107 allocate on the host first)
        actions = th.zeros((horizon_len, n, self.action_dim), dtype=th.float32).to(dev)
        logprobs = th.zeros((horizon_len, n), dtype=th.float32).to(dev)
        rewards = th.zeros((horizon_len, n), dtype=th.float32).to(dev)
        terminals = th.zeros((horizon_len, n), dtype=th.bool).to(dev)
        truncates = th.zeros((horizon_len, n), dtype=th.bool).to(dev)
        state = self.last_state
        with th.no_grad():
            for t in range(horizon_len):
                mean = self.act.net(self.act.state_norm(state))
                dist = th.distributions.normal.Normal(mean, self.act.action_std_log.exp())
                action = dist.sample()
                logprob = dist.log_prob(action).sum(1)
                states[t], actions[t], logprobs[t] = state, action, logprob
                state, reward, terminal, truncate, _ = env.step(action.tanh())
                rewards[t], terminals[t], truncates[t] = reward, terminal, truncate
        self.last_state = state
        rewards *= self.reward_scale
        return states, actions, logprobs, rewards, th.logical_not(terminals), th.logical_not(truncates)

    # ---- GAE: boolean-mask fix-up + Python reverse loop (:207-232)
    def get_advantages(self, states, rewards, undones, unmasks, values):
        advantages = th.empty_like(values)
        truncated = th.logical_not(unmasks)
        if th.any(truncated):
            rewards[truncated] += self.cri(states[truncated]).squeeze(1).detach()
            undones[truncated] = False
        masks = undones * self.gamma
        next_value = self.cri(self.last_state.clone()).detach().squeeze(-1)
        advantage = th.zeros_like(next_value)
        if self.if_use_v_trace:
            for t in range(rewards.shape[0] - 1, -1, -1):
                next_value = rewards[t] + masks[t] * next_value
                advantages[t] = advantage = next_value - values[t] + masks[t] * self.lambda_gae_adv * advantage
                next_value = values[t]
        else:
            for t in range(rewards.shape[0] - 1, -1, -1):
                advantages[t] = rewards[t] - values[t] + masks[t] * advantage
                advantage = values[t] + self.lambda_gae_adv * advantages[t]
        return advantages

    def _step(self, optimizer, objective):  # AgentBase.optimizer_backward, AgentBase.py:239-248
        optimizer.zero_grad()
        objective.backward()
        clip_grad_norm_(parameters=optimizer.param_groups[0]["params"], max_norm=self.clip_grad_norm)
        optimizer.step()

    # ---- one minibatch (:173-205)
    def update_objectives(self, buffer):
        states, actions, unmasks, logprobs, advantages, reward_sums = buffer
        sample_len, num_seqs = states.shape[0], states.shape[1]
        ids = th.randint(sample_len * num_seqs, size=(self.batch_size,), requires_grad=False, device=self.device)
        ids0 = th.fmod(ids, sample_len)
        ids1 = th.div(ids, sample_len, rounding_mode='floor')
        state, action, unmask = states[ids0, ids1], actions[ids0, ids1], unmasks[ids0, ids1]
        logprob, advantage, reward_sum = logprobs[ids0, ids1], advantages[ids0, ids1], reward_sums[ids0, ids1]

        value = self.cri(state).squeeze(1)
        obj_critic = (self.criterion(value, reward_sum) * unmask).mean()
        self._step(self.cri_optimizer, obj_critic)

        mean = self.act.net(self.act.state_norm(state))
        dist = th.distributions.normal.Normal(mean, self.act.action_std_log.exp())
        new_logprob, entropy = dist.log_prob(action).sum(1), dist.entropy().sum(1)
        ratio = (new_logprob - logprob.detach()).exp()
        surrogate = advantage * ratio * th.where(advantage.gt(0), 1 - self.ratio_clip, 1 + self.ratio_clip)
        obj_surrogate = (surrogate * unmask).mean()
        obj_entropy = (entropy * unmask).mean()
        self._step(self.act_optimizer, -(obj_surrogate - obj_entropy * self.lambda_entropy))
        return obj_critic.item(), obj_surrogate.item(), obj_entropy.item()

    # ---- whole update (:135-171)
    def update_net(self, buffer):
        buffer_size = buffer[0].shape[0]
        with th.no_grad():
            states, actions, logprobs, rewards, undones, unmasks = buffer
            bs = max(1, 2 ** 10 // self.num_envs)
            values = th.cat([self.cri(states[i:i + bs]) for i in range(0, buffer_size, bs)], dim=0).squeeze(-1)
            advantages = self.get_advantages(states, rewards, undones, unmasks, values)
            reward_sums = advantages + values
            advantages = (advantages - advantages.mean()) / (advantages[::4, ::4].std() + 1e-5)
        train_buffer = states, actions, unmasks, logprobs, advantages, reward_sums
        update_times = int(buffer_size * self.repeat_times / self.batch_size)
        assert update_times >= 1
        logs = []
        with th.enable_grad():
            for _ in range(update_times):
                logs.append(self.update_objectives(train_buffer))
        logs = th.tensor(logs, dtype=th.float64)
        return tuple(logs.mean(dim=0).tolist())


def time_cpu_cycles(num_envs, horizon_len, net_dims=(64, 64), warmup=1, cycles=3, threads=None, seed=0, device="cpu", **hyper):
    """Time explore_env + update_net of the port on the torch Pendulum vec env, all host threads (``device="cuda:0"``:
    the same eager op sequence on the GPU, synchronised around each phase).
    Returns dict(env_steps_per_sec, explore_s, update_s, cycle_s (lists), threads)."""
    if threads is None:
        threads = th.get_num_threads()
    th.set_num_threads(threads)
    th.manual_seed(seed)
    on_gpu = th.device(device).type == "cuda"
    sync = th.cuda.synchronize if on_gpu else (lambda: None)
    agent = CpuPPO(list(net_dims), 3, 1, num_envs, device=device, **hyper)
    env = PendulumVecEnv(num_envs=num_envs, gpu_id=th.device(device).index or 0 if on_gpu else -1, max_step=200, seed=seed)
    agent.last_state = env.reset()[0]
    explore_s, update_s = [], []
    for i in range(warmup + cycles):
        sync()
        t0 = time.perf_counter()
        buffer = agent.explore_env(env, horizon_len)
        sync()
        t1 = time.perf_counter()
        agent.update_net(list(buffer))
        sync()
        t2 = time.perf_counter()
        if i >= warmup:
            explore_s.append(t1 - t0)
            update_s.append(t2 - t1)
    total = sum(explore_s) + sum(update_s)
    return dict(env_steps_per_sec=num_envs * horizon_len * cycles / total, explore_s=explore_s, update_s=update_s,
                threads=threads, cycles=cycles)


def best_thread_count(net_dims=(64, 64), candidates=None, **hyper):
    """PyTorch's intra-op threading scales NEGATIVELY on this workload beyond a handful of threads (a 64-thread run
    was measured 8x slower than an 8-thread one): pick the thread count that maximises env-steps/s on a reduced
    cycle (16 384 envs x 32 steps) so that the CPU baseline is the best the host can do, not the worst."""
    import os
    ncpu = os.cpu_count() or 8
    if candidates is None:
        # (one thread per logical CPU is pathological on big hosts -- 128 threads measured 400x slower than 16 -- and
        # would make the calibration itself take a minute, so the sweep stops at 64)
        candidates = sorted({t for t in (4, 8, 16, 32, 64, min(ncpu, 64)) if t <= ncpu})
    scores = {}
    for t in candidates:
        r = time_cpu_cycles(16384, 32, net_dims, warmup=1, cycles=1, threads=t, **hyper)
        scores[t] = r["env_steps_per_sec"]
    best = max(scores, key=scores.get)
    th.set_num_threads(best)
    return best, scores

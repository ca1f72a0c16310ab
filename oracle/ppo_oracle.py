"""CPU oracle for the on-policy hot path (rollout -> GAE -> PPO update).  TEST INFRASTRUCTURE ONLY.

A numpy restatement (closed form, hand-written backward, no autograd, no torch) of the arithmetic of the
reference's on-policy agent.  Every function cites the reference file:line it follows (paths relative to
the reference checkout, ``elegantrl/...``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` leg of ``bench.py`` may import this module; the product path (``elegantrl_b200``) never
does, and fails loudly when ``libb200rl.so`` is missing.

Pinning: the reference's own tests hold no numeric goldens for this path, so the oracle
is pinned against vectors minted *from the importable Python reference itself* by ``oracle/make_golden.py``
(committed under ``tests/golden/``) -- see ``tests/test_oracle_golden.py``.

Conventions
-----------
* a "net" is a dict ``{"W": [W0, W1, ...], "b": [b0, b1, ...], "state_avg", "state_std",
  "action_std_log" (actor only, shape [1, A]), "activation": "gelu" | "relu"}``; ``W[l]`` has the
  ``nn.Linear.weight`` layout ``[out, in]``.
* dtype: everything is computed in the dtype of the inputs (float32 to mirror the reference, float64 for
  a higher-precision yardstick); integer / mask paths are exact.
"""
import math

import numpy as np
from scipy.special import erf as _erf

SQRT_HALF = math.sqrt(0.5)
INV_SQRT_2PI = 1.0 / math.sqrt(2.0 * math.pi)
LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))


# ----------------------------------------------------------------------------------------------- nets
def act_fn(z, activation):
    """nn.GELU() (exact erf form) / nn.ReLU -- reference AgentBase.py:353-354 (build_mlp default GELU)."""
    if activation == "gelu":
        return (z * 0.5 * (1.0 + _erf(z * SQRT_HALF))).astype(z.dtype)
    if activation == "relu":
        return np.maximum(z, 0).astype(z.dtype)
    raise ValueError(activation)


def act_grad(z, activation):
    """d act / d z.  GELU' = Phi(z) + z * phi(z)."""
    if activation == "gelu":
        cdf = 0.5 * (1.0 + _erf(z * SQRT_HALF))
        pdf = np.exp(-0.5 * z * z) * INV_SQRT_2PI
        return (cdf + z * pdf).astype(z.dtype)
    if activation == "relu":
        return (z > 0).astype(z.dtype)
    raise ValueError(activation)


def state_norm(net, state):
    """(s - avg) / (std + 1e-4) -- reference AgentPPO.py:360-361 (actor) / :440-441 (critic)."""
    dt = state.dtype
    if net.get("state_avg") is None:
        return state
    return ((state - net["state_avg"].astype(dt)) / (net["state_std"].astype(dt) + dt.type(1e-4))).astype(dt)


def mlp_forward(net, x, keep=False):
    """Linear-act-...-Linear on an already-normalised input x [B, in] -- reference AgentBase.py:345-360."""
    acts, pre = [x], []
    h = x
    n_layers = len(net["W"])
    for l in range(n_layers):
        z = (h @ net["W"][l].astype(x.dtype).T + net["b"][l].astype(x.dtype)).astype(x.dtype)
        pre.append(z)
        h = act_fn(z, net["activation"]) if l < n_layers - 1 else z
        acts.append(h)
    return (h, acts, pre) if keep else h


def mlp_backward(net, acts, pre, d_out):
    """Hand-written backward of ``mlp_forward``: returns (dW list, db list, d_input)."""
    n_layers = len(net["W"])
    dW, db = [None] * n_layers, [None] * n_layers
    dz = d_out
    for l in range(n_layers - 1, -1, -1):
        dW[l] = (dz.T @ acts[l]).astype(dz.dtype)
        db[l] = dz.sum(axis=0).astype(dz.dtype)
        dx = (dz @ net["W"][l].astype(dz.dtype)).astype(dz.dtype)
        if l > 0:
            dz = (dx * act_grad(pre[l - 1], net["activation"])).astype(dz.dtype)
    return dW, db, dx


def critic_value(critic, state):
    """CriticPPO.forward -- reference AgentPPO.py:435-438.  state [..., S] -> value [...]."""
    lead = state.shape[:-1]
    v = mlp_forward(critic, state_norm(critic, state.reshape(-1, state.shape[-1])))
    return v.reshape(lead)


def actor_mean(actor, state):
    """action_avg = net(state_norm(s)) -- reference AgentPPO.py:369-370."""
    return mlp_forward(actor, state_norm(actor, state))


def actor_forward(actor, state):
    """ActorPPO.forward (deterministic, tanh) -- reference AgentPPO.py:363-366."""
    return np.tanh(actor_mean(actor, state))


def gaussian_logprob(mean, std_log, action):
    """Normal(mean, exp(std_log)).log_prob(action).sum(1) -- reference AgentPPO.py:371-375 +
    torch.distributions.Normal.log_prob: -((x-mu)^2)/(2 var) - log(scale) - log(sqrt(2 pi))."""
    dt = mean.dtype
    std = np.exp(std_log.astype(dt))
    var = std * std
    lp = -((action - mean) ** 2) / (2 * var) - np.log(std) - dt.type(LOG_SQRT_2PI)
    return lp.sum(axis=1).astype(dt)


def gaussian_entropy(std_log, batch, dt):
    """Normal.entropy().sum(1) = sum_a (0.5 + 0.5 log 2pi + log scale) -- reference AgentPPO.py:385."""
    std = np.exp(std_log.astype(dt))
    ent = (dt.type(0.5 + 0.5 * math.log(2 * math.pi)) + np.log(std)).sum()
    return np.full((batch,), ent, dtype=dt)


def sample_action(actor, state, eps):
    """ActorPPO.get_action with the noise injected: a = mu + sigma * eps -- reference AgentPPO.py:368-376
    (torch draws Normal.sample() as randn * sigma + mu)."""
    mean = actor_mean(actor, state)
    std = np.exp(actor["action_std_log"].astype(mean.dtype))
    action = (eps.astype(mean.dtype) * std + mean).astype(mean.dtype)
    return action, gaussian_logprob(mean, actor["action_std_log"], action)


def logprob_entropy(actor, state, action):
    """ActorPPO.get_logprob_entropy -- reference AgentPPO.py:378-386."""
    mean = actor_mean(actor, state)
    return (gaussian_logprob(mean, actor["action_std_log"], action),
            gaussian_entropy(actor["action_std_log"], state.shape[0], mean.dtype))


# ------------------------------------------------------------------------------- discrete (categorical) actor
def categorical_from_logits(z):
    """softmax + torch.distributions.Categorical(probs=...) -- reference AgentPPO.py:407-418 (ActorDiscretePPO):
    ``a_prob = Softmax(net(s))``; Categorical renormalises ``probs / probs.sum(-1)`` and keeps
    ``logits = log(clamp(probs, eps, 1 - eps))`` (torch/distributions/utils.py probs_to_logits).  Returns (probs, logits)."""
    dt = z.dtype
    e = np.exp(z - z.max(axis=1, keepdims=True))
    p = (e / e.sum(axis=1, keepdims=True)).astype(dt)
    p = (p / p.sum(axis=1, keepdims=True)).astype(dt)
    eps = np.finfo(dt).eps
    return p, np.log(np.clip(p, eps, 1 - eps)).astype(dt)


def categorical_sample(probs, expo):
    """Categorical.sample() = torch.multinomial(probs, 1): ATen's one-draw fast path is the exponential race
    ``argmax(probs / q)``, q ~ Exp(1) (aten/src/ATen/native/Sampling / multinomial "fast path").  ``expo`` [B, A] injects q."""
    return np.argmax(probs / expo.astype(probs.dtype), axis=1).astype(np.int64)


def sample_action_discrete(actor, state, expo):
    """ActorDiscretePPO.get_action with the Exp(1) noise injected -- reference AgentPPO.py:407-413."""
    p, logits = categorical_from_logits(actor_mean(actor, state))
    action = categorical_sample(p, expo)
    return action, np.take_along_axis(logits, action[:, None], axis=1)[:, 0]


def logprob_entropy_discrete(actor, state, action):
    """ActorDiscretePPO.get_logprob_entropy -- reference AgentPPO.py:415-421 (entropy = -sum p * logits)."""
    p, logits = categorical_from_logits(actor_mean(actor, state))
    logprob = np.take_along_axis(logits, action.astype(np.int64)[:, None], axis=1)[:, 0]
    return logprob, (-(p * logits).sum(axis=1)).astype(p.dtype)


# ------------------------------------------------------------------------------------------------ env
def pendulum_observe(theta, theta_dot):
    return np.stack((np.cos(theta), np.sin(theta), theta_dot), axis=1).astype(theta.dtype)


def pendulum_step(theta, theta_dot, cur_step, env_action, reset_u, max_step=200):
    """One step of ``elegantrl_b200.envs.PendulumVecEnv.step`` (gymnasium Pendulum-v1 physics with the
    scaling of reference elegantrl/envs/CustomGymEnv.py:39-44: torque = 2 * action, reward * 0.5;
    truncation at max_step, auto-reset -- contract of reference elegantrl/train/config.py:243-247).
    env_action [N] already tanh'ed.  reset_u [N, 2] in [0, 1).  Returns new (theta, theta_dot, cur_step),
    reward, terminal, truncate."""
    dt = theta.dtype
    f = dt.type
    torque = np.clip(env_action.astype(dt) * f(2.0), f(-2.0), f(2.0))
    theta_norm = np.remainder(theta + f(math.pi), f(2 * math.pi)) - f(math.pi)
    cost = theta_norm * theta_norm + f(0.1) * (theta_dot * theta_dot) + f(0.001) * (torque * torque)
    reward = (cost * f(-0.5)).astype(dt)
    accel = f(15.0) * np.sin(theta) + f(3.0) * torque
    new_theta_dot = np.clip(theta_dot + accel * f(0.05), f(-8.0), f(8.0)).astype(dt)
    new_theta = (theta + new_theta_dot * f(0.05)).astype(dt)
    cur_step = cur_step + 1
    truncate = cur_step >= max_step
    terminal = np.zeros_like(truncate)
    u = reset_u.astype(dt)
    new_theta = np.where(truncate, (u[:, 0] * f(2.0) - f(1.0)) * f(math.pi), new_theta).astype(dt)
    new_theta_dot = np.where(truncate, u[:, 1] * f(2.0) - f(1.0), new_theta_dot).astype(dt)
    cur_step = np.where(truncate, 0, cur_step).astype(np.int32)
    return new_theta, new_theta_dot, cur_step, reward, terminal, truncate


def rollout_pendulum(actor, critic, theta, theta_dot, cur_step, horizon_len, eps, reset_u,
                     reward_scale=1.0, max_step=200):
    """AgentPPO._explore_vec_env on the Pendulum vec env -- reference AgentPPO.py:87-129: record pre-step
    state, raw (pre-tanh) action, logprob; env gets tanh(action); rewards *= reward_scale;
    undones = ~terminals; unmasks = ~truncates.  Also returns V(s_t) of the pre-update critic
    (what reference update_net :141-143 recomputes) and the final env state."""
    dt = theta.dtype
    n = theta.shape[0]
    a_dim = actor["W"][-1].shape[0]
    states = np.zeros((horizon_len, n, 3), dt)
    actions = np.zeros((horizon_len, n, a_dim), dt)
    logprobs = np.zeros((horizon_len, n), dt)
    rewards = np.zeros((horizon_len, n), dt)
    terminals = np.zeros((horizon_len, n), bool)
    truncates = np.zeros((horizon_len, n), bool)
    values = np.zeros((horizon_len, n), dt)
    for t in range(horizon_len):
        state = pendulum_observe(theta, theta_dot)
        action, logprob = sample_action(actor, state, eps[t])
        states[t], actions[t], logprobs[t] = state, action, logprob
        if critic is not None:
            values[t] = critic_value(critic, state)
        theta, theta_dot, cur_step, reward, terminal, truncate = pendulum_step(
            theta, theta_dot, cur_step, np.tanh(action[:, 0]), reset_u[t], max_step)
        rewards[t], terminals[t], truncates[t] = reward, terminal, truncate
    rewards = (rewards * dt.type(reward_scale)).astype(dt)
    out = dict(states=states, actions=actions, logprobs=logprobs, rewards=rewards,
               undones=~terminals, unmasks=~truncates, values=values,
               last_state=pendulum_observe(theta, theta_dot), theta=theta, theta_dot=theta_dot, cur_step=cur_step)
    return out


def cartpole_step(state, cur_step, action, reset_u, max_step=500):
    """One step of ``elegantrl_b200.envs.CartPoleVecEnv.step`` (classic cart-pole, Euler, tau 0.02; terminal beyond
    +-2.4 m / +-12 deg, truncation at max_step, auto-reset to U(-0.05, 0.05) -- vec-env contract of reference
    elegantrl/train/config.py:243-247).  action [N] int.  Returns (state, cur_step, reward, terminal, truncate)."""
    dt = state.dtype
    f = dt.type
    x, x_dot, theta, theta_dot = (state[:, i] for i in range(4))
    force = np.where(action > 0, f(10.0), f(-10.0)).astype(dt)
    cos_t, sin_t = np.cos(theta), np.sin(theta)
    temp = (force + f(0.1 * 0.5) * (theta_dot * theta_dot) * sin_t) / f(1.1)
    theta_acc = (f(9.8) * sin_t - cos_t * temp) / (f(0.5) * (f(4.0 / 3.0) - f(0.1) * (cos_t * cos_t) / f(1.1)))
    x_acc = temp - f(0.1 * 0.5) * theta_acc * cos_t / f(1.1)
    x = x + f(0.02) * x_dot
    x_dot = x_dot + f(0.02) * x_acc
    theta = theta + f(0.02) * theta_dot
    theta_dot = theta_dot + f(0.02) * theta_acc
    new_state = np.stack((x, x_dot, theta, theta_dot), axis=1).astype(dt)
    cur_step = cur_step + 1
    terminal = (np.abs(x) > f(2.4)) | (np.abs(theta) > f(12 * 2 * math.pi / 360))
    truncate = (cur_step >= max_step) & ~terminal
    done = terminal | truncate
    fresh = (reset_u.astype(dt) * f(0.1) - f(0.05)).astype(dt)
    new_state = np.where(done[:, None], fresh, new_state).astype(dt)
    cur_step = np.where(done, 0, cur_step).astype(np.int32)
    return new_state, cur_step, np.ones(state.shape[0], dt), terminal, truncate


def rollout_cartpole(actor, critic, state, cur_step, horizon_len, expo, reset_u, reward_scale=1.0, max_step=500):
    """AgentPPO._explore_vec_env, discrete branch (actions int32 [H, N]) -- reference AgentPPO.py:87-129 with
    ActorDiscretePPO.get_action (:407-413) and convert_action_for_env = .long() (:423-425)."""
    dt = state.dtype
    n = state.shape[0]
    states = np.zeros((horizon_len, n, 4), dt)
    actions = np.zeros((horizon_len, n), np.int32)
    logprobs = np.zeros((horizon_len, n), dt)
    rewards = np.zeros((horizon_len, n), dt)
    terminals = np.zeros((horizon_len, n), bool)
    truncates = np.zeros((horizon_len, n), bool)
    values = np.zeros((horizon_len, n), dt)
    for t in range(horizon_len):
        action, logprob = sample_action_discrete(actor, state, expo[t])
        states[t], actions[t], logprobs[t] = state, action, logprob
        if critic is not None:
            values[t] = critic_value(critic, state)
        state, cur_step, reward, terminal, truncate = cartpole_step(state, cur_step, action, reset_u[t], max_step)
        rewards[t], terminals[t], truncates[t] = reward, terminal, truncate
    rewards = (rewards * dt.type(reward_scale)).astype(dt)
    return dict(states=states, actions=actions, logprobs=logprobs, rewards=rewards, undones=~terminals,
                unmasks=~truncates, values=values, last_state=state, cur_step=cur_step)


# ------------------------------------------------------------------------------------------------ GAE
def gae(rewards, undones, unmasks, values, last_value, gamma, lambda_gae, if_use_v_trace=True,
        trunc_values=None):
    """AgentPPO.get_advantages -- reference AgentPPO.py:207-232.

    (i) truncation fix-up, IN PLACE on the caller's arrays (:211-214): rewards[trunc] += V(s_trunc);
        undones[trunc] = False.  ``trunc_values`` = V at the stored (pre-reset) state; it equals ``values``
        because the same critic produced both (:143 vs :213).
    (ii) masks = undones * gamma (:216); (iii) next_value = V(last_state) (:219-220);
    (iv) reverse scan, default branch :223-227, alternative branch :228-231."""
    dt = values.dtype
    if trunc_values is None:
        trunc_values = values
    truncated = np.logical_not(unmasks)
    if truncated.any():
        rewards[truncated] += trunc_values[truncated]
        undones[truncated] = False
    masks = (undones * dt.type(gamma)).astype(dt)
    horizon_len = rewards.shape[0]
    advantages = np.empty_like(values)
    next_value = last_value.astype(dt).copy()
    advantage = np.zeros_like(next_value)
    lam = dt.type(lambda_gae)
    if if_use_v_trace:
        for t in range(horizon_len - 1, -1, -1):
            next_value = rewards[t] + masks[t] * next_value
            advantages[t] = advantage = next_value - values[t] + masks[t] * lam * advantage
            next_value = values[t]
    else:
        for t in range(horizon_len - 1, -1, -1):
            advantages[t] = rewards[t] - values[t] + masks[t] * advantage
            advantage = values[t] + lam * advantages[t]
    return advantages


def advantage_stats(advantages):
    """mean over everything, unbiased std over the [::4, ::4] sub-lattice -- reference AgentPPO.py:149."""
    mean = advantages.mean(dtype=np.float64)
    std = advantages[::4, ::4].std(ddof=1, dtype=np.float64)
    return mean, std


def normalize_advantages(advantages, full_std=False):
    """(adv - adv.mean()) / (adv[::4, ::4].std() + 1e-5) -- reference AgentPPO.py:149; full_std: the helloworld
    variant (adv - mean) / (adv.std() + 1e-5), helloworld_PPO_single_file.py:296."""
    dt = advantages.dtype
    mean, std = advantage_stats(advantages)
    if full_std:
        std = advantages.std(ddof=1, dtype=np.float64)
    return ((advantages - dt.type(mean)) / (dt.type(std) + dt.type(1e-5))).astype(dt)


def values_gae_pass(critic, states, rewards, undones, unmasks, last_state, gamma, lambda_gae,
                    if_use_v_trace=True):
    """First half of AgentPPO.update_net -- reference AgentPPO.py:139-150: values, advantages (mutating
    rewards/undones), reward_sums = adv + values, normalised advantages."""
    values = critic_value(critic, states)
    last_value = critic_value(critic, last_state)
    advantages = gae(rewards, undones, unmasks, values, last_value, gamma, lambda_gae, if_use_v_trace)
    reward_sums = (advantages + values).astype(values.dtype)
    return values, advantages, reward_sums, normalize_advantages(advantages)


# --------------------------------------------------------------------------------------------- update
def split_ids(ids, horizon_len):
    """ids0 = ids % H (time), ids1 = ids // H (env) -- reference AgentPPO.py:178-180.  Exact (int64)."""
    return np.fmod(ids, horizon_len), ids // horizon_len


def new_adam_state(net, has_std):
    st = {"step": 0, "m_W": [np.zeros_like(w) for w in net["W"]], "v_W": [np.zeros_like(w) for w in net["W"]],
          "m_b": [np.zeros_like(b) for b in net["b"]], "v_b": [np.zeros_like(b) for b in net["b"]]}
    if has_std:
        st["m_std"] = np.zeros_like(net["action_std_log"])
        st["v_std"] = np.zeros_like(net["action_std_log"])
    return st


def clip_grads(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ -- called per net by reference AgentBase.py:239-248:
    coef = max_norm / (total_norm + 1e-6), clamped to <= 1, multiplied into every grad."""
    dt = grads[0].dtype
    total_norm = dt.type(math.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads)))
    if max_norm is None or max_norm <= 0:
        return grads, total_norm
    coef = min(dt.type(max_norm) / (total_norm + dt.type(1e-6)), dt.type(1.0))
    return [(g * coef).astype(dt) for g in grads], total_norm


def adam_step(params, grads, ms, vs, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (defaults: no weight decay / amsgrad), op order of torch's ``_single_tensor_adam``:
    m.lerp_(g, 1-b1); v = v*b2 + (1-b2) g*g; denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= lr/(1-b1^t) * m/denom.
    Instantiated by reference AgentPPO.py:24-25, stepped by AgentBase.py:248."""
    step = step + 1
    bc1 = 1.0 - beta1 ** step
    bc2_sqrt = math.sqrt(1.0 - beta2 ** step)
    step_size = lr / bc1
    for p, g, m, v in zip(params, grads, ms, vs):
        dt = p.dtype
        m += (g - m) * dt.type(1.0 - beta1)
        v *= dt.type(beta2)
        v += dt.type(1.0 - beta2) * g * g
        denom = np.sqrt(v) / dt.type(bc2_sqrt) + dt.type(eps)
        p -= dt.type(step_size) * (m / denom)
    return step


def ppo_minibatch(actor, critic, opt_a, opt_c, batch, hp):
    """AgentPPO.update_objectives on an already-gathered minibatch -- reference AgentPPO.py:189-205 +
    AgentBase.optimizer_backward :239-248.  ``batch`` = dict(state, action, unmask (bool), logprob,
    advantage (normalised), reward_sum).  ``hp`` = dict(ratio_clip, lambda_entropy, clip_grad_norm,
    learning_rate).  Mutates nets and Adam states in place.  Returns the three logged scalars and a dict of
    the (clipped) gradients for inspection."""
    state, action = batch["state"], batch["action"]
    dt = state.dtype
    unmask = batch["unmask"].astype(dt)
    bsz = dt.type(state.shape[0])

    # critic: obj_critic = mean(MSE(V(s), reward_sum) * unmask)   (:189-191)
    xc = state_norm(critic, state)
    value, acts_c, pre_c = mlp_forward(critic, xc, keep=True)
    err = value[:, 0] - batch["reward_sum"]
    if hp.get("critic_loss", "mse") == "smooth_l1":  # helloworld_PPO_single_file.py:246 SmoothL1Loss(beta=1)
        small = np.abs(err) < 1
        loss_el = np.where(small, dt.type(0.5) * err * err, np.abs(err) - dt.type(0.5)).astype(dt)
        d_el = np.where(small, err, np.sign(err)).astype(dt)
    else:
        loss_el, d_el = err * err, dt.type(2.0) * err
    if hp.get("critic_mask", "elementwise") == "batch_mean":
        # helloworld multiplies criterion(...) [B] by unmask [B, 1]: a [B, B] outer product whose mean is
        # mean(loss) * mean(unmask)  (helloworld_PPO_single_file.py:325, 332)
        um_c = np.full_like(unmask, unmask.mean(dtype=dt))
    else:
        um_c = unmask
    obj_critic = (loss_el * um_c).mean(dtype=dt)
    d_value = (d_el * um_c / bsz)[:, None].astype(dt)
    dWc, dbc, _ = mlp_backward(critic, acts_c, pre_c, d_value)
    gc = [g for pair in zip(dWc, dbc) for g in pair]
    gc, norm_c = clip_grads(gc, hp["clip_grad_norm"])

    # actor: ratio, "clip" as a constant factor, entropy sign as in the reference   (:193-204)
    xa = state_norm(actor, state)
    mean, acts_a, pre_a = mlp_forward(actor, xa, keep=True)
    discrete = hp.get("discrete", False)
    if discrete:  # ActorDiscretePPO.get_logprob_entropy (:415-421): `mean` holds the logits, `action` the int indices
        probs, logits = categorical_from_logits(mean)
        act_idx = np.asarray(action).astype(np.int64).reshape(-1)
        new_logprob = np.take_along_axis(logits, act_idx[:, None], axis=1)[:, 0]
        entropy = (-(probs * logits).sum(axis=1)).astype(dt)
    else:
        std_log = actor["action_std_log"].astype(dt)
        std = np.exp(std_log)
        var = std * std
        diff = action - mean
        new_logprob = (-(diff ** 2) / (2 * var) - np.log(std) - dt.type(LOG_SQRT_2PI)).sum(axis=1)
        entropy = gaussian_entropy(std_log, state.shape[0], dt)
    ratio = np.exp(new_logprob - batch["logprob"])
    adv = batch["advantage"]
    mask_a = unmask if hp.get("mask_actor", True) else np.ones_like(unmask)  # helloworld does not mask the actor terms
    if hp.get("surrogate", "factor") == "a2c":
        # AgentA2C.update_objectives (AgentPPO.py:306-311), single-env buffers [H, 1, ...]: new_logprob is summed over the
        # size-1 env axis, stays [B, A], and obj_actor = mean over [B, A] of adv * logprob = mean_b(adv_b * logp_b) / A
        inv_a = dt.type(1.0 / mean.shape[1])
        surrogate = (adv * new_logprob * inv_a).astype(dt)
        ratio = np.ones_like(adv)
        d_ratio = (adv * inv_a).astype(dt)                                   # d surrogate / d new_logprob
    elif hp.get("surrogate", "factor") == "min_clip":  # helloworld_PPO_single_file.py:337-339
        clipped = np.clip(ratio, dt.type(1 - hp["ratio_clip"]), dt.type(1 + hp["ratio_clip"]))
        s1, s2 = adv * ratio, adv * clipped
        take1 = s1 <= s2
        surrogate = np.where(take1, s1, s2).astype(dt)
        d_ratio = np.where(take1, adv, np.where(clipped == ratio, adv, dt.type(0))).astype(dt)
    else:  # the reference's constant factor, AgentPPO.py:196-199
        kappa = np.where(adv > 0, dt.type(1 - hp["ratio_clip"]), dt.type(1 + hp["ratio_clip"])).astype(dt)
        surrogate = adv * ratio * kappa
        d_ratio = (adv * kappa).astype(dt)
    obj_surrogate = (surrogate * mask_a).mean(dtype=dt)
    obj_entropy = (entropy * mask_a).mean(dtype=dt)
    ent_sign = dt.type(-1.0 if hp.get("entropy_bonus", False) else 1.0)  # d loss / d entropy: helloworld ADDS the bonus
    lam_ent = dt.type(hp["lambda_entropy"])
    # loss = -(obj_surrogate - obj_entropy * lambda_entropy)
    g_logp = (-(d_ratio * ratio * mask_a) / bsz).astype(dt)                 # d loss / d new_logprob
    if discrete:
        # d logp[a] / d z = onehot(a) - p;  d entropy / d z_j = -p_j (log p_j + entropy)   (log-softmax algebra; the
        # eps clamp of probs_to_logits is inactive unless a probability underflows 1.2e-7)
        onehot = np.zeros_like(probs)
        onehot[np.arange(probs.shape[0]), act_idx] = 1
        g_ent = (ent_sign * lam_ent * mask_a / bsz).astype(dt)              # d loss / d entropy_b
        d_mean = (g_logp[:, None] * (onehot - probs) - g_ent[:, None] * probs * (logits + entropy[:, None])).astype(dt)
        dWa, dba, _ = mlp_backward(actor, acts_a, pre_a, d_mean)
        ga = [g for pair in zip(dWa, dba) for g in pair]
    else:
        d_mean = (g_logp[:, None] * diff / var).astype(dt)                  # d logp / d mu = (a - mu) / var
        d_std_log = (g_logp[:, None] * (diff * diff / var - dt.type(1.0))).sum(axis=0, keepdims=True) \
            + ent_sign * lam_ent * mask_a.mean(dtype=dt)                    # d entropy / d std_log = 1
        dWa, dba, _ = mlp_backward(actor, acts_a, pre_a, d_mean)
        ga = [g for pair in zip(dWa, dba) for g in pair] + [d_std_log.astype(dt)]
    ga, norm_a = clip_grads(ga, hp["clip_grad_norm"])

    # Adam, critic first (:191) then actor (:204); the nets are disjoint so the order is immaterial
    pc = [p for pair in zip(critic["W"], critic["b"]) for p in pair]
    mc = [p for pair in zip(opt_c["m_W"], opt_c["m_b"]) for p in pair]
    vc = [p for pair in zip(opt_c["v_W"], opt_c["v_b"]) for p in pair]
    opt_c["step"] = adam_step(pc, gc, mc, vc, opt_c["step"], hp["learning_rate"])
    pa = [p for pair in zip(actor["W"], actor["b"]) for p in pair] + ([] if discrete else [actor["action_std_log"]])
    ma = [p for pair in zip(opt_a["m_W"], opt_a["m_b"]) for p in pair] + ([] if discrete else [opt_a["m_std"]])
    va = [p for pair in zip(opt_a["v_W"], opt_a["v_b"]) for p in pair] + ([] if discrete else [opt_a["v_std"]])
    opt_a["step"] = adam_step(pa, ga, ma, va, opt_a["step"], hp["learning_rate"])
    return (float(obj_critic), float(obj_surrogate), float(obj_entropy)), \
        dict(critic=gc, actor=ga, norm_critic=float(norm_c), norm_actor=float(norm_a))


def gather_minibatch(buffer, ids, adv_mean=None, adv_std=None):
    """The six advanced-index gathers of reference AgentPPO.py:178-187.  ``buffer`` = dict(states [H,N,S],
    actions [H,N,A], unmasks, logprobs, advantages, reward_sums [H,N]).  If adv_mean/adv_std are given the
    advantage is normalised at gather time (what the CUDA engine does) instead of beforehand."""
    horizon_len = buffer["states"].shape[0]
    ids0, ids1 = split_ids(ids, horizon_len)
    adv = buffer["advantages"][ids0, ids1]
    if adv_mean is not None:
        dt = adv.dtype
        adv = ((adv - dt.type(adv_mean)) / (dt.type(adv_std) + dt.type(1e-5))).astype(dt)
    return dict(state=buffer["states"][ids0, ids1], action=buffer["actions"][ids0, ids1],
                unmask=buffer["unmasks"][ids0, ids1], logprob=buffer["logprobs"][ids0, ids1],
                advantage=adv, reward_sum=buffer["reward_sums"][ids0, ids1])


def update_net(actor, critic, opt_a, opt_c, rollout, last_state, ids_per_update, hp):
    """AgentPPO.update_net -- reference AgentPPO.py:135-171, with the minibatch indices injected
    (``ids_per_update`` [update_times, batch_size] int64 replaces th.randint :178).
    update_times = int(H * repeat_times / batch_size) is the caller's business (:159)."""
    states = rollout["states"]
    values, advantages, reward_sums, adv_norm = values_gae_pass(
        critic, states, rollout["rewards"], rollout["undones"], rollout["unmasks"], last_state,
        hp["gamma"], hp["lambda_gae_adv"], hp.get("if_use_v_trace", True))
    buffer = dict(states=states, actions=rollout["actions"], unmasks=rollout["unmasks"],
                  logprobs=rollout["logprobs"], advantages=adv_norm, reward_sums=reward_sums)
    logs = []
    for ids in ids_per_update:
        scalars, _ = ppo_minibatch(actor, critic, opt_a, opt_c, gather_minibatch(buffer, ids), hp)
        logs.append(scalars)
    logs = np.array(logs, dtype=np.float64)
    return tuple(logs.mean(axis=0)), dict(values=values, advantages=advantages, reward_sums=reward_sums,
                                          adv_norm=adv_norm, per_update=logs)


# ------------------------------------------------------------------------------------------ utilities
def net_from_torch(module, dtype=np.float32):
    """Copy a torch ActorPPO / CriticPPO (reference's or this repo's: same attribute names) to a net dict."""
    import torch.nn as nn
    linears = [m for m in module.net if isinstance(m, nn.Linear)]
    activation = "relu" if any(isinstance(m, nn.ReLU) for m in module.net) else "gelu"
    net = {"W": [l.weight.detach().cpu().numpy().astype(dtype).copy() for l in linears],
           "b": [l.bias.detach().cpu().numpy().astype(dtype).copy() for l in linears],
           "activation": activation, "state_avg": None, "state_std": None}
    if getattr(module, "state_avg", None) is not None:
        net["state_avg"] = module.state_avg.detach().cpu().numpy().astype(dtype).copy()
        net["state_std"] = module.state_std.detach().cpu().numpy().astype(dtype).copy()
    if hasattr(module, "action_std_log"):
        net["action_std_log"] = module.action_std_log.detach().cpu().numpy().astype(dtype).copy()
    return net


def net_astype(net, dtype):
    out = dict(net)
    out["W"] = [w.astype(dtype) for w in net["W"]]
    out["b"] = [b.astype(dtype) for b in net["b"]]
    for k in ("state_avg", "state_std", "action_std_log"):
        if net.get(k) is not None:
            out[k] = net[k].astype(dtype)
    return out


# ------------------------------------------------------------------------- env-sharded (multi-GPU) update
def ppo_minibatch_grads(actor, critic, batch, hp, denominator=None):
    """Unclipped gradient SUMS of one rank's share of a minibatch, divided by ``denominator`` (the GLOBAL batch
    size), and this rank's share of the three logged scalars -- what ``b200rl_ppo_grads`` leaves in the flat
    buffer.  Summing the outputs over ranks gives the gradients / scalars of reference AgentPPO.py:189-204 on the
    union minibatch.  Nets are not modified."""
    import copy
    a2, c2 = copy.deepcopy(actor), copy.deepcopy(critic)
    dt = batch["state"].dtype
    local = batch["state"].shape[0]
    denominator = local if denominator is None else denominator
    hp_raw = dict(hp, clip_grad_norm=0.0, learning_rate=0.0)
    scalars, grads = ppo_minibatch(a2, c2, new_adam_state(a2, not hp.get("discrete", False)), new_adam_state(c2, False), batch, hp_raw)
    scale = dt.type(local / denominator)  # ppo_minibatch averaged over the local samples
    return tuple(s * float(scale) for s in scalars), [g * scale for g in grads["actor"]], [g * scale for g in grads["critic"]]


def ppo_apply_grads(actor, critic, opt_a, opt_c, grads_actor, grads_critic, hp):
    """clip_grad_norm_ + Adam on (all-reduced) gradients, per net -- what ``b200rl_ppo_apply`` does on every rank."""
    ga, _ = clip_grads([g.copy() for g in grads_actor], hp["clip_grad_norm"])
    gc, _ = clip_grads([g.copy() for g in grads_critic], hp["clip_grad_norm"])
    pc = [p for pair in zip(critic["W"], critic["b"]) for p in pair]
    mc = [p for pair in zip(opt_c["m_W"], opt_c["m_b"]) for p in pair]
    vc = [p for pair in zip(opt_c["v_W"], opt_c["v_b"]) for p in pair]
    opt_c["step"] = adam_step(pc, gc, mc, vc, opt_c["step"], hp["learning_rate"])
    discrete = hp.get("discrete", False)
    pa = [p for pair in zip(actor["W"], actor["b"]) for p in pair] + ([] if discrete else [actor["action_std_log"]])
    ma = [p for pair in zip(opt_a["m_W"], opt_a["m_b"]) for p in pair] + ([] if discrete else [opt_a["m_std"]])
    va = [p for pair in zip(opt_a["v_W"], opt_a["v_b"]) for p in pair] + ([] if discrete else [opt_a["v_std"]])
    opt_a["step"] = adam_step(pa, ga, ma, va, opt_a["step"], hp["learning_rate"])


def lattice_stat_sums(advantages, env_offset):
    """(sum adv, sum and sum of squares over the [::4, ::4] lattice taken on the GLOBAL env index) of one env shard --
    the three doubles ``b200rl_gae`` emits; all-reduced they give reference AgentPPO.py:149's mean / std."""
    adv = advantages.astype(np.float64)
    cols = (np.arange(adv.shape[1]) + env_offset) % 4 == 0
    lat = adv[::4][:, cols]
    return np.array([adv.sum(), lat.sum(), (lat ** 2).sum(), 0.0])


def stats_from_sums(sums, count_all, count_lattice):
    mean = sums[0] / count_all
    m_lat = sums[1] / count_lattice
    var = (sums[2] - count_lattice * m_lat * m_lat) / (count_lattice - 1.0)
    return mean, math.sqrt(max(var, 0.0))


# ------------------------------------------------------------------------------------- A2C variant
A2C_FLAVOUR = dict(surrogate="a2c", mask_actor=False, lambda_entropy=0.0)


def update_net_a2c(actor, critic, opt_a, opt_c, rollout, last_state, ids_per_update, hp):
    """AgentA2C.update_net / update_objectives -- reference AgentPPO.py:257-311, meaningful for single-env buffers
    ``[H, 1, ...]`` only (its ``states[indices]`` indexes the time axis alone): same values / GAE /
    normalisation pass as PPO, then per minibatch the masked-MSE critic step and ``obj_actor = (advantage * new_logprob).mean()``
    (no ratio, no clip, no entropy term, not masked).  ``ids_per_update`` are time indices.  Returns
    (obj_critic_avg, obj_actor_avg, 0)."""
    assert rollout["states"].shape[1] == 1, "A2C of the reference is only coherent for num_envs == 1"
    hp = dict(hp, **A2C_FLAVOUR)
    (obj_c, obj_a, _), aux = update_net(actor, critic, opt_a, opt_c, rollout, last_state, ids_per_update, hp)
    return (obj_c, obj_a, 0.0), aux


# ------------------------------------------------------------------------------- helloworld variant
HELLOWORLD_FLAVOUR = dict(critic_loss="smooth_l1", critic_mask="batch_mean", surrogate="min_clip", entropy_bonus=True,
                          mask_actor=False, clip_grad_norm=0.0)


def update_net_helloworld(actor, critic, opt_a, opt_c, buf, last_state, ids_per_update, hp):
    """helloworld AgentPPO.update_net -- reference helloworld/helloworld_PPO_single_file.py:283-364: single-env buffer
    (states [H, S], actions [H, A], logprobs [H], rewards / undones / unmasks [H, 1]), ReLU nets without state_norm,
    full-buffer std in the advantage normalisation (:296), SmoothL1 critic (:246, 332), min/clamp clip (:337-339),
    entropy bonus ADDED (:340), actor terms not masked, no grad clipping (:366-370).
    Returns (obj_critic_avg, obj_actor_avg, 0.0) as :314-317 (its a_std_log attribute does not exist -> 0)."""
    states, actions, logprobs, rewards, undones, unmasks = buf
    hp = dict(hp, **HELLOWORLD_FLAVOUR)
    values = critic_value(critic, states)                                   # [H]
    last_value = critic_value(critic, last_state[None, :])                  # [1]
    advantages = gae(rewards, undones, unmasks, values[:, None], last_value, hp["gamma"], hp["lambda_gae_adv"], True)
    reward_sums = (advantages + values[:, None]).astype(values.dtype)
    adv_norm = normalize_advantages(advantages, full_std=True)
    buffer = dict(states=states[:, None, :], actions=actions[:, None, :], unmasks=unmasks, logprobs=logprobs[:, None],
                  advantages=adv_norm, reward_sums=reward_sums)
    logs = []
    for ids in ids_per_update:
        (obj_c, obj_s, obj_e), _ = ppo_minibatch(actor, critic, opt_a, opt_c, gather_minibatch(buffer, ids), hp)
        logs.append((obj_c, obj_s + obj_e * hp["lambda_entropy"]))
    logs = np.array(logs, dtype=np.float64)
    return (logs[:, 0].mean(), logs[:, 1].mean(), 0.0), dict(values=values, advantages=advantages, adv_norm=adv_norm)

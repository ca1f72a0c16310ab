"""CPU oracle for the off-policy path the hot-path scope table lists NEXT: ReplayBuffer + SAC.
TEST INFRASTRUCTURE ONLY -- groundwork for the next round; no CUDA path exists for it yet and nothing under
``elegantrl_b200/`` imports this module.

A numpy restatement (closed form, hand-written backward, no autograd) of

* ``ReplayBuffer.update`` / ``ReplayBuffer.sample``          reference ``elegantrl/train/replay_buffer.py:78-134``
* ``ActorSAC`` / ``CriticEnsemble``                          reference ``elegantrl/agents/AgentSAC.py:167-198, 244-259``
* ``AgentSAC.update_objectives`` (no PER, lambda_fit_cum_r = 0) reference ``elegantrl/agents/AgentSAC.py:42-86``
* ``AgentBase.optimizer_backward`` / ``soft_update``         reference ``elegantrl/agents/AgentBase.py:239-248, 270-278``

pinned by ``tests/test_sac_oracle_golden.py`` against vectors minted from the importable reference
(``oracle/make_golden.py::main_sac`` -> ``tests/golden/sac_*.npz``).  Quirks kept on purpose (they are what a drop-in
must reproduce): the log-prob is evaluated at the MEAN (``dist.log_prob(a_avg)``, so only ``-log std`` and the tanh
correction survive), the tanh correction uses ``1.000001 - tanh^2``, ``target_entropy = +log(action_dim)``, the critic's
shared encoder is a bare Linear followed directly by the decoders' first Linear, the actor step uses ``cri_target`` and the
temperature AFTER its own Adam step but BEFORE the clamp of ``alpha_log``.
"""
import math

import numpy as np

from oracle.ppo_oracle import act_fn, act_grad, adam_step, clip_grads

LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))


# ------------------------------------------------------------------------------------------- replay buffer
class ReplayBuffer:
    """Ring buffer ``[max_size, num_seqs, dim]`` over TIME (``max_size`` is the time length) -- replay_buffer.py:11-134."""

    def __init__(self, max_size, state_dim, action_dim, num_seqs=1, dtype=np.float32):
        self.p, self.if_full, self.cur_size, self.add_size = 0, False, 0, 0
        self.max_size, self.num_seqs = max_size, num_seqs
        self.states = np.zeros((max_size, num_seqs, state_dim), dtype)
        self.actions = np.zeros((max_size, num_seqs, action_dim), dtype)
        self.rewards = np.zeros((max_size, num_seqs), dtype)
        self.undones = np.zeros((max_size, num_seqs), dtype)   # float32 in the reference (:57-58)
        self.unmasks = np.zeros((max_size, num_seqs), dtype)

    def update(self, items):
        """replay_buffer.py:78-118: append ``add_size`` time rows at the pointer, wrapping around once."""
        states, actions, rewards, undones, unmasks = items
        self.add_size = rewards.shape[0]
        p = self.p + self.add_size
        fields = ((self.states, states), (self.actions, actions), (self.rewards, rewards), (self.undones, undones),
                  (self.unmasks, unmasks))
        if p > self.max_size:
            self.if_full = True
            p0, p1, p2 = self.p, self.max_size, self.max_size - self.p
            p = p - self.max_size
            for dst, src in fields:
                dst[p0:p1], dst[0:p] = src[:p2], src[-p:]
        else:
            for dst, src in fields:
                dst[self.p:p] = src
        self.p = p
        self.cur_size = self.max_size if self.if_full else self.p

    def split_ids(self, ids):
        """ids in [0, (cur_size - 1) * num_seqs) -> (time index, sequence index) -- replay_buffer.py:120-125."""
        sample_len = self.cur_size - 1
        return np.fmod(ids, sample_len), ids // sample_len

    def sample(self, ids):
        """replay_buffer.py:120-134 with the indices injected; the next state is the NEXT TIME ROW of the same sequence."""
        ids0, ids1 = self.split_ids(ids)
        return (self.states[ids0, ids1], self.actions[ids0, ids1], self.rewards[ids0, ids1], self.undones[ids0, ids1],
                self.unmasks[ids0, ids1], self.states[ids0 + 1, ids1])


# ------------------------------------------------------------------------------------------------- nets
def mlp_forward(layers, x, act_last, activation="gelu"):
    """``build_mlp`` (AgentBase.py:345-360): Linear + GELU per layer; the last activation only when if_raw_out=False.
    ``layers`` = [(W [out, in], b [out]), ...].  Returns (output, cache)."""
    inputs, pre = [], []
    h = x
    for i, (w, b) in enumerate(layers):
        inputs.append(h)
        z = (h @ w.T + b).astype(x.dtype)
        pre.append(z)
        h = act_fn(z, activation) if (i < len(layers) - 1 or act_last) else z
    return h, (inputs, pre, act_last, activation)


def mlp_backward(layers, cache, d_out):
    """Backward of ``mlp_forward``: ([(dW, db), ...], d_input)."""
    inputs, pre, act_last, activation = cache
    grads = [None] * len(layers)
    d = d_out
    for i in range(len(layers) - 1, -1, -1):
        if i < len(layers) - 1 or act_last:
            d = (d * act_grad(pre[i], activation)).astype(d.dtype)
        grads[i] = ((d.T @ inputs[i]).astype(d.dtype), d.sum(axis=0).astype(d.dtype))
        d = (d @ layers[i][0]).astype(d.dtype)
    return grads, d


def actor_forward(actor, state, eps):
    """ActorSAC.get_action_logprob (AgentSAC.py:184-196) with the rsample noise injected.
    actor = {"net_s": layers (activation on EVERY layer), "net_a": [(W [2A, d], b)]}.
    Returns (tanh(action), logprob [B], cache)."""
    dt = state.dtype
    h, cache_s = mlp_forward(actor["net_s"], state, act_last=True)
    out, cache_a = mlp_forward(actor["net_a"], h, act_last=False)
    a_dim = out.shape[1] // 2
    a_avg, lsd_raw = out[:, :a_dim], out[:, a_dim:]
    lsd = np.clip(lsd_raw, dt.type(-16), dt.type(2))
    std = np.exp(lsd)
    action = (a_avg + std * eps.astype(dt)).astype(dt)                        # Normal.rsample(): loc + eps * scale
    t = np.tanh(action)
    logprob = (-np.log(std) - dt.type(LOG_SQRT_2PI)) - np.log(-t * t + dt.type(1.000001))   # log_prob at the MEAN (:193)
    return t, logprob.sum(axis=1).astype(dt), (cache_s, cache_a, lsd_raw, std, t, eps.astype(dt))


def actor_backward(actor, cache, d_tanh, d_logprob):
    """Gradients of sum(d_tanh * tanh(action)) + sum(d_logprob * logprob) w.r.t. the actor's parameters."""
    cache_s, cache_a, lsd_raw, std, t, eps = cache
    dt = t.dtype
    one_m_t2 = dt.type(1.0) - t * t
    d_lp = d_logprob[:, None]
    d_action = d_tanh * one_m_t2 + d_lp * (dt.type(2.0) * t * one_m_t2 / (dt.type(1.000001) - t * t))
    inside = ((lsd_raw >= dt.type(-16)) & (lsd_raw <= dt.type(2))).astype(dt)  # gradient of clamp
    d_lsd = (d_action * std * eps - d_lp) * inside                            # d(-log std) / d lsd = -1
    d_out = np.concatenate((d_action, d_lsd), axis=1).astype(dt)
    g_a, d_h = mlp_backward(actor["net_a"], cache_a, d_out)
    g_s, _ = mlp_backward(actor["net_s"], cache_s, d_h)
    return g_s + g_a                                                          # parameter order: net_s then net_a (:170-171)


def critic_forward(critic, state, action):
    """CriticEnsemble.get_q_values (AgentSAC.py:256-259): shared raw Linear encoder, E decoders -> [B, E]."""
    sa = np.concatenate((state, action), axis=1)
    enc, cache_e = mlp_forward(critic["encoder"], sa, act_last=False)
    qs, caches = [], []
    for dec in critic["decoders"]:
        q, c = mlp_forward(dec, enc, act_last=False)
        qs.append(q)
        caches.append(c)
    return np.concatenate(qs, axis=1), (cache_e, caches)


def critic_backward(critic, cache, d_q):
    """d_q [B, E] -> ([(dW, db)] in parameter order encoder, decoder 0, 1, ..., d_state_action)."""
    cache_e, caches = cache
    d_enc = 0
    grads_dec = []
    for e, dec in enumerate(critic["decoders"]):
        g, d = mlp_backward(dec, caches[e], d_q[:, e:e + 1])
        grads_dec += g
        d_enc = d_enc + d
    g_enc, d_sa = mlp_backward(critic["encoder"], cache_e, d_enc.astype(d_q.dtype))
    return g_enc + grads_dec, d_sa


def critic_params(critic):
    out = [p for layer in critic["encoder"] for p in layer]
    for dec in critic["decoders"]:
        out += [p for layer in dec for p in layer]
    return out


def actor_params(actor):
    return [p for layer in actor["net_s"] + actor["net_a"] for p in layer]


def new_adam(params):
    return {"step": 0, "m": [np.zeros_like(p) for p in params], "v": [np.zeros_like(p) for p in params]}


def soft_update(target, current, tau):
    """AgentBase.soft_update (AgentBase.py:270-278): tar = cur * tau + tar * (1 - tau), parameter by parameter."""
    for tar, cur in zip(target, current):
        dt = tar.dtype
        tar[...] = cur * dt.type(tau) + tar * dt.type(1.0 - tau)


def _step(params, grads, opt, hp):
    """optimizer_backward (AgentBase.py:239-248): clip_grad_norm_ over the optimizer's parameters, then Adam."""
    grads, norm = clip_grads(grads, hp["clip_grad_norm"])
    opt["step"] = adam_step(params, grads, opt["m"], opt["v"], opt["step"], hp["learning_rate"])
    return norm


# ------------------------------------------------------------------------------------------------- update
def sac_update(agent, batch, eps_next, eps_pg, hp):
    """One ``AgentSAC.update_objectives`` (AgentSAC.py:42-86) on an already sampled batch.
    agent = dict(actor, critic, critic_target, alpha_log [1], opt_actor, opt_critic, opt_alpha); mutated in place.
    batch = (state, action, reward, undone, unmask, next_state); eps_* = the two rsample draws, in call order.
    Returns (obj_critic, obj_actor)."""
    state, action, reward, undone, unmask, next_state = batch
    dt = state.dtype
    bsz = dt.type(state.shape[0])
    actor, critic, critic_target = agent["actor"], agent["critic"], agent["critic_target"]

    # ---- q_label (no grad)                                                                   (:52-55)
    next_action, next_logprob, _ = actor_forward(actor, next_state, eps_next)
    next_q = critic_forward(critic_target, next_state, next_action)[0].min(axis=1)
    alpha = np.exp(agent["alpha_log"].astype(dt))[0]
    q_label = (reward + undone * dt.type(hp["gamma"]) * (next_q - next_logprob * alpha)).astype(dt)

    # ---- critic: mean_b(unmask * mean_e (q - label)^2), Adam, soft update of the target       (:57-69)
    q_values, cache_c = critic_forward(critic, state, action)
    n_ens = q_values.shape[1]
    err = q_values - q_label[:, None]
    td_error = (err * err).mean(axis=1) * unmask
    obj_critic = td_error.mean(dtype=dt)
    d_q = (dt.type(2.0) * err * unmask[:, None] / (bsz * dt.type(n_ens))).astype(dt)
    g_c, _ = critic_backward(critic, cache_c, d_q)
    _step(critic_params(critic), [g for pair in g_c for g in pair], agent["opt_critic"], hp)
    soft_update(critic_params(critic_target), critic_params(critic), hp["soft_update_tau"])

    # ---- temperature                                                                         (:71-74)
    action_pg, logprob, cache_a = actor_forward(actor, state, eps_pg)
    g_alpha = np.array([(dt.type(hp["target_entropy"]) - logprob).mean(dtype=dt)], dtype=dt)
    _step([agent["alpha_log"]], [g_alpha], agent["opt_alpha"], hp)

    # ---- actor: maximise mean Q_target(s, a_pg) - alpha * mean logprob                        (:76-83)
    alpha = np.exp(agent["alpha_log"].astype(dt))[0]                           # after its step, before the clamp
    agent["alpha_log"][...] = np.clip(agent["alpha_log"], dt.type(-16), dt.type(2))
    q_pg, cache_t = critic_forward(critic_target, state, action_pg)
    q_value_pg = q_pg.mean(axis=1, keepdims=True).mean(dtype=dt)
    obj_actor = (q_value_pg - logprob * alpha).mean(dtype=dt)
    d_q_pg = np.full_like(q_pg, dt.type(-1.0) / (bsz * dt.type(n_ens)))       # d(-obj_actor) / d q
    _, d_sa = critic_backward(critic_target, cache_t, d_q_pg)
    d_tanh = d_sa[:, state.shape[1]:]
    d_logprob = np.full_like(logprob, alpha / bsz)
    g_a = actor_backward(actor, cache_a, d_tanh, d_logprob)
    _step(actor_params(actor), [g for pair in g_a for g in pair], agent["opt_actor"], hp)
    return float(obj_critic), float(obj_actor)


# ------------------------------------------------------------------------------------- rollout + update_net
def explore_action(actor, state, eps):
    """ActorSAC.get_action (AgentSAC.py:176-182): tanh(mean + std * eps) -- the tanh'ed action is what the buffer stores."""
    return actor_forward(actor, state, eps)[0]


def explore_pendulum(actor, theta, theta_dot, cur_step, horizon_len, eps, reset_u, reward_scale=1.0, max_step=200):
    """AgentBase._explore_vec_env, off-policy flavour (AgentBase.py:130-170) on the Pendulum vec env: records
    (states, actions, rewards * reward_scale, undones, unmasks); no log-probs.  eps [H, N, A], reset_u [H, N, 2]."""
    from oracle.ppo_oracle import pendulum_observe, pendulum_step
    dt = theta.dtype
    n = theta.shape[0]
    states = np.zeros((horizon_len, n, 3), dt)
    actions = np.zeros((horizon_len, n, eps.shape[2]), dt)
    rewards = np.zeros((horizon_len, n), dt)
    terminals = np.zeros((horizon_len, n), bool)
    truncates = np.zeros((horizon_len, n), bool)
    for t in range(horizon_len):
        state = pendulum_observe(theta, theta_dot)
        action = explore_action(actor, state, eps[t])
        states[t], actions[t] = state, action
        theta, theta_dot, cur_step, reward, terminal, truncate = pendulum_step(theta, theta_dot, cur_step, action[:, 0], reset_u[t], max_step)
        rewards[t], terminals[t], truncates[t] = reward, terminal, truncate
    rewards = (rewards * dt.type(reward_scale)).astype(dt)
    return dict(states=states, actions=actions, rewards=rewards, undones=~terminals, unmasks=~truncates,
                last_state=pendulum_observe(theta, theta_dot), theta=theta, theta_dot=theta_dot, cur_step=cur_step)


def update_net(agent, buffer, ids_per_update, eps_next, eps_pg, hp):
    """AgentBase.update_net, off-policy (AgentBase.py:172-189): ``int(cur_size * repeat_times / batch_size)`` calls of
    update_objectives; returns the (nan)means of (obj_critic, obj_actor)."""
    logs = [sac_update(agent, buffer.sample(ids), eps_next[u], eps_pg[u], hp) for u, ids in enumerate(ids_per_update)]
    logs = np.array(logs, dtype=np.float64)
    return float(np.nanmean(logs[:, 0])), float(np.nanmean(logs[:, 1]))

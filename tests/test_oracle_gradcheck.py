"""Finite-difference check of the oracle's hand-written backward passes (float64): an independent leg under the parity
argument -- the goldens pin the oracle's OUTPUTS to the reference, this pins its GRADIENTS to its own forward.  CPU only."""
import copy

import numpy as np
import pytest

from oracle import ppo_oracle as po
from oracle import sac_oracle as so
from tests import golden_utils as gu
from tests.test_sac_oracle_golden import SAC_CASES, agent_of, build_buffer


def _losses(actor, critic, batch, hp, discrete):
    a2, c2 = copy.deepcopy(actor), copy.deepcopy(critic)
    (obj_c, obj_s, obj_e), grads = po.ppo_minibatch(a2, c2, po.new_adam_state(a2, not discrete), po.new_adam_state(c2, False),
                                                    batch, dict(hp, clip_grad_norm=0.0, learning_rate=0.0))
    sign = 1.0 if hp.get("entropy_bonus", False) else -1.0
    return obj_c, -(obj_s + sign * hp["lambda_entropy"] * obj_e), grads


FLAVOURS = [("synth_s5_a3_64x48x32", {}, False), ("synth_s8_a2_128x64", dict(po.HELLOWORLD_FLAVOUR, clip_grad_norm=0.0), False),
            ("a2c_s8_a2_64x32", po.A2C_FLAVOUR, False), ("discrete_s8_a4_128x64", dict(discrete=True), True)]


@pytest.mark.parametrize("case,flavour,discrete", FLAVOURS)
def test_ppo_minibatch_gradients(case, flavour, discrete):
    g = gu.load(case)
    hp = dict(gu.hyper_of(g), **flavour)
    actor = po.net_astype((gu.discrete_net_of if discrete else gu.net_of)(g, "actor"), np.float64)
    critic = po.net_astype(gu.net_of(g, "critic"), np.float64)
    h = g["buf.states"].shape[0]
    rng = np.random.default_rng(0)
    ids = g["update_net.ids"][0][:24] if "update_net.ids" in g else rng.integers(0, h, 24)
    buf = dict(states=g["buf.states"].astype(np.float64), actions=g["buf.actions"] if discrete else g["buf.actions"].astype(np.float64),
               unmasks=g["buf.unmasks"], logprobs=g["buf.logprobs"].astype(np.float64),
               advantages=g["gae.adv_norm"].astype(np.float64), reward_sums=g["gae.reward_sums"].astype(np.float64))
    batch = po.gather_minibatch(buf, ids)
    _, _, grads = _losses(actor, critic, batch, hp, discrete)
    eps = 1e-6
    for which, net, analytic in (("actor", actor, grads["actor"]), ("critic", critic, grads["critic"])):
        params = gu.flat_params(net)
        assert len(params) == len(analytic)
        for p, ga in zip(params, analytic):
            for _ in range(3):  # three random elements per tensor
                idx = tuple(rng.integers(0, s) for s in p.shape)
                old = p[idx]
                p[idx] = old + eps
                up = _losses(actor, critic, batch, hp, discrete)
                p[idx] = old - eps
                dn = _losses(actor, critic, batch, hp, discrete)
                p[idx] = old
                k = 0 if which == "critic" else 1
                numeric = (up[k] - dn[k]) / (2 * eps)
                assert abs(numeric - ga[idx]) <= 1e-6 + 1e-5 * abs(numeric), (which, p.shape, idx, numeric, ga[idx])


@pytest.mark.parametrize("case", SAC_CASES)
def test_sac_gradients(case):
    """Actor objective of SAC: d(-(mean Q_target(s, tanh(a)) - alpha * mean logprob)) / d actor parameters, through the target
    critic's input gradient and the tanh / clamp / log-std chain."""
    g = gu.load(case)
    agent = agent_of(g, "init")
    for key in ("actor", "critic_target"):
        agent[key] = copy.deepcopy(agent[key])
    to64 = lambda layers: [(w.astype(np.float64), b.astype(np.float64)) for w, b in layers]
    actor = {k: to64(v) for k, v in agent["actor"].items()}
    target = {"encoder": to64(agent["critic_target"]["encoder"]), "decoders": [to64(d) for d in agent["critic_target"]["decoders"]]}
    buf = build_buffer(g)
    state = buf.sample(g["update.ids"][0])[0].astype(np.float64)[:16]
    eps_pg = g["update.eps_pg"][0].astype(np.float64)[:16]
    alpha = 0.37

    def loss():
        t, logprob, cache = so.actor_forward(actor, state, eps_pg)
        q, cache_t = so.critic_forward(target, state, t)
        return -(q.mean(axis=1).mean() - alpha * logprob.mean()), (t, logprob, cache, q, cache_t)

    _, (t, logprob, cache, q, cache_t) = loss()
    bsz = state.shape[0]
    _, d_sa = so.critic_backward(target, cache_t, np.full_like(q, -1.0 / (bsz * q.shape[1])))
    grads = so.actor_backward(actor, cache, d_sa[:, state.shape[1]:], np.full_like(logprob, alpha / bsz))
    analytic = [x for pair in grads for x in pair]
    rng = np.random.default_rng(1)
    step = 1e-6
    for p, ga in zip(so.actor_params(actor), analytic):
        for _ in range(4):
            idx = tuple(rng.integers(0, s) for s in p.shape)
            old = p[idx]
            p[idx] = old + step
            up = loss()[0]
            p[idx] = old - step
            dn = loss()[0]
            p[idx] = old
            numeric = (up - dn) / (2 * step)
            assert abs(numeric - ga[idx]) <= 1e-7 + 1e-5 * abs(numeric), (p.shape, idx, numeric, ga[idx])

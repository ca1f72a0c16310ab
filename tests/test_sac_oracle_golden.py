"""Pin the off-policy oracle (oracle/sac_oracle.py: ReplayBuffer + SAC) to vectors minted from the Python
reference (oracle/make_golden.py::main_sac).  CPU only; groundwork for the next round -- no CUDA path exists yet."""
import glob
import os

import numpy as np
import pytest

from oracle import sac_oracle as so
from tests import golden_utils as gu

SAC_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(gu.GOLDEN_DIR, "sac_*.npz")))


def layers_of(g, prefix):
    n = int(g[f"{prefix}.n_layers"])
    return [(g[f"{prefix}.W{i}"].copy(), g[f"{prefix}.b{i}"].copy()) for i in range(n)]


def agent_of(g, prefix):
    n_ens = int(g["hp.num_ensembles"])
    critic = lambda name: {"encoder": layers_of(g, f"{prefix}.{name}.encoder"),
                           "decoders": [layers_of(g, f"{prefix}.{name}.decoder{e}") for e in range(n_ens)]}
    agent = {"actor": {"net_s": layers_of(g, f"{prefix}.actor.net_s"), "net_a": layers_of(g, f"{prefix}.actor.net_a")},
             "critic": critic("critic"), "critic_target": critic("critic_target"), "alpha_log": g[f"{prefix}.alpha_log"].copy()}
    agent["opt_actor"] = so.new_adam(so.actor_params(agent["actor"]))
    agent["opt_critic"] = so.new_adam(so.critic_params(agent["critic"]))
    agent["opt_alpha"] = so.new_adam([agent["alpha_log"]])
    return agent


def flat(agent):
    return (so.actor_params(agent["actor"]) + so.critic_params(agent["critic"]) + so.critic_params(agent["critic_target"])
            + [agent["alpha_log"]])


def build_buffer(g):
    s_dim, a_dim, num_seqs, max_size = (int(x) for x in g["dims"][:4])
    buf = so.ReplayBuffer(max_size, s_dim, a_dim, num_seqs)
    for i in range(3):
        buf.update(tuple(g[f"append{i}.{k}"] for k in ("states", "actions", "rewards", "undones", "unmasks")))
        assert buf.p == int(g[f"append{i}.p"]) and buf.cur_size == int(g[f"append{i}.cur_size"])  # pointer arithmetic: exact
    return buf


@pytest.mark.parametrize("case", SAC_CASES)
def test_replay_buffer_ring(case):
    """ReplayBuffer.update incl. the wrap-around branch and the (time, sequence) index split: bit-exact."""
    g = gu.load(case)
    buf = build_buffer(g)
    assert buf.if_full
    for k in ("states", "actions", "rewards", "undones", "unmasks"):
        assert np.array_equal(getattr(buf, k), g[f"buffer.{k}"]), k
    ids = g["update.ids"][0]
    ids0, ids1 = buf.split_ids(ids)
    assert ids0.max() < buf.cur_size - 1 and ids1.max() < buf.num_seqs
    state, action, reward, undone, unmask, next_state = buf.sample(ids)
    assert np.array_equal(next_state, g["buffer.states"][ids0 + 1, ids1]) and np.array_equal(reward, g["buffer.rewards"][ids0, ids1])


@pytest.mark.parametrize("case", SAC_CASES)
def test_sac_update_objectives(case):
    """AgentSAC.update_objectives x3: obj_critic / obj_actor per call, every parameter (actor, critic ensemble, target,
    alpha_log) after the first and after the last call."""
    g = gu.load(case)
    buf = build_buffer(g)
    agent = agent_of(g, "init")
    hp = {k: float(g[f"hp.{k}"]) for k in ("gamma", "clip_grad_norm", "learning_rate", "soft_update_tau", "target_entropy")}
    for u, ids in enumerate(g["update.ids"]):
        scalars = so.sac_update(agent, buf.sample(ids), g["update.eps_next"][u], g["update.eps_pg"][u], hp)
        np.testing.assert_allclose(scalars, g["update.scalars"][u], rtol=1e-4, atol=2e-6)
        if u == 0:
            for mine, ref in zip(flat(agent), flat(agent_of(g, "after1"))):
                np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=2e-6)
    for mine, ref in zip(flat(agent), flat(agent_of(g, "after"))):
        np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=3e-6)


SACCYCLE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(gu.GOLDEN_DIR, "saccycle_*.npz")))


@pytest.mark.parametrize("case", SACCYCLE_CASES)
def test_sac_full_cycle(case):
    """explore_env (off-policy, tanh'ed actions stored) on the Pendulum vec env -> ReplayBuffer.update -> update_net."""
    g = gu.load(case)
    agent = agent_of(g, "init")
    h, max_step, n, max_size = int(g["horizon_len"]), int(g["max_step"]), int(g["dims"][2]), int(g["dims"][3])
    hp = {k: float(g[f"hp.{k}"]) for k in ("gamma", "clip_grad_norm", "learning_rate", "soft_update_tau", "target_entropy")}
    roll = so.explore_pendulum(agent["actor"], g["env.theta0"], g["env.theta_dot0"], g["env.cur_step0"], h, g["explore.eps"],
                               g["env.reset_noise"], float(g["hp.reward_scale"]), max_step)
    for k in ("states", "actions", "rewards"):
        np.testing.assert_allclose(roll[k], g[f"explore.{k}"], rtol=1e-4, atol=1e-5, err_msg=k)
    assert np.array_equal(roll["undones"], g["explore.undones"]) and np.array_equal(roll["unmasks"], g["explore.unmasks"])
    assert not g["explore.unmasks"].all()
    np.testing.assert_allclose(roll["last_state"], g["explore.last_state"], rtol=1e-4, atol=1e-5)
    buf = so.ReplayBuffer(max_size, 3, 1, n)
    buf.update(tuple(g[f"explore.{k}"] for k in ("states", "actions", "rewards", "undones", "unmasks")))
    assert buf.cur_size == h and not buf.if_full
    assert len(g["update.ids"]) == int(buf.cur_size * float(g["hp.repeat_times"]) / int(g["hp.batch_size"]))
    result = so.update_net(agent, buf, g["update.ids"], g["update.eps_next"], g["update.eps_pg"], hp)
    np.testing.assert_allclose(result, g["update_net.result"], rtol=1e-4, atol=2e-6)
    for mine, ref in zip(flat(agent), flat(agent_of(g, "after"))):
        np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=3e-6)

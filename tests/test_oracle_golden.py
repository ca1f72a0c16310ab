"""Pin the numpy oracle (oracle/ppo_oracle.py) to vectors minted from the Python reference itself.

CPU only.  Tolerances: float paths rtol 1e-5 / small atol (numpy vs ATen differ in summation order and erf/exp
rounding); index and mask paths bit-exact.
"""
import numpy as np
import pytest

from oracle import ppo_oracle as po
from tests import golden_utils as gu

F32 = dict(rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("case", gu.SYNTH_CASES)
def test_nets(case):
    g = gu.load(case)
    actor, critic = gu.net_of(g, "actor"), gu.net_of(g, "critic")
    state, action = g["nets.state"], g["nets.action"]
    np.testing.assert_allclose(po.actor_mean(actor, state), g["nets.actor_mean"], **F32)
    np.testing.assert_allclose(po.actor_forward(actor, state), g["nets.actor_forward"], **F32)
    logprob, entropy = po.logprob_entropy(actor, state, action)
    np.testing.assert_allclose(logprob, g["nets.logprob"], **F32)
    np.testing.assert_allclose(entropy, g["nets.entropy"], **F32)
    np.testing.assert_allclose(po.critic_value(critic, state), g["nets.value"], **F32)


@pytest.mark.parametrize("case", gu.SYNTH_CASES + gu.ROLLOUT_CASES)
@pytest.mark.parametrize("tag", ["gae", "gae_alt"])
def test_gae(case, tag):
    g = gu.load(case)
    if f"{tag}.values" not in g:
        pytest.skip("branch not recorded for this case")
    hp = gu.hyper_of(g)
    v_trace = hp["if_use_v_trace"] if tag == "gae" else not hp["if_use_v_trace"]
    critic = gu.net_of(g, "critic")
    src = "buf" if "buf.states" in g else "rollout"
    rewards, undones = g[f"{src}.rewards"].copy(), g[f"{src}.undones"].copy()
    values, adv, rsum, adv_norm = po.values_gae_pass(
        critic, g[f"{src}.states"], rewards, undones, g[f"{src}.unmasks"], g[f"{src}.last_state"],
        hp["gamma"], hp["lambda_gae_adv"], v_trace)
    np.testing.assert_allclose(values, g[f"{tag}.values"], **F32)
    np.testing.assert_allclose(adv, g[f"{tag}.advantages"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(rsum, g[f"{tag}.reward_sums"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(adv_norm, g[f"{tag}.adv_norm"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(rewards, g[f"{tag}.rewards_after"], **F32)
    assert np.array_equal(undones, g[f"{tag}.undones_after"])  # mask path: bit-exact
    mean, std = po.advantage_stats(g[f"{tag}.advantages"])
    np.testing.assert_allclose(mean, g[f"{tag}.adv_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(std, g[f"{tag}.adv_std"], rtol=1e-5)


def test_gae_scan_exact_given_reference_values():
    """With the reference's own values the scan itself is bit-exact (same op order)."""
    for case in gu.SYNTH_CASES:
        g = gu.load(case)
        hp = gu.hyper_of(g)
        rewards, undones = g["buf.rewards"].copy(), g["buf.undones"].copy()
        adv = po.gae(rewards, undones, g["buf.unmasks"], g["gae.values"], g["gae.last_value"],
                     hp["gamma"], hp["lambda_gae_adv"], hp["if_use_v_trace"])
        assert np.array_equal(adv, g["gae.advantages"]), case
        assert np.array_equal(rewards, g["gae.rewards_after"]) and np.array_equal(undones, g["gae.undones_after"])


@pytest.mark.parametrize("case", gu.SYNTH_CASES)
def test_index_split_exact(case):
    g = gu.load(case)
    horizon_len = g["buf.states"].shape[0]
    ids0, ids1 = po.split_ids(g["update.ids"], horizon_len)
    assert np.array_equal(ids0, g["update.ids0"]) and np.array_equal(ids1, g["update.ids1"])


@pytest.mark.parametrize("case", gu.SYNTH_CASES)
def test_update_objectives(case):
    g = gu.load(case)
    hp = gu.hyper_of(g)
    actor, critic = gu.net_of(g, "actor"), gu.net_of(g, "critic")
    opt_a, opt_c = po.new_adam_state(actor, True), po.new_adam_state(critic, False)
    buffer = dict(states=g["buf.states"], actions=g["buf.actions"], unmasks=g["buf.unmasks"],
                  logprobs=g["buf.logprobs"], advantages=g["gae.adv_norm"], reward_sums=g["gae.reward_sums"])
    for u, ids in enumerate(g["update.ids"]):
        scalars, _ = po.ppo_minibatch(actor, critic, opt_a, opt_c, po.gather_minibatch(buffer, ids), hp)
        np.testing.assert_allclose(scalars, g["update.scalars"][u], rtol=1e-4, atol=1e-6)
        if u == 0:
            for mine, ref in zip(gu.flat_params(actor), gu.flat_params(gu.net_of(g, "update.after1.actor"))):
                np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=1e-6)
    for prefix, net, opt in (("actor", actor, opt_a), ("critic", critic, opt_c)):
        for mine, ref in zip(gu.flat_params(net), gu.flat_params(gu.net_of(g, f"update.after.{prefix}"))):
            np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=2e-6)
        assert opt["step"] == int(g[f"update.after.{prefix}_adam.step"])
        for i in range(len(net["W"])):
            m_ref, v_ref = g[f"update.after.{prefix}_adam.m.W{i}"], g[f"update.after.{prefix}_adam.v.W{i}"]
            np.testing.assert_allclose(opt["m_W"][i], m_ref, rtol=1e-4, atol=1e-5 * np.abs(m_ref).max())
            np.testing.assert_allclose(opt["v_W"][i], v_ref, rtol=2e-4, atol=1e-5 * np.abs(v_ref).max())


@pytest.mark.parametrize("case", gu.SYNTH_CASES + gu.ROLLOUT_CASES)
def test_update_net(case):
    g = gu.load(case)
    hp = gu.hyper_of(g)
    actor, critic = gu.net_of(g, "actor"), gu.net_of(g, "critic")
    opt_a, opt_c = po.new_adam_state(actor, True), po.new_adam_state(critic, False)
    src = "buf" if "buf.states" in g else "rollout"
    rollout = {k: g[f"{src}.{k}"].copy() for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")}
    result, _ = po.update_net(actor, critic, opt_a, opt_c, rollout, g[f"{src}.last_state"], g["update_net.ids"], hp)
    np.testing.assert_allclose(result, g["update_net.result"], rtol=1e-4, atol=1e-6)
    for prefix, net in (("actor", actor), ("critic", critic)):
        for mine, ref in zip(gu.flat_params(net), gu.flat_params(gu.net_of(g, f"update_net.after.{prefix}"))):
            np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("case", gu.ROLLOUT_CASES)
def test_rollout(case):
    g = gu.load(case)
    hp = gu.hyper_of(g)
    actor, critic = gu.net_of(g, "actor"), gu.net_of(g, "critic")
    horizon_len = g["rollout.states"].shape[0]
    out = po.rollout_pendulum(actor, critic, g["env.theta0"], g["env.theta_dot0"], g["env.cur_step0"],
                              horizon_len, g["eps"], g["env.reset_noise"], hp["reward_scale"], int(g["max_step"]))
    for k in ("states", "actions", "logprobs", "rewards"):
        np.testing.assert_allclose(out[k], g[f"rollout.{k}"], rtol=1e-4, atol=1e-5, err_msg=k)
    assert np.array_equal(out["undones"], g["rollout.undones"])  # masks: bit-exact
    assert np.array_equal(out["unmasks"], g["rollout.unmasks"])
    assert np.array_equal(out["cur_step"], g["rollout.cur_step"])
    np.testing.assert_allclose(out["last_state"], g["rollout.last_state"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["values"], g["gae.values"], rtol=1e-4, atol=1e-5)


def test_one_env_rollout():
    """Row a5 golden (reference ``_explore_one_env``, AgentPPO.py:34-85): the oracle's policy on the golden's own states
    and replayed noise reproduces its actions / log-probs; the [H, 1, ...] buffer goes through the oracle's update_net."""
    g = gu.load("oneenv_pendulum_h48")
    hp = gu.hyper_of(g)
    actor, critic = gu.net_of(g, "actor"), gu.net_of(g, "critic")
    states = g["rollout.states"]
    assert states.shape[1] == 1 and g["rollout.undones"].dtype == np.bool_
    action, logprob = po.sample_action(actor, states[:, 0], g["eps"][:, 0])
    np.testing.assert_allclose(action, g["rollout.actions"][:, 0], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(logprob, g["rollout.logprobs"][:, 0], rtol=1e-4, atol=1e-5)
    opt_a, opt_c = po.new_adam_state(actor, True), po.new_adam_state(critic, False)
    rollout = {k: g[f"rollout.{k}"].copy() for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")}
    result, _ = po.update_net(actor, critic, opt_a, opt_c, rollout, g["rollout.last_state"], g["update_net.ids"], hp)
    np.testing.assert_allclose(result, g["update_net.result"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("case", gu.HELLOWORLD_CASES)
def test_helloworld_update_net(case):
    """The helloworld flavour of the oracle against helloworld_PPO_single_file.AgentPPO.update_net."""
    g = gu.load(case)
    hp = gu.helloworld_hyper_of(g)
    actor, critic = gu.plain_net_of(g, "actor"), gu.plain_net_of(g, "critic")
    opt_a, opt_c = po.new_adam_state(actor, True), po.new_adam_state(critic, False)
    buf = [g[f"buf.{k}"].copy() for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")]
    result, info = po.update_net_helloworld(actor, critic, opt_a, opt_c, buf, g["buf.last_state"], g["update_net.ids"], hp)
    np.testing.assert_allclose(info["values"], g["values"], **F32)
    np.testing.assert_allclose(result, g["update_net.result"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(buf[3], g["update_net.rewards_after"], **F32)
    assert np.array_equal(buf[4], g["update_net.undones_after"])
    for prefix, net in (("actor", actor), ("critic", critic)):
        for mine, ref in zip(gu.flat_params(net), gu.flat_params(gu.plain_net_of(g, f"update_net.after.{prefix}"))):
            np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=2e-6)


# ---------------------------------------------------------------------------------- discrete PPO
@pytest.mark.parametrize("case", gu.DISCRETE_CASES)
def test_discrete_nets_and_sampling(case):
    """ActorDiscretePPO (reference AgentPPO.py:393-425): logits, greedy action, log-prob / entropy of Categorical, and
    get_action with torch.multinomial's Exp(1) noise replayed -- sampled indices bit-exact."""
    g = gu.load(case)
    actor, critic = gu.discrete_net_of(g, "actor"), gu.net_of(g, "critic")
    state = g["nets.state"]
    np.testing.assert_allclose(po.actor_mean(actor, state), g["nets.logits"], **F32)
    assert np.array_equal(np.argmax(po.actor_mean(actor, state), axis=1), g["nets.actor_forward"])
    logprob, entropy = po.logprob_entropy_discrete(actor, state, g["nets.action"])
    np.testing.assert_allclose(logprob, g["nets.logprob"], **F32)
    np.testing.assert_allclose(entropy, g["nets.entropy"], **F32)
    np.testing.assert_allclose(po.critic_value(critic, state), g["nets.value"], **F32)
    action, sampled_logprob = po.sample_action_discrete(actor, state, g["sample.expo"])
    assert np.array_equal(action, g["sample.action"])
    np.testing.assert_allclose(sampled_logprob, g["sample.logprob"], **F32)


@pytest.mark.parametrize("case", gu.DISCRETE_CASES)
def test_discrete_update_objectives(case):
    g = gu.load(case)
    hp = dict(gu.hyper_of(g), discrete=True)
    actor, critic = gu.discrete_net_of(g, "actor"), gu.net_of(g, "critic")
    opt_a, opt_c = po.new_adam_state(actor, False), po.new_adam_state(critic, False)
    buffer = dict(states=g["buf.states"], actions=g["buf.actions"], unmasks=g["buf.unmasks"],
                  logprobs=g["buf.logprobs"], advantages=g["gae.adv_norm"], reward_sums=g["gae.reward_sums"])
    for u, ids in enumerate(g["update.ids"]):
        scalars, _ = po.ppo_minibatch(actor, critic, opt_a, opt_c, po.gather_minibatch(buffer, ids), hp)
        np.testing.assert_allclose(scalars, g["update.scalars"][u], rtol=1e-4, atol=1e-6)
    for prefix, net, opt in (("actor", actor, opt_a), ("critic", critic, opt_c)):
        for mine, ref in zip(gu.flat_params(net), gu.flat_params(gu.discrete_net_of(g, f"update.after.{prefix}"))):
            np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=2e-6)
        for i in range(len(net["W"])):
            np.testing.assert_allclose(opt["m_W"][i], g[f"update.after.{prefix}_adam.m.W{i}"], rtol=1e-3, atol=1e-7)
    # the inherited action_std_log never moves (no gradient reaches it)
    assert np.array_equal(g["update.after.actor.action_std_log"], g["actor.action_std_log"])


@pytest.mark.parametrize("case", gu.DISCRETE_CASES + gu.CARTPOLE_CASES)
def test_discrete_update_net(case):
    g = gu.load(case)
    hp = dict(gu.hyper_of(g), discrete=True)
    actor, critic = gu.discrete_net_of(g, "actor"), gu.net_of(g, "critic")
    opt_a, opt_c = po.new_adam_state(actor, False), po.new_adam_state(critic, False)
    src = "buf" if "buf.states" in g else "rollout"
    rollout = {k: g[f"{src}.{k}"].copy() for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")}
    result, _ = po.update_net(actor, critic, opt_a, opt_c, rollout, g[f"{src}.last_state"], g["update_net.ids"], hp)
    np.testing.assert_allclose(result, g["update_net.result"], rtol=1e-4, atol=1e-6)
    for prefix, net in (("actor", actor), ("critic", critic)):
        for mine, ref in zip(gu.flat_params(net), gu.flat_params(gu.discrete_net_of(g, f"update_net.after.{prefix}"))):
            np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("case", gu.CARTPOLE_CASES)
def test_discrete_rollout_cartpole(case):
    """Reference _explore_vec_env (discrete branch) on the torch CartPole vec env: integer actions and both masks
    bit-exact (real terminals AND truncations occur), floats to 1e-4."""
    g = gu.load(case)
    hp = gu.hyper_of(g)
    actor, critic = gu.discrete_net_of(g, "actor"), gu.net_of(g, "critic")
    horizon_len = g["rollout.states"].shape[0]
    out = po.rollout_cartpole(actor, critic, g["env.state0"], g["env.cur_step0"], horizon_len, g["expo"],
                              g["env.reset_noise"], hp["reward_scale"], int(g["max_step"]))
    assert np.array_equal(out["actions"], g["rollout.actions"]) and out["actions"].dtype == g["rollout.actions"].dtype
    assert np.array_equal(out["undones"], g["rollout.undones"]) and np.array_equal(out["unmasks"], g["rollout.unmasks"])
    assert not g["rollout.undones"].all() and not g["rollout.unmasks"].all()
    assert np.array_equal(out["cur_step"], g["rollout.cur_step"])
    for k in ("states", "logprobs", "rewards"):
        np.testing.assert_allclose(out[k], g[f"rollout.{k}"], rtol=1e-4, atol=1e-5, err_msg=k)
    np.testing.assert_allclose(out["last_state"], g["rollout.last_state"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["values"], g["gae.values"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("case", gu.A2C_CASES)
def test_a2c_update_net(case):
    """The A2C flavour of the oracle against elegantrl.agents.AgentA2C.update_net on single-env buffers [H, 1, ...]."""
    g = gu.load(case)
    hp = gu.hyper_of(g)
    actor, critic = gu.net_of(g, "actor"), gu.net_of(g, "critic")
    opt_a, opt_c = po.new_adam_state(actor, True), po.new_adam_state(critic, False)
    rollout = {k: g[f"buf.{k}"].copy() for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")}
    result, _ = po.update_net_a2c(actor, critic, opt_a, opt_c, rollout, g["buf.last_state"], g["update_net.ids"], hp)
    np.testing.assert_allclose(result, g["update_net.result"], rtol=1e-4, atol=1e-6)
    assert result[2] == 0.0 and g["update_net.result"][2] == 0.0
    for prefix, net in (("actor", actor), ("critic", critic)):
        for mine, ref in zip(gu.flat_params(net), gu.flat_params(gu.net_of(g, f"update_net.after.{prefix}"))):
            np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=2e-6)

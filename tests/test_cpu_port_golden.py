"""Pin the PyTorch-CPU port that bench.py times as the CPU baseline (oracle/cpu_port.py) to the goldens minted
from the real reference: same RNG replay -> same numbers (it issues the same ATen ops)."""
import numpy as np
import pytest
import torch as th

from oracle.cpu_port import CpuPPO
from tests import golden_utils as gu
from tests.gpu_utils import load_module

CASE_SEEDS = {"synth_s3_a1_64x64": 11, "synth_s8_a2_128x64": 23, "synth_s5_a3_64x48x32": 37, "synth_s3_a1_n1": 41,
              "rollout_pendulum_n32_h40": 53, "rollout_pendulum_n8_h16": 61, "rollout_pendulum_n1344_h40": 71}


def port_from_golden(g):
    dims = [int(x) for x in g["dims"]]
    hp = gu.hyper_of(g)
    hp.pop("reward_scale_unused", None)
    port = CpuPPO(dims[4:], dims[0], dims[1], dims[2], **hp)
    load_module(port.act, gu.net_of(g, "actor"))
    load_module(port.cri, gu.net_of(g, "critic"))
    return port


@pytest.mark.parametrize("case", sorted(CASE_SEEDS))
def test_update_net_replays_reference(case):
    g = gu.load(case)
    port = port_from_golden(g)
    src = "buf" if "buf.states" in g else "rollout"
    buffer = [th.from_numpy(g[f"{src}.{k}"].copy()) for k in ("states", "actions", "logprobs", "rewards", "undones", "unmasks")]
    port.last_state = th.from_numpy(g[f"{src}.last_state"].copy())
    th.manual_seed(CASE_SEEDS[case] + 5)
    result = port.update_net(buffer)
    np.testing.assert_allclose(result, g["update_net.result"], rtol=1e-6, atol=1e-7)
    ref = gu.net_of(g, "update_net.after.actor")
    for layer, w in zip([m for m in port.act.net if hasattr(m, "weight")], ref["W"]):
        np.testing.assert_allclose(layer.weight.detach().numpy(), w, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("case", gu.ROLLOUT_CASES)
def test_rollout_replays_reference(case):
    from elegantrl_b200.envs import PendulumVecEnv
    g = gu.load(case)
    port = port_from_golden(g)
    n, h = int(g["dims"][2]), int(g["dims"][3])
    env = PendulumVecEnv(num_envs=n, gpu_id=-1, max_step=int(g["max_step"]))
    env.theta, env.theta_dot = th.from_numpy(g["env.theta0"].copy()), th.from_numpy(g["env.theta_dot0"].copy())
    env.cur_step = th.from_numpy(g["env.cur_step0"].copy())
    env.inject_reset_noise(th.from_numpy(g["env.reset_noise"]))
    port.last_state = th.from_numpy(g["state0"].copy())
    th.manual_seed(CASE_SEEDS[case] + 3)
    out = port.explore_env(env, h)
    for name, got in zip(("states", "actions", "logprobs", "rewards", "undones", "unmasks"), out):
        assert np.array_equal(got.numpy(), g[f"rollout.{name}"]), name

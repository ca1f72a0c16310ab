"""Parity tests of kernels under ``csrc/experimental/`` (written, compile-checked, NOT yet validated on a GPU; built only with
``B200RL_BUILD_EXPERIMENTAL=1``).  With the default build every test here is skipped."""
import numpy as np
import pytest
import torch as th

from elegantrl_b200 import _build
from tests import golden_utils as gu

pytestmark = pytest.mark.gpu


def _has_ws():
    """True when libb200rl.so contains the experimental kernel (its name appears in the ELF / fatbin symbol tables)."""
    try:
        with open(_build.LIB_PATH, "rb") as f:
            return b"rollout_pendulum_tc_ws_kernel" in f.read()
    except OSError:
        return False


needs_ws = pytest.mark.skipif(not _has_ws(), reason="libb200rl.so was built without csrc/experimental (B200RL_BUILD_EXPERIMENTAL=1)")


@needs_ws
@pytest.mark.parametrize("case", gu.ROLLOUT_CASES)
def test_ws_rollout_against_reference(case, monkeypatch):
    """The warp-specialised rollout (dedicated MMA-issuer warps) against the same goldens as the production kernel."""
    from tests import gpu_utils as G
    from tests.test_gpu_parity import test_fused_rollout_against_reference
    monkeypatch.setenv("B200RL_ROLLOUT", "ws")
    test_fused_rollout_against_reference(case, "ws")


@needs_ws
def test_ws_rollout_equals_production_kernel_at_full_size(monkeypatch):
    from elegantrl_b200 import Config
    from elegantrl_b200.agents import AgentPPO
    from elegantrl_b200.envs import PendulumVecEnv
    outs = []
    for mode in ("tc", "ws"):
        monkeypatch.setenv("B200RL_ROLLOUT", mode)
        n, h = 65536, 64
        args = Config(AgentPPO, None, {'env_name': 'x', 'num_envs': n, 'max_step': 200, 'state_dim': 3, 'action_dim': 1, 'if_discrete': False})
        args.net_dims, args.random_seed = [64, 64], 3
        th.manual_seed(0)
        agent = AgentPPO(args.net_dims, 3, 1, gpu_id=0, args=args)
        env = PendulumVecEnv(num_envs=n, gpu_id=0, max_step=37, seed=5)
        agent.last_state = env.reset()[0]
        outs.append([t.clone() for t in agent.explore_env(env, h)] + [agent._value_cache[1].clone(), agent.last_state.clone()])
    for a, b in zip(*outs):
        if a.dtype == th.bool:
            assert th.equal(a, b)
        else:
            np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-5, atol=1e-6)

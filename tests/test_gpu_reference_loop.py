"""The drop-in agent under the REFERENCE's own training loop (north_star: "drops in behind examples/demo_A2C_PPO.py").

``oracle/_ref`` holds a byte-for-byte copy of the reference package (``oracle/make_ref.py``; git-ignored, travels to the
GPU box).  This test hands ``elegantrl_b200.agents.AgentPPO`` + the tensor Pendulum vec env to the reference's
``train_agent`` (``elegantrl/train/run.py:28-138``: build_env, agent construction, ``explore_env`` / ``update_net`` loop,
``agent.explore_rate``, ``Evaluator.evaluate_and_save`` calling ``actor(state)`` on the eval env, ``th.save(actor)``,
``save_or_load_agent``) exactly as ``examples/demo_A2C_PPO.py:57-88`` configures it, for a few cycles on the B200, and checks
what the loop leaves behind.  matplotlib is absent from the image: ``draw_learning_curve`` (``evaluator.py:238-300``) gets a
stub module."""
import os
import sys
import types
from unittest import mock

import numpy as np
import pytest
import torch as th

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "oracle", "_ref")


@pytest.fixture
def reference_package(monkeypatch):
    if not os.path.isdir(os.path.join(REF, "elegantrl")):
        pytest.skip("oracle/_ref is absent: run `python oracle/make_ref.py` in the build container")
    monkeypatch.syspath_prepend(REF)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = mock.MagicMock(name="matplotlib.pyplot")
        plt.subplots.side_effect = lambda *a, **k: (mock.MagicMock(), [mock.MagicMock(), mock.MagicMock()])
        mpl.pyplot = plt
        monkeypatch.setitem(sys.modules, "matplotlib", mpl)
        monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    yield
    for name in [m for m in sys.modules if m == "elegantrl" or m.startswith("elegantrl.")]:
        sys.modules.pop(name, None)


@pytest.mark.parametrize("num_envs", [512, 1])
def test_reference_train_agent_runs_the_dropin(reference_package, tmp_path, num_envs):
    from elegantrl.train.config import Config          # the reference's own Config / loop
    from elegantrl.train.run import train_agent
    from elegantrl_b200.agents import AgentPPO
    from elegantrl_b200.envs import PendulumEnv, PendulumVecEnv

    th.cuda.set_device(0)
    horizon_len, cycles = 32, 4
    env_class = PendulumVecEnv if num_envs > 1 else PendulumEnv
    env_args = {'env_name': 'Pendulum-v1', 'num_envs': num_envs, 'max_step': 40, 'state_dim': 3, 'action_dim': 1,
                'if_discrete': False}
    args = Config(agent_class=AgentPPO, env_class=env_class, env_args=env_args)   # demo_A2C_PPO.py:73-88
    args.net_dims = [64, 64]
    args.horizon_len = horizon_len
    args.batch_size, args.repeat_times = (128, 8) if num_envs > 1 else (16, 2)
    args.gamma, args.state_value_tau = 0.97, 0.1
    args.break_step = horizon_len * (cycles - 1)
    args.eval_per_step, args.eval_times, args.save_gap = horizon_len, 2, 1
    args.gpu_id, args.num_workers, args.num_threads = 0, 1, 2
    args.cwd, args.if_remove = str(tmp_path / "run"), True
    try:
        train_agent(args, if_single_process=True)
    finally:
        th.set_grad_enabled(True)     # the loop switches autograd off globally (run.py:41)

    cwd = args.cwd
    recorder = np.load(os.path.join(cwd, "recorder.npy"))
    assert recorder.ndim == 2 and recorder.shape[0] >= cycles - 1 and recorder.shape[1] >= 7   # step, avgR, stdR, expR, objC, objA, ...
    assert np.isfinite(recorder[:, :6]).all()
    assert list(recorder[:, 0]) == sorted(recorder[:, 0]) and recorder[-1, 0] >= args.break_step
    assert any(f.startswith("actor__") and f.endswith(".pt") for f in os.listdir(cwd))   # th.save(actor) by the Evaluator
    for name in ("act.pth", "cri.pth", "act_optimizer.pth", "cri_optimizer.pth"):             # save_or_load_agent (AgentBase.py:280-297)
        assert os.path.isfile(os.path.join(cwd, name)), name
    actor = th.load(os.path.join(cwd, "act.pth"), map_location="cuda:0", weights_only=False)
    out = actor(th.zeros((5, 3), device="cuda:0"))
    assert out.shape == (5, 1) and bool((out.abs() <= 1).all())

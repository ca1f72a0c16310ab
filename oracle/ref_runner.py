"""Run the UNMODIFIED reference (``oracle/_ref/elegantrl``, placed there by ``oracle/make_ref.py``) on the host cores.
TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by the product package.

The reference's own ``AgentPPO`` (``elegantrl/agents/AgentPPO.py:12-232``) is built with ``gpu_id=-1`` and driven through
its own public API -- ``explore_env(env, horizon_len)`` then ``update_net(buffer)`` -- on the torch Pendulum vec env of
``elegantrl_b200/envs/pendulum.py`` (gymnasium's physics is not part of the reference tree; this env is the one the
goldens were minted on).  ``bench.py --impl reference`` and the ``cpu_baseline`` leg time exactly this; when ``oracle/_ref``
is absent they fall back to the PyTorch-CPU port (``oracle/cpu_port.py``) and say so (``kind: "port"``).
"""
import importlib
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "elegantrl", "agents", "AgentPPO.py"))


def load():
    """Import the reference package from oracle/_ref (and nowhere else).  Returns the ``elegantrl`` module."""
    if not available():
        raise ImportError("oracle/_ref/elegantrl is missing: run `python oracle/make_ref.py` where /root/reference exists")
    mod = sys.modules.get("elegantrl")
    if mod is not None and os.path.abspath(os.path.dirname(mod.__file__)) != os.path.join(REF_ROOT, "elegantrl"):
        raise ImportError(f"another `elegantrl` is already imported from {mod.__file__}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    mod = importlib.import_module("elegantrl")
    assert os.path.abspath(os.path.dirname(mod.__file__)) == os.path.join(REF_ROOT, "elegantrl"), mod.__file__
    return mod


def make_agent(net_dims, num_envs, *, batch_size=128, repeat_times=8.0, seed=0, **hyper):
    """The reference's AgentPPO on CPU for the Pendulum dims, Config defaults otherwise."""
    import torch as th
    load()
    from elegantrl.agents.AgentPPO import AgentPPO
    from elegantrl.train.config import Config
    env_args = {'env_name': 'Pendulum-v1', 'num_envs': num_envs, 'max_step': 200, 'state_dim': 3, 'action_dim': 1,
                'if_discrete': False}
    args = Config(agent_class=AgentPPO, env_class=None, env_args=env_args)
    args.net_dims, args.batch_size, args.repeat_times = list(net_dims), batch_size, repeat_times
    for k, v in hyper.items():
        setattr(args, k, v)
    th.manual_seed(seed)
    return AgentPPO(list(net_dims), 3, 1, gpu_id=-1, args=args)


def time_ref_cycles(num_envs, horizon_len, net_dims=(64, 64), warmup=1, cycles=3, threads=None, seed=0, **hyper):
    """Time explore_env + update_net of the reference agent, all host threads torch is told to use.
    Returns dict(env_steps_per_sec, explore_s, update_s (lists), threads, cycles)."""
    import torch as th
    from elegantrl_b200.envs import PendulumVecEnv
    if threads is None:
        threads = th.get_num_threads()
    th.set_num_threads(threads)
    agent = make_agent(net_dims, num_envs, seed=seed, **hyper)
    env = PendulumVecEnv(num_envs=num_envs, gpu_id=-1, max_step=200, seed=seed)
    agent.last_state = env.reset()[0]
    explore_s, update_s = [], []
    th.set_grad_enabled(False)   # as the reference's training loop does (run.py:41, 124-127)
    for i in range(warmup + cycles):
        t0 = time.perf_counter()
        buffer = agent.explore_env(env, horizon_len)
        t1 = time.perf_counter()
        th.set_grad_enabled(True)
        agent.update_net(buffer)
        th.set_grad_enabled(False)
        t2 = time.perf_counter()
        if i >= warmup:
            explore_s.append(t1 - t0)
            update_s.append(t2 - t1)
    th.set_grad_enabled(True)
    total = sum(explore_s) + sum(update_s)
    return dict(env_steps_per_sec=num_envs * horizon_len * cycles / total, explore_s=explore_s, update_s=update_s,
                threads=threads, cycles=cycles)

"""Diagnostic: update_net (values from the rollout cache + GAE + minibatch updates) for the update schedules of SURVEY.md 8(d):
primary = Config defaults (8 minibatches of 128), secondary = 4 minibatches of 65 536, plus two in between."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from elegantrl_b200 import Config
from elegantrl_b200.agents import AgentPPO
from elegantrl_b200.envs import PendulumVecEnv

N, H = 65536, 128
env_args = {'env_name': 'Pendulum-v1', 'num_envs': N, 'max_step': 200, 'state_dim': 3, 'action_dim': 1, 'if_discrete': False}
for batch, repeat in ((128, 8.0), (1024, 64.0), (8192, 256.0), (65536, 2048.0)):
    cfg = Config(AgentPPO, PendulumVecEnv, env_args)
    cfg.net_dims, cfg.batch_size, cfg.repeat_times, cfg.random_seed = [64, 64], batch, repeat, 0
    agent = AgentPPO([64, 64], 3, 1, gpu_id=0, args=cfg)
    env = PendulumVecEnv(num_envs=N, gpu_id=0, max_step=200, seed=0)
    agent.last_state = env.reset()[0]
    rows = []
    for it in range(7):
        buf = agent.explore_env(env, H)
        a = th.cuda.Event(enable_timing=True); b = th.cuda.Event(enable_timing=True)
        a.record(); out = agent.update_net_device(list(buf)); b.record()
        th.cuda.synchronize()
        rows.append(a.elapsed_time(b))
    rows = sorted(rows[2:])
    updates = int(H * repeat / batch)
    samples = updates * batch
    print(f"batch {batch:6d} x {updates} updates: update_net {rows[len(rows) // 2]:7.3f} ms  "
          f"({samples / (rows[len(rows) // 2] * 1e-3) / 1e6:8.1f} M samples/s, finite={bool(th.isfinite(out).all())})", flush=True)

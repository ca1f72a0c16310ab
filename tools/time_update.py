"""Diagnostic: time the phases of update_net (GAE, minibatch updates) for both update drivers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from elegantrl_b200 import Config
from elegantrl_b200.agents import AgentPPO
from elegantrl_b200.envs import PendulumVecEnv

N, H = 65536, 128
env_args = {'env_name': 'Pendulum-v1', 'num_envs': N, 'max_step': 200, 'state_dim': 3, 'action_dim': 1, 'if_discrete': False}
cfg = Config(AgentPPO, PendulumVecEnv, env_args)
cfg.net_dims, cfg.batch_size, cfg.repeat_times, cfg.random_seed = [64, 64], 128, 8.0, 0
agent = AgentPPO([64, 64], 3, 1, gpu_id=0, args=cfg)
env = PendulumVecEnv(num_envs=N, gpu_id=0, max_step=200, seed=0)
agent.last_state = env.reset()[0]
env.cur_step[:] = th.randint(0, 200, (N,), device="cuda:0", dtype=th.int32)


def ev():
    e = th.cuda.Event(enable_timing=True); e.record(); return e


for mode in ("cluster", "multilaunch", "cluster"):
    os.environ["B200RL_UPDATE"] = mode
    rows = []
    for it in range(8):
        a = ev(); buf = agent.explore_env(env, H); b = ev()
        out = agent.update_net_device(list(buf)); c = ev()
        th.cuda.synchronize()
        rows.append((round(a.elapsed_time(b), 3), round(b.elapsed_time(c), 3)))
    print(mode, "explore / update_net ms:", rows[3:])

"""Host-side cost of the two public calls of a cycle (the GPU idles for exactly this long after a synchronisation point):
wall time of ``explore_env`` / ``update_net_device`` when they only ENQUEUE work (device idle before, no sync inside)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from elegantrl_b200 import Config
from elegantrl_b200.agents import AgentPPO
from elegantrl_b200.envs import PendulumVecEnv

N, H = 65536, 128
cfg = Config(AgentPPO, PendulumVecEnv, {'env_name': 'Pendulum-v1', 'num_envs': N, 'max_step': 200, 'state_dim': 3, 'action_dim': 1, 'if_discrete': False})
cfg.net_dims, cfg.random_seed = [64, 64], 0
agent = AgentPPO([64, 64], 3, 1, gpu_id=0, args=cfg)
env = PendulumVecEnv(num_envs=N, gpu_id=0, max_step=200, seed=0)
agent.last_state = env.reset()[0]
for _ in range(5):
    agent.update_net(list(agent.explore_env(env, H)))
te = tu = ts = 0.0
K = 50
for _ in range(K):
    th.cuda.synchronize()
    a = time.perf_counter()
    buf = agent.explore_env(env, H)
    b = time.perf_counter()
    out = agent.update_net_device(list(buf))
    c = time.perf_counter()
    res = out.tolist()
    d = time.perf_counter()
    te += b - a; tu += c - b; ts += d - c
print(f"host time per call (us): explore_env {1e6 * te / K:.1f}  update_net_device {1e6 * tu / K:.1f}  tolist (GPU work + sync) {1e6 * ts / K:.1f}")

"""Debug: phase timestamps of the persistent update kernel (library built with -DB200RL_PROFILE_PHASES)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as th
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
for _ in range(4):
    agent.update_net(list(agent.explore_env(env, H)))
th.cuda.synchronize()
marks = agent._workspace[64:64 + 14 * 8].view(th.int64).cpu().numpy()
names = ["start", "gather done", "stage+X0 done", "fwd L0", "fwd L1", "fwd L2", "loss", "bwd l2", "bwd l1", "bwd l0", "grads done", "cluster.sync", "apply done", "zero+sync"]
print("phase marks of CTA 0 (actor, tile 0), last minibatch; cycles since start:")
for i in range(1, 14):
    print(f"  {names[i]:16s} +{marks[i] - marks[i - 1]:7d}   (t = {marks[i] - marks[0]:7d})")

"""Debug: clock64 phase marks of the persistent tcgen05 update kernel (B200RL_PROFILE=1; actor CTA, last minibatch)."""
import os, sys
os.environ["B200RL_PROFILE"] = "1"
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
marks = agent._workspace[64:64 + 11 * 8].view(th.int64).cpu().numpy()
names = ["start", "stage params + layer-1 planes", "gather + x~ rows written", "L1 done", "H1 + L2 done", "head + loss", "dZ2 + dH1 + G2 pass 1", "G2 pass 2",
         "dZ1 + G1", "gradients out + loss sums", "exchange + clip + Adam"]
print("phase marks of the actor CTA, last minibatch; SM cycles:")
for i in range(1, 11):
    print(f"  {names[i]:28s} +{marks[i] - marks[i - 1]:7d}   (t = {marks[i] - marks[0]:7d})")

"""Diagnostic: external-env (per-step) rollout, eager loop vs the CUDA-graph captured loop, torch CartPole / Pendulum envs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from elegantrl_b200 import Config
from elegantrl_b200.agents import AgentDiscretePPO, AgentPPO
from elegantrl_b200.envs import CartPoleVecEnv, PendulumVecEnv

H = 128
for name, agent_class, env_class, sd, ad, disc in (("cartpole", AgentDiscretePPO, CartPoleVecEnv, 4, 2, True),
                                                   ("pendulum(torch env)", AgentPPO, PendulumVecEnv, 3, 1, False)):
    for n in (4096, 65536):
        for graph in (False, True):
            args = Config(agent_class, None, {'env_name': 'x', 'num_envs': n, 'max_step': 200, 'state_dim': sd, 'action_dim': ad, 'if_discrete': disc})
            args.net_dims = [64, 32]  # not a fused-kernel shape -> external-env path
            agent = agent_class(args.net_dims, sd, ad, gpu_id=0, args=args)
            agent.cuda_graph_rollout = graph
            env = env_class(num_envs=n, gpu_id=0, max_step=200)
            agent.last_state = env.reset()[0]
            for _ in range(2):
                buf = agent.explore_env(env, H)
            th.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                buf = agent.explore_env(env, H)
            th.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
            print(f"{name:20s} N={n:6d} graph={graph!s:5s}: {1e3 * dt:8.2f} ms per {H}-step rollout = {n * H / dt / 1e6:9.1f} M env-steps/s", flush=True)

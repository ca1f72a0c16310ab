"""Diagnostic: time the fused rollout alone, back-to-back vs synchronised, and the update alone."""
import os, sys, time
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


def timed(fn, reps, sync_each):
    evs = []
    for _ in range(reps):
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record()
        evs.append((a, b))
        if sync_each:
            th.cuda.synchronize()
    th.cuda.synchronize()
    return [round(a.elapsed_time(b), 3) for a, b in evs], out


for mode in ("tc", "ffma"):
    os.environ["B200RL_ROLLOUT"] = mode
    for _ in range(3):
        buf = agent.explore_env(env, H)
    th.cuda.synchronize()
    t_sync, _ = timed(lambda: agent.explore_env(env, H), 6, True)
    t_async, buf = timed(lambda: agent.explore_env(env, H), 6, False)
    print(mode, "rollout sync ", t_sync)
    print(mode, "rollout async", t_async)
    t0 = time.perf_counter(); buf = agent.explore_env(env, H); t1 = time.perf_counter(); th.cuda.synchronize(); t2 = time.perf_counter()
    print(mode, f"host launch {1e3*(t1-t0):.3f} ms, total {1e3*(t2-t0):.3f} ms")
    print("theta range", float(env.theta.min()), float(env.theta.max()), "finite", bool(th.isfinite(env.theta).all()))
    t_upd, _ = timed(lambda: agent.update_net_device(list(agent.explore_env(env, H))), 4, False)
    print(mode, "explore+update async", t_upd)

"""Diagnostic: find what stalls the un-synchronised cycle loop (per-step GPU and CPU timings, allocator counters)."""
import gc, os, sys, time
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
for _ in range(5):
    agent.update_net_device(list(agent.explore_env(env, H)))
th.cuda.synchronize()
for label, disable_gc in (("gc on", False), ("gc off", True)):
    if disable_gc:
        gc.disable()
    st0 = th.cuda.memory_stats()
    evs, cpu = [], []
    for i in range(120):
        a = th.cuda.Event(enable_timing=True); b = th.cuda.Event(enable_timing=True); c = th.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        buf = agent.explore_env(env, H)
        t1 = time.perf_counter(); b.record()
        out = agent.update_net_device(list(buf))
        t2 = time.perf_counter(); c.record()
        evs.append((a, b, c)); cpu.append((t1 - t0, t2 - t1))
    th.cuda.synchronize()
    st1 = th.cuda.memory_stats()
    gpu = [(a.elapsed_time(b), b.elapsed_time(c)) for a, b, c in evs]
    worst = sorted(range(len(gpu)), key=lambda i: -gpu[i][0])[:4]
    print(label, "median rollout", sorted(g[0] for g in gpu)[60], "update", sorted(g[1] for g in gpu)[60])
    for i in sorted(worst):
        print(f"   step {i}: gpu rollout {gpu[i][0]:.2f} ms update {gpu[i][1]:.2f} | cpu explore {1e3*cpu[i][0]:.2f} ms update {1e3*cpu[i][1]:.2f} ms")
    print("   cpu explore max %.2f ms, cpu update max %.2f ms" % (1e3 * max(c[0] for c in cpu), 1e3 * max(c[1] for c in cpu)))
    for k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "num_sync_all_streams"):
        print("   ", k, st1.get(k, 0) - st0.get(k, 0))

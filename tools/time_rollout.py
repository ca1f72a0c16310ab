"""Time the fused Pendulum rollout kernel alone (CUDA events on the launching stream, median of K launches after W
warm-ups) for one or more ``B200RL_ROLLOUT`` implementations and env counts.

    python tools/time_rollout.py --modes tc,ts --envs 65536,32768,8192 --horizon 128 --launches 12

Prints one line per (mode, N): median / min ms, env-steps/s.  Each launch writes a fresh trajectory (> L2 at the
BASELINE size), so no explicit L2 flush is needed.  Development tool -- bench.py is the judged measurement.
"""
import argparse
import os
import statistics
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="tc")
    ap.add_argument("--envs", default="65536")
    ap.add_argument("--horizon", type=int, default=128)
    ap.add_argument("--launches", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch as th
    from elegantrl_b200 import Config
    from elegantrl_b200.agents import AgentPPO
    from elegantrl_b200.envs import PendulumVecEnv

    for n in [int(x) for x in args.envs.split(",")]:
        for mode in args.modes.split(","):
            os.environ["B200RL_ROLLOUT"] = mode
            cfg = Config(AgentPPO, PendulumVecEnv, {'env_name': 'Pendulum-v1', 'num_envs': n, 'max_step': 200, 'state_dim': 3,
                                                    'action_dim': 1, 'if_discrete': False})
            cfg.net_dims, cfg.random_seed = [64, 64], 0
            th.manual_seed(0)
            agent = AgentPPO(cfg.net_dims, 3, 1, gpu_id=0, args=cfg)
            env = PendulumVecEnv(num_envs=n, gpu_id=0, max_step=200, seed=0)
            agent.last_state = env.reset()[0]
            env.cur_step[:] = th.randint(0, 200, (n,), device="cuda:0", dtype=th.int32)
            for _ in range(args.warmup):
                agent.explore_env(env, args.horizon)
            th.cuda.synchronize()
            times = []
            for _ in range(args.launches):
                a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
                a.record()
                agent.explore_env(env, args.horizon)
                b.record()
                th.cuda.synchronize()
                times.append(a.elapsed_time(b))
            med = statistics.median(times)
            print(f"mode={mode:5s} N={n:6d} H={args.horizon}: median {med:.4f} ms  min {min(times):.4f} ms  "
                  f"{n * args.horizon / med / 1e6:.3f} G env-steps/s", flush=True)


if __name__ == "__main__":
    main()

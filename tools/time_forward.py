"""tcgen05 forward kernel (csrc/forward_tc.cu) against the CUDA-core kernel (csrc/forward.cu, B200RL_FORWARD=ffma) on
S -> 64 -> 64 -> OUT GELU nets: the per-step policy call, the values pass over a whole [H * N, S] buffer, and the
graph-captured external-env rollout (torch CartPole) built on the policy step.   python tools/time_forward.py"""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from elegantrl_b200 import Config
from elegantrl_b200.agents import AgentDiscretePPO, AgentPPO
from elegantrl_b200.envs import CartPoleVecEnv


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        th.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


def set_impl(impl):
    if impl == "ffma":
        os.environ["B200RL_FORWARD"] = "ffma"
    else:
        os.environ.pop("B200RL_FORWARD", None)


def agent_of(cls, sd, ad, disc, n):
    args = Config(cls, None, {'env_name': 'x', 'num_envs': n, 'max_step': 200, 'state_dim': sd, 'action_dim': ad, 'if_discrete': disc})
    args.net_dims = [64, 64]
    return cls(args.net_dims, sd, ad, gpu_id=0, args=args)


print("policy step (actor forward + sampling + log-prob) and values (critic forward), one launch each; median of 20, us")
for sd, ad, disc in ((3, 1, False), (8, 2, False), (11, 3, False), (4, 2, True)):
    for rows in (4096, 65536, 1 << 20):
        agent = agent_of(AgentDiscretePPO if disc else AgentPPO, sd, ad, disc, rows)
        state = th.randn((rows, sd), device="cuda:0")
        line = f"  S={sd:2d} A={ad} {'categorical' if disc else 'gaussian   '} rows={rows:8d}:"
        for impl in ("tc", "ffma"):
            set_impl(impl)
            tp = timed(lambda: agent._policy_step(state))
            tv = timed(lambda: agent.get_values(state))
            line += f"   {impl}: policy {tp:8.1f}  values {tv:8.1f}"
        print(line, flush=True)

print("values pass over a [128 x 65 536, S] buffer (get_advantages when the buffer did not come from the fused rollout), ms")
for sd in (3, 11):
    agent = agent_of(AgentPPO, sd, 1, False, 65536)
    states = th.randn((128 * 65536, sd), device="cuda:0")
    line = f"  S={sd:2d}:"
    for impl in ("tc", "ffma"):
        set_impl(impl)
        line += f"   {impl} {timed(lambda: agent.get_values(states), reps=5) / 1e3:8.3f}"
    print(line, flush=True)
    del states

print("graph-captured external-env rollout (torch CartPole, 64 x 64 nets), 128 steps")
for n in (4096, 65536):
    line = f"  N={n:6d}:"
    for impl in ("tc", "ffma"):
        set_impl(impl)
        agent = agent_of(AgentDiscretePPO, 4, 2, True, n)
        agent.cuda_graph_rollout = True
        env = CartPoleVecEnv(num_envs=n, gpu_id=0, max_step=200)
        agent.last_state = env.reset()[0]
        for _ in range(2):
            agent.explore_env(env, 128)
        th.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            agent.explore_env(env, 128)
        th.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        line += f"   {impl} {1e3 * dt:7.2f} ms = {n * 128 / dt / 1e6:7.1f} M env-steps/s"
    print(line, flush=True)

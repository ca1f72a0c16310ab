"""Where the end-to-end cycle (bench.py's ``e2e``: host buffers in and out, a blocking read of the three scalars) loses time
against the device-resident cycle: GPU-side segments between CUDA events and the host time in front of the rollout launch.

    python tools/time_e2e_gaps.py [side]      # "side": what bench.py does -- the two D2H copies on a copy stream, update_net()
                                            # with its pinned result block; default: one stream, device result + .tolist()
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from elegantrl_b200 import Config, _lib
from elegantrl_b200.agents import AgentPPO
from elegantrl_b200.envs import PendulumVecEnv

N, H = 65536, 128
side = len(sys.argv) > 1 and sys.argv[1] == "side"
cfg = Config(AgentPPO, PendulumVecEnv, {'env_name': 'Pendulum-v1', 'num_envs': N, 'max_step': 200, 'state_dim': 3, 'action_dim': 1, 'if_discrete': False})
cfg.net_dims, cfg.random_seed = [64, 64], 0
agent = AgentPPO([64, 64], 3, 1, gpu_id=0, args=cfg)
env = PendulumVecEnv(num_envs=N, gpu_id=0, max_step=200, seed=0)
agent.last_state = env.reset()[0]
host_in = th.empty((3, N), dtype=th.float32).pin_memory()
host_last_state = th.empty((N, 3), dtype=th.float32).pin_memory()
host_in.copy_(env.engine_state_block())
copy_stream = th.cuda.Stream()
lib = _lib.load()
stamp = [0.0]
orig = lib.b200rl_rollout_pendulum


def stamped(*a):
    stamp[0] = time.perf_counter()
    return orig(*a)


lib.b200rl_rollout_pendulum = stamped


def cycle(ev):
    t0 = time.perf_counter()
    ev[0].record()
    block = env.engine_state_block()
    block.copy_(host_in, non_blocking=True)
    ev[1].record()
    buffer = agent.explore_env(env, H)
    ev[2].record()
    if side:
        copy_stream.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(copy_stream):
            host_last_state.copy_(agent.last_state, non_blocking=True)
            host_in.copy_(block, non_blocking=True)
    else:
        host_last_state.copy_(agent.last_state, non_blocking=True)
        host_in.copy_(block, non_blocking=True)
    ev[3].record()
    if side:   # the public call: scalars land in pinned host memory, the host waits for the stream
        agent.update_net(list(buffer))
        ev[4].record()
        copy_stream.synchronize()
    else:
        out = agent.update_net_device(list(buffer))
        ev[4].record()
        res = out.tolist()
    t1 = time.perf_counter()
    return t0, stamp[0], t1


for _ in range(5):
    cycle([th.cuda.Event(enable_timing=True) for _ in range(5)])
K = 40
evs = [[th.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
th.cuda.synchronize()
w0 = time.perf_counter()
stamps = [cycle(evs[i]) for i in range(K)]
w1 = time.perf_counter()
names = ["H2D (issue + DMA)", "rollout (host pre-launch gap + kernel)", "D2H copies", "GAE + update"]
seg = [sum(evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(K)) / K * 1e3 for j in range(4)]
tail = sum(evs[i][4].elapsed_time(evs[i + 1][0]) for i in range(K - 1)) / (K - 1) * 1e3
pre = sum(s[1] - s[0] for s in stamps) / K * 1e6
print(f"e2e cycle ({'copy stream' if side else 'one stream'}): wall {1e6 * (w1 - w0) / K:.1f} us per cycle")
for nm, v in zip(names, seg):
    print(f"  {nm:42s} {v:8.1f} us")
print(f"  {'last kernel -> next cycle first event':42s} {tail:8.1f} us   (sync wake-up + tolist + Python)")
print(f"  host: cycle start -> rollout launch call {pre:8.1f} us")

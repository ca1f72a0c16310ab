"""Off-policy cycle at BASELINE configs[3] shapes (AgentSAC, BipedalWalker-v3 dims: S = 24, A = 4, net_dims [256, 128],
32 768 envs, batch 512; reference recipe examples/demo_DDPG_TD3_SAC.py:286-316): explore_env (per-step policy kernel around
a torch vec env) -> ReplayBuffer.update (ring-write kernel) -> update_net (four launches per minibatch).  Box2D is absent
from the image, so the env is a synthetic torch vec env with the same tensor contract and dims (linear dynamics, quadratic
cost) -- the agent-side hot path, which is what the engine owns, is the real one.  Development measurement; prints one JSON
line.  With --reference the same cycle runs on the unmodified reference agent from oracle/_ref (eager PyTorch on the same GPU).

    python tools/bench_sac.py [--reference] [--envs 32768] [--cycles 5]
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch as th


class SyntheticVecEnv:
    """state' = clip(A s + B a + noise), reward = -|s'|^2 / S; truncation every max_step steps; auto-reset."""

    def __init__(self, num_envs, state_dim, action_dim, max_step, device):
        g = th.Generator(device="cpu").manual_seed(0)
        self.num_envs, self.state_dim, self.action_dim, self.max_step, self.device = num_envs, state_dim, action_dim, max_step, device
        self.env_name, self.if_discrete = "SyntheticWalker", False
        self.A = (th.eye(state_dim) * 0.95 + 0.02 * th.randn((state_dim, state_dim), generator=g)).to(device)
        self.B = (0.3 * th.randn((action_dim, state_dim), generator=g)).to(device)
        self.state = th.zeros((num_envs, state_dim), device=device)
        self.cur = th.zeros(num_envs, dtype=th.int32, device=device)

    def reset(self):
        self.state = th.randn((self.num_envs, self.state_dim), device=self.device)
        self.cur.zero_()
        return self.state, {}

    def step(self, action):
        s = (self.state @ self.A + action @ self.B + 0.01 * th.randn_like(self.state)).clamp(-5, 5)
        reward = -(s * s).mean(dim=1)
        self.cur += 1
        truncate = self.cur >= self.max_step
        s = th.where(truncate[:, None], th.randn_like(s), s)
        self.cur = th.where(truncate, th.zeros_like(self.cur), self.cur)
        self.state = s
        return s, reward, th.zeros_like(truncate), truncate, {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--envs", type=int, default=32768)
    ap.add_argument("--horizon", type=int, default=50)        # max_step // 32 of the reference recipe
    ap.add_argument("--buffer-rows", type=int, default=200)   # time length of the ring: 200 x 32 768 = 6.5 M transitions
    ap.add_argument("--updates", type=int, default=8)
    ap.add_argument("--cycles", type=int, default=5)
    args = ap.parse_args()
    dev = th.device("cuda:0")
    th.cuda.set_device(0)
    s_dim, a_dim, net_dims, batch = 24, 4, [256, 128], 512
    if args.reference:
        sys.path.insert(0, os.path.join(REPO, "oracle", "_ref"))
        from elegantrl.agents.AgentSAC import AgentSAC
        from elegantrl.train.config import Config
        from elegantrl.train.replay_buffer import ReplayBuffer
    else:
        from elegantrl_b200 import Config
        from elegantrl_b200.agents import AgentSAC
        from elegantrl_b200.train import ReplayBuffer
    cfg = Config()
    cfg.num_envs, cfg.batch_size, cfg.learning_rate, cfg.gamma = args.envs, batch, 1e-4, 0.99
    th.manual_seed(0)
    agent = AgentSAC(net_dims, s_dim, a_dim, gpu_id=0, args=cfg)
    buffer = ReplayBuffer(max_size=args.buffer_rows, state_dim=s_dim, action_dim=a_dim, gpu_id=0, num_seqs=args.envs, args=cfg)
    env = SyntheticVecEnv(args.envs, s_dim, a_dim, 1600, dev)
    agent.last_state = env.reset()[0]

    def cycle():
        th.set_grad_enabled(False)
        items = agent.explore_env(env, args.horizon)
        buffer.update(items)
        agent.repeat_times = (args.updates + 0.5) * batch / buffer.cur_size   # update_times = int(cur_size * repeat / batch)
        th.set_grad_enabled(True)
        return agent.update_net(buffer)

    for _ in range(2):
        res = cycle()
    th.cuda.synchronize()
    t_ex = t_up = 0.0
    t0 = time.perf_counter()
    for _ in range(args.cycles):
        a = time.perf_counter()
        th.set_grad_enabled(False)
        items = agent.explore_env(env, args.horizon)
        buffer.update(items)
        th.cuda.synchronize()
        b = time.perf_counter()
        agent.repeat_times = (args.updates + 0.5) * batch / buffer.cur_size
        th.set_grad_enabled(True)
        res = agent.update_net(buffer)
        th.cuda.synchronize()
        c = time.perf_counter()
        t_ex += b - a
        t_up += c - b
    total = time.perf_counter() - t0
    print(json.dumps({"impl": "reference (eager PyTorch, same GPU)" if args.reference else "b200", "workload": "AgentSAC, S=24 A=4 net [256,128], "
                      f"{args.envs} envs x {args.horizon} steps + {args.updates} updates of batch {batch} (BASELINE configs[3] shapes, synthetic env)",
                      "env_steps_per_s": args.envs * args.horizon * args.cycles / total, "ms_per_cycle": 1e3 * total / args.cycles,
                      "explore_ms": 1e3 * t_ex / args.cycles, "update_ms": 1e3 * t_up / args.cycles,
                      "update_ms_per_minibatch": 1e3 * t_up / args.cycles / args.updates, "result": [float(x) for x in res]}))


if __name__ == "__main__":
    main()

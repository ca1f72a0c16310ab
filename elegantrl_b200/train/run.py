"""Minimal single-process training loop for boxes where the reference package is not installed.

The drop-in agent is meant to run under the reference's own ``train_agent``; this loop only
mirrors the order of calls of its single-process path (reference ``elegantrl/train/run.py:39-138``: build env ->
agent -> ``last_state = env.reset()`` -> loop { ``explore_env`` ; ``buffer[:] = items`` ; ``update_net`` ; evaluate })
without the Evaluator / checkpoint machinery, which stays out of scope.
"""
import time

import torch as th


def evaluate_vec_env(actor, env_class, env_args, device_id: int, episodes_envs: int = 256):
    """Mean undiscounted episode return of the deterministic policy ``actor(state)`` (what the reference
    Evaluator reports as avgR) on a fresh tensor vec env."""
    args = dict(env_args)
    args["num_envs"] = episodes_envs
    env = env_class(**{k: v for k, v in args.items() if k in ("num_envs", "max_step")}, gpu_id=device_id)
    state, _ = env.reset()
    returns = th.zeros(episodes_envs, device=state.device)
    alive = th.ones(episodes_envs, dtype=th.bool, device=state.device)  # first episode of every sub-env only (auto-reset envs)
    with th.no_grad():
        for _ in range(env.max_step):
            state, reward, terminal, truncate, _ = env.step(actor(state))
            returns += reward * alive
            alive &= ~(terminal | truncate)
    return float(returns.mean()), float(returns.std())


def train_agent(args, max_cycles: int = 100, eval_every: int = 10, log=print):
    """args: Config with agent_class / env_class / env_args and the usual hyper-parameters."""
    th.set_grad_enabled(False)
    if args.random_seed is None:
        args.random_seed = max(0, args.gpu_id)
    th.manual_seed(args.random_seed)
    env_kwargs = {k: v for k, v in args.env_args.items() if k in ("num_envs", "max_step")}
    env = args.env_class(**env_kwargs, gpu_id=args.gpu_id, seed=args.random_seed)
    agent = args.agent_class(args.net_dims, args.state_dim, args.action_dim, gpu_id=args.gpu_id, args=args)
    agent.last_state = env.reset()[0]
    buffer, history, start = [], [], time.time()
    for cycle in range(1, max_cycles + 1):
        buffer[:] = agent.explore_env(env, args.horizon_len)
        exp_r = buffer[3].mean().item() / max(args.reward_scale, 1e-12)
        obj_critic, obj_actor, obj_entropy = agent.update_net(buffer)
        if cycle % eval_every == 0 or cycle == max_cycles:
            avg_r, std_r = evaluate_vec_env(agent.act, args.env_class, args.env_args, args.gpu_id)
            steps = cycle * args.horizon_len * args.num_envs
            history.append((steps, avg_r))
            log(f"| cycle {cycle:5d}  env-steps {steps:.2e}  time {time.time() - start:7.1f}s  avgR {avg_r:9.2f} +- {std_r:7.2f}"
                f"  expR {exp_r:7.3f}  objC {obj_critic:8.3f}  objA {obj_actor:8.3f}  entropy {obj_entropy:6.3f}")
    return agent, history

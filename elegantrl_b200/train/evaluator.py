"""B200-native evaluation rollout: drop-in for the reference's
``elegantrl.train.evaluator.get_cumulative_rewards_and_step_from_vec_env(env, actor)`` (``elegantrl/train/evaluator.py:200-238``),
the function its ``Evaluator.evaluate_and_save`` calls for vec envs (``:56-64, 140-145``) -- SURVEY.md section 8 row f4.

The reference resets the env, runs ``max_step`` steps of the DETERMINISTIC policy ``action = actor(state)`` in a Python loop
(45-60 tiny launches per step), copies ``returns [max_step, N]`` and ``dones`` to the host and then walks every env in a
Python loop with an ``.item()`` per episode -- at 65 536 envs that loop alone takes minutes.  Here

* the rollout is ONE launch of the fused Pendulum kernel with ``B200RL_ROLLOUT_DETERMINISTIC`` (zero policy noise; the env
  receives ``tanh(mean)`` exactly as ``ActorPPO.forward`` :363-366) when the env is the built-in Pendulum vec env and the
  actor has the kernel's shape -- otherwise the same loop as the reference with the actor's own ``forward``;
* the episode segmentation (per env: cut at every done, sum the rewards of each piece, count its steps) runs on the GPU
  with a cumulative sum and one ``nonzero``, followed by ONE device-to-host copy.

Same return value: a list of ``(cumulative_return, steps)`` per finished episode, env-major, episodes of an env in time order.
"""
import ctypes as C
from typing import List, Tuple

import torch as th

from .. import _lib

TEN = th.Tensor


def episodes_from_returns_dones(returns: TEN, dones: TEN) -> List[Tuple[float, int]]:
    """``returns`` fp32 [T, N], ``dones`` bool [T, N] -> [(episode return, episode steps)] in the reference's order
    (``evaluator.py:222-237``): for every env, its finished episodes in time order; unfinished tails are dropped."""
    t_len, n = returns.shape
    csum = th.cumsum(returns.to(th.float64), dim=0)                       # inclusive prefix sums per env
    env_idx, t_idx = th.nonzero(dones.T, as_tuple=True)                   # sorted by env, then time
    if env_idx.numel() == 0:
        return []
    end_sum = csum[t_idx, env_idx]
    same_env = th.zeros_like(env_idx, dtype=th.bool)
    same_env[1:] = env_idx[1:] == env_idx[:-1]                            # previous done belongs to the same env
    prev_t = th.where(same_env, th.roll(t_idx, 1), th.full_like(t_idx, -1))
    prev_sum = th.where(same_env, th.roll(end_sum, 1), th.zeros_like(end_sum))
    out = th.stack((end_sum - prev_sum, (t_idx - prev_t).to(th.float64)), dim=1).cpu()   # the one D2H copy
    return [(float(r), int(s)) for r, s in out.tolist()]


def _fused_pendulum_ok(env, actor) -> bool:
    from ..agents.nets import ActorPPO
    if getattr(env, "env_kind", None) != "pendulum" or not isinstance(actor, ActorPPO) or env.device.type != "cuda":
        return False
    linears = [m for m in actor.net if isinstance(m, th.nn.Linear)]
    dims = [linears[0].in_features] + [layer.out_features for layer in linears]
    return dims == [3, 64, 64, 1] and getattr(actor, "activation", "gelu") == "gelu" and next(actor.parameters()).device == env.device


def rollout_returns_dones(env, actor) -> Tuple[TEN, TEN]:
    """``env.reset()`` then ``max_step`` deterministic steps: (returns [max_step, N], dones [max_step, N]) on ``env.device``."""
    device, n, max_step = env.device, env.num_envs, env.max_step
    state, _ = env.reset()   # must reset in vectorized env (reference :209)
    if _fused_pendulum_ok(env, actor):
        from ..agents.AgentPPO import AgentPPO
        lib = _lib.load()
        shim = AgentPPO.__new__(AgentPPO)          # descriptor helper only: no nets / optimizers are created
        shim._desc_cache, shim.device = {}, device
        desc = shim._build_net_desc(actor)
        f32 = th.float32
        states = th.empty((max_step, n, 3), dtype=f32, device=device)
        actions = th.empty((max_step, n, 1), dtype=f32, device=device)
        logprobs = th.empty((max_step, n), dtype=f32, device=device)
        rewards = th.empty((max_step, n), dtype=f32, device=device)
        undones = th.empty((max_step, n), dtype=th.bool, device=device)
        unmasks = th.empty((max_step, n), dtype=th.bool, device=device)
        last_state = th.empty((n, 3), dtype=f32, device=device)
        theta, theta_dot, cur_step = env.engine_state()
        # the kernel evaluates an actor AND a critic of the same 3 -> 64 -> 64 -> 1 shape: the actor stands in for the critic
        # (its "values" are not stored), which keeps the evaluation on the tcgen05 kernel
        args = _lib.RolloutArgs(actor=C.pointer(desc), critic=C.pointer(desc), num_envs=n, horizon_len=max_step,
                                max_step=max_step, reward_scale=1.0, theta=theta.data_ptr(), theta_dot=theta_dot.data_ptr(),
                                cur_step=cur_step.data_ptr(), states=states.data_ptr(), actions=actions.data_ptr(),
                                logprobs=logprobs.data_ptr(), rewards=rewards.data_ptr(), undones=undones.data_ptr(),
                                unmasks=unmasks.data_ptr(), values=None, last_state=last_state.data_ptr(), last_value=None,
                                eps=None, reset_noise=None, seed=getattr(env, "seed", 0), step_offset=env.global_step,
                                env_offset=0, flags=_lib.ROLLOUT_DETERMINISTIC)
        with th.cuda.device(device):
            _lib.check(lib.b200rl_rollout_pendulum(C.byref(args), th.cuda.current_stream(device).cuda_stream), "rollout_pendulum")
        env.global_step += max_step
        return rewards, th.logical_not(th.logical_and(undones, unmasks))
    returns = th.empty((max_step, n), dtype=th.float32, device=device)
    dones = th.empty((max_step, n), dtype=th.bool, device=device)
    with th.no_grad():
        for t in range(max_step):
            state, reward, terminal, truncate, _ = env.step(actor(state.to(device)))
            returns[t] = reward
            dones[t] = th.logical_or(terminal, truncate)
    return returns, dones


def get_cumulative_rewards_and_step_from_vec_env(env, actor) -> List[Tuple[float, int]]:
    """Drop-in for the reference function of the same name (``elegantrl/train/evaluator.py:200-238``)."""
    returns, dones = rollout_returns_dones(env, actor)
    if hasattr(env, "cumulative_returns"):  # envs that track their own episode returns (reference :219-220)
        return [(ret, env.max_step) for ret in env.cumulative_returns]
    return episodes_from_returns_dones(returns, dones)

"""Drop-in for the ``AgentPPO`` of the reference's tutorial ``helloworld/helloworld_PPO_single_file.py`` (BASELINE
configs[0]) on the same B200 engine.

The tutorial agent differs from ``elegantrl.agents.AgentPPO`` in its arithmetic;
every difference is a flag of the same CUDA kernels:

=======================================  ============================  =====================================
                                          elegantrl AgentPPO            helloworld AgentPPO (this class)
=======================================  ============================  =====================================
nets                                      GELU, state_norm              ReLU, no state_norm (:172-212)
advantage normalisation                   std over [::4, ::4]           std over the whole buffer (:296)
critic criterion                          MSE * unmask                  SmoothL1, weighted by mean(unmask)
                                                                        ([B] x [B, 1] broadcast, :325, 332)
surrogate                                 adv*ratio*const factor        min(adv*ratio, adv*clamp(ratio)) (:337-339)
entropy term                              subtracted                    added (:340); lambda_entropy 0.01
actor terms masked by unmask              yes                           no
grad clipping                             clip_grad_norm_ 3.0           none (:366-370)
rollout                                   vec / single env              single gym env, numpy state; ``logprobs`` is
                                                                        never written and stays 0 (:258-271)
buffer shapes                             [H, N, ...]                   [H, S], [H, A], [H], [H, 1] x3 (:248-277)
update_net returns                        (objC, objA_surrogate, ent)   (objC, objA_full, 0.0) (:314-317)
=======================================  ============================  =====================================
"""
from typing import Tuple

import copy

import numpy as np
import torch as th

from .. import _lib
from .AgentPPO import AgentPPO as _EngineAgentPPO

TEN = th.Tensor


class AgentPPO(_EngineAgentPPO):
    def __init__(self, net_dims, state_dim: int, action_dim: int, gpu_id: int = 0, args=None):
        if args is None:
            from ..config import Config
            args = Config()
            args.learning_rate, args.lambda_entropy = 6e-5, 0.01
        args = copy.copy(args)  # the overrides below must not leak into a Config the caller reuses (evaluator, other agents)
        for name, value in (("activation", "relu"), ("use_state_norm", False), ("clip_grad_norm", 0.0), ("num_envs", 1),
                            ("if_discrete", False)):
            setattr(args, name, value)
        if not hasattr(args, "lambda_entropy"):
            args.lambda_entropy = 0.01  # helloworld default (:241)
        super().__init__(net_dims, state_dim, action_dim, gpu_id, args)
        self._ppo_flags = _lib.PPO_HELLOWORLD
        self._full_std = True
        self.if_vec_env = False

    def explore_env(self, env, horizon_len: int, **_kwargs) -> Tuple[TEN, TEN, TEN, TEN, TEN, TEN]:
        """helloworld_PPO_single_file.py:248-277: one gym env, numpy observations, engine policy step."""
        self._require_engine()
        h, dev = int(horizon_len), self.device
        states = th.empty((h, self.state_dim), dtype=th.float32, device=dev)
        actions = th.empty((h, self.action_dim), dtype=th.float32, device=dev)
        logprobs = th.zeros(h, dtype=th.float32, device=dev)  # never written by the tutorial either (quirk #13)
        rewards = th.zeros(h, dtype=th.float32)
        terminals = th.zeros(h, dtype=th.bool)
        truncates = th.zeros(h, dtype=th.bool)
        ary_state = self.last_state
        for i in range(h):
            state = th.as_tensor(np.asarray(ary_state), dtype=th.float32, device=dev).reshape(1, self.state_dim)
            action, _, env_action = self._policy_step(state)
            ary_state, reward, terminal, truncate, _ = env.step(env_action[0].cpu().numpy())
            if terminal or truncate:
                ary_state, _ = env.reset()
            states[i], actions[i] = state[0], action[0]
            rewards[i], terminals[i], truncates[i] = float(reward), bool(terminal), bool(truncate)
        self.last_state = ary_state
        return (states, actions, logprobs, rewards.unsqueeze(1).to(dev), th.logical_not(terminals).unsqueeze(1).to(dev),
                th.logical_not(truncates).unsqueeze(1).to(dev))

    def update_net(self, buffer) -> Tuple[float, float, float]:
        """helloworld_PPO_single_file.py:283-317 on the engine; accepts the tutorial's 2-D buffer."""
        states, actions, logprobs, rewards, undones, unmasks = buffer
        h = states.shape[0]
        last_state = self.last_state
        self.last_state = th.as_tensor(np.asarray(last_state), dtype=th.float32, device=self.device).reshape(1, self.state_dim)
        rewards2, undones2 = rewards.reshape(h, 1), undones.reshape(h, 1)
        engine_buffer = [states.reshape(h, 1, self.state_dim), actions.reshape(h, 1, self.action_dim), logprobs.reshape(h, 1),
                         rewards2, undones2, unmasks.reshape(h, 1)]
        obj_critic, obj_surrogate, obj_entropy = self.update_net_device(engine_buffer).tolist()
        self.last_state = last_state
        return obj_critic, obj_surrogate + obj_entropy * float(self.lambda_entropy), 0.0

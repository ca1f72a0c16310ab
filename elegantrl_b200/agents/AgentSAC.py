"""Drop-in ``AgentSAC`` (reference ``elegantrl/agents/AgentSAC.py:16-86`` on ``AgentBase``, ``elegantrl/agents/AgentBase.py:
16-74, 130-189``) whose off-policy hot path runs in ``libb200rl.so`` -- SURVEY.md section 8 row f3, BASELINE configs[3].

Boundary mirrored: ``AgentSAC(net_dims, state_dim, action_dim, gpu_id, args)``; ``explore_env(env, horizon_len) -> (states,
actions, rewards, undones, unmasks)`` with the tanh'ed actions the ReplayBuffer stores; ``update_net(buffer) -> (obj_critic,
obj_actor)``; ``act`` / ``cri`` / ``cri_target`` as picklable ``nn.Module`` s with the reference's state_dict keys;
``alpha_log``; ``last_state``; ``save_or_load_agent``.  The class name has no on-policy marker, so ``Config.get_if_off_policy``
(``elegantrl/train/config.py:108-111``) classifies it off-policy and the reference's ``train_agent`` builds a ReplayBuffer for
it -- ``elegantrl_b200.train.replay_buffer.ReplayBuffer`` is the drop-in for that class.

What runs where
    explore_env  one ``b200rl_sac_policy_step`` kernel per time step (ActorSAC.get_action, AgentSAC.py:176-182) around the
                 env's own ``step``; no fused env here: configs[3]'s BipedalWalker is a Box2D env the engine does not own
    update_net   ``b200rl_sac_update``: ``int(cur_size * repeat_times / batch_size)`` minibatches, four launches each, the
                 replay gathers fused into them; one D2H copy of two floats
There is no fallback: without ``libb200rl.so`` or without a CUDA device the agent raises.
"""
import copy
import ctypes as C
import math
import os
from typing import Optional, Tuple

import torch as th
from torch import nn

from .. import _lib
from .AgentPPO import _on_device
from .nets import ActorSAC, CriticEnsemble

TEN = th.Tensor


def _mlp_desc(seq: nn.Sequential) -> _lib.Net:
    linears = [m for m in seq if isinstance(m, nn.Linear)]
    assert 1 <= len(linears) <= _lib.MAX_LINEAR
    net = _lib.Net()
    net.num_linear, net.activation = len(linears), _lib.ACT_GELU
    net.dims[0] = linears[0].in_features
    for i, layer in enumerate(linears):
        assert layer.weight.is_cuda and layer.weight.is_contiguous() and layer.weight.dtype == th.float32
        net.dims[i + 1] = layer.out_features
        net.weight[i], net.bias[i] = layer.weight.data_ptr(), layer.bias.data_ptr()
    return net


class AgentSAC:
    def __init__(self, net_dims, state_dim: int, action_dim: int, gpu_id: int = 0, args=None):
        if args is None:
            from ..config import Config
            args = Config()
        # ---- fields of AgentBase.__init__ (reference AgentBase.py:27-68)
        self.if_discrete = False
        self.if_off_policy = True
        self.net_dims, self.state_dim, self.action_dim = list(net_dims), int(state_dim), int(action_dim)
        self.gamma = args.gamma
        self.max_step = getattr(args, "max_step", 12345)
        self.num_envs = getattr(args, "num_envs", None) or 1
        self.batch_size = int(args.batch_size)
        self.repeat_times = args.repeat_times
        self.reward_scale = args.reward_scale
        self.learning_rate = args.learning_rate
        self.clip_grad_norm = args.clip_grad_norm
        self.soft_update_tau = getattr(args, "soft_update_tau", 5e-3)
        self.explore_noise_std = getattr(args, "explore_noise_std", 0.05)
        self.explore_rate = getattr(args, "explore_rate", 1.0)
        self.if_use_per = getattr(args, "if_use_per", False)
        self.lambda_fit_cum_r = getattr(args, "lambda_fit_cum_r", 0.0)
        assert not self.if_use_per and not self.lambda_fit_cum_r, "PER / lambda_fit_cum_r are outside the engine's scope"
        self.buffer_init_size = getattr(args, "buffer_init_size", None)
        self.last_state: Optional[TEN] = None
        self.device = th.device(f"cuda:{gpu_id}" if (th.cuda.is_available() and gpu_id >= 0) else "cpu")
        self.if_vec_env = self.num_envs > 1

        # ---- fields of AgentSAC.__init__ (reference AgentSAC.py:17-31)
        self.num_ensembles = getattr(args, "num_ensembles", 4)
        self.act = ActorSAC(self.net_dims, self.state_dim, self.action_dim).to(self.device)
        self.cri = CriticEnsemble(self.net_dims, self.state_dim, self.action_dim, self.num_ensembles).to(self.device)
        self.act_target = None
        self.cri_target = copy.deepcopy(self.cri)
        self.act_optimizer = th.optim.Adam(self.act.parameters(), self.learning_rate)
        self.cri_optimizer = th.optim.Adam(self.cri.parameters(), self.learning_rate)
        self.alpha_log = th.tensor((-1,), dtype=th.float32, requires_grad=True, device=self.device)
        self.alpha_optim = th.optim.Adam((self.alpha_log,), lr=self.learning_rate)
        self.target_entropy = math.log(self.action_dim)
        self.save_attr_names = {"act", "act_target", "act_optimizer", "cri", "cri_target", "cri_optimizer"}

        seed = getattr(args, "random_seed", None)
        self.seed = int(max(0, gpu_id) if seed is None else seed)
        self._policy_steps = 0
        self._update_draws = 0
        self._workspace: Optional[TEN] = None
        self.last_update_info = {}

    def __getstate__(self):
        state = dict(self.__dict__)
        state.update(_workspace=None)
        return state

    # --------------------------------------------------------------------------------------- plumbing
    def _require_engine(self):
        if self.device.type != "cuda":
            raise _lib.B200RLError("AgentSAC (B200 engine) needs a CUDA device; there is no CPU path")
        return _lib.load()

    def _stream(self) -> int:
        return th.cuda.current_stream(self.device).cuda_stream

    def _actor_desc(self) -> _lib.SacActor:
        return _lib.SacActor(net_s=_mlp_desc(self.act.net_s), net_a=_mlp_desc(self.act.net_a))

    @staticmethod
    def _critic_desc(cri: CriticEnsemble) -> _lib.SacCritic:
        desc = _lib.SacCritic(encoder=_mlp_desc(cri.encoder_sa), num_ensembles=cri.num_ensembles)
        for i, dec in enumerate(cri.decoder_qs):
            desc.decoder[i] = _mlp_desc(dec)
        return desc

    @staticmethod
    def _group_desc(optimizer: th.optim.Adam) -> _lib.ParamGroup:
        """C view of one torch Adam: parameters in optimizer order with their exp_avg / exp_avg_sq (created here, exactly as
        torch creates them, if the optimizer has not stepped yet)."""
        pg = optimizer.param_groups[0]
        assert not pg.get("amsgrad", False) and pg.get("weight_decay", 0) == 0 and not pg.get("maximize", False)
        grp = _lib.ParamGroup()
        params = pg["params"]
        assert len(params) <= _lib.MAX_GROUP_TENSORS
        grp.num_tensors = len(params)
        grp.lr, grp.eps = pg["lr"], pg["eps"]
        grp.beta1, grp.beta2 = pg["betas"]
        step = None
        for i, p in enumerate(params):
            st = optimizer.state[p]
            if len(st) == 0:
                st["step"] = th.tensor(0.0, dtype=th.float32)
                st["exp_avg"] = th.zeros_like(p, memory_format=th.preserve_format)
                st["exp_avg_sq"] = th.zeros_like(p, memory_format=th.preserve_format)
            assert p.is_cuda and p.is_contiguous() and p.dtype == th.float32
            grp.param[i], grp.exp_avg[i], grp.exp_avg_sq[i] = p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            grp.numel[i] = p.numel()
            step = float(st["step"]) if step is None else step
        grp.step = int(step or 0)
        return grp

    @staticmethod
    def _set_steps(optimizer, step: int):
        for p in optimizer.param_groups[0]["params"]:
            optimizer.state[p]["step"] = th.tensor(float(step), dtype=th.float32)

    # ---------------------------------------------------------------------------------------- rollout
    @_on_device
    def explore_action(self, state: TEN, eps: Optional[TEN] = None) -> TEN:
        """ActorSAC.get_action through the engine (reference AgentSAC.py:33-40, 176-182): tanh(mean + std * eps)."""
        lib = self._require_engine()
        state = state.to(self.device, th.float32).contiguous()
        rows = state.shape[0]
        action = th.empty((rows, self.action_dim), dtype=th.float32, device=self.device)
        desc = self._actor_desc()
        _lib.check(lib.b200rl_sac_policy_step(C.byref(desc), state.data_ptr(), rows, _lib.ptr(eps), self.seed, self._policy_steps,
                                              0, action.data_ptr(), self._stream()), "sac_policy_step")
        self._policy_steps += 1
        return action

    def explore_env(self, env, horizon_len: int) -> Tuple[TEN, TEN, TEN, TEN, TEN]:
        """AgentBase.explore_env dispatch + _explore_vec_env / _explore_one_env, off-policy flavour (AgentBase.py:70-170)."""
        self._require_engine()
        h, dev, n = int(horizon_len), self.device, self.num_envs
        states = th.empty((h, n, self.state_dim), dtype=th.float32, device=dev)
        actions = th.empty((h, n, self.action_dim), dtype=th.float32, device=dev)
        noise = getattr(self, "_inject_eps", None)   # parity tests: [H, N, A]
        state = self.last_state.to(dev)
        if self.if_vec_env:
            rewards = th.empty((h, n), dtype=th.float32, device=dev)
            terminals = th.empty((h, n), dtype=th.bool, device=dev)
            truncates = th.empty((h, n), dtype=th.bool, device=dev)
            for t in range(h):
                action = self.explore_action(state, None if noise is None else noise[t].contiguous())
                states[t], actions[t] = state, action
                state, reward, terminal, truncate, _ = env.step(action)
                state = state.to(dev)
                rewards[t], terminals[t], truncates[t] = reward, terminal, truncate
            rewards *= self.reward_scale
        else:
            import numpy as np
            rewards_h = th.zeros((h, 1), dtype=th.float32)
            terminals = th.zeros((h, 1), dtype=th.bool)
            truncates = th.zeros((h, 1), dtype=th.bool)
            for t in range(h):
                action = self.explore_action(state, None if noise is None else noise[t].contiguous())
                states[t], actions[t] = state, action
                ary_state, reward, terminal, truncate, _ = env.step(action[0].cpu().numpy())
                if terminal or truncate:
                    ary_state, _ = env.reset()
                state = th.as_tensor(np.asarray(ary_state), dtype=th.float32, device=dev).reshape(1, self.state_dim)
                rewards_h[t], terminals[t], truncates[t] = float(reward), bool(terminal), bool(truncate)
            rewards = (rewards_h * self.reward_scale).to(dev)
            terminals, truncates = terminals.to(dev), truncates.to(dev)
        self.last_state = state
        return states, actions, rewards, th.logical_not(terminals), th.logical_not(truncates)

    # ----------------------------------------------------------------------------------------- update
    def update_net(self, buffer) -> Tuple[float, float]:
        """AgentBase.update_net, off-policy (AgentBase.py:172-189): (mean obj_critic, mean obj_actor)."""
        out = self.update_net_device(buffer)
        if out is None:
            return 0.0, 0.0   # update_times == 0: the reference returns the defaults of its empty lists
        obj_critic, obj_actor = out.tolist()
        return obj_critic, obj_actor

    @_on_device
    def update_net_device(self, buffer) -> Optional[TEN]:
        lib = self._require_engine()
        update_times = int(buffer.cur_size * self.repeat_times / self.batch_size)
        if update_times < 1:
            return None
        actor, critic, target = self._actor_desc(), self._critic_desc(self.cri), self._critic_desc(self.cri_target)
        g_act, g_cri, g_alpha = (self._group_desc(o) for o in (self.act_optimizer, self.cri_optimizer, self.alpha_optim))
        need = int(lib.b200rl_sac_workspace_bytes(C.byref(actor), C.byref(critic), self.batch_size))
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = th.zeros(need, dtype=th.uint8, device=self.device)
        hp = _lib.SacHyper(gamma=float(self.gamma), soft_update_tau=float(self.soft_update_tau),
                           clip_grad_norm=float(self.clip_grad_norm or 0.0), target_entropy=float(self.target_entropy))
        out = th.empty(2, dtype=th.float32, device=self.device)
        ids = getattr(self, "_inject_ids", None)          # parity tests: [update_times, batch]
        eps_next = getattr(self, "_inject_eps_next", None)  # [update_times, batch, A]
        eps_pg = getattr(self, "_inject_eps_pg", None)
        rb = buffer.descriptor()
        _lib.check(lib.b200rl_sac_update(C.byref(actor), C.byref(critic), C.byref(target), C.byref(g_act), C.byref(g_cri),
                                         C.byref(g_alpha), C.byref(rb), int(buffer.cur_size), C.byref(hp), self.batch_size,
                                         update_times, _lib.ptr(ids), _lib.ptr(eps_next), _lib.ptr(eps_pg), self.seed,
                                         self._update_draws, out.data_ptr(), self._workspace.data_ptr(), self._workspace.numel(),
                                         self._stream()), "sac_update")
        self._update_draws += update_times
        for opt, grp in ((self.act_optimizer, g_act), (self.cri_optimizer, g_cri), (self.alpha_optim, g_alpha)):
            self._set_steps(opt, grp.step)
        th.autograd.graph.increment_version(list(self.act.parameters()) + list(self.cri.parameters()) +
                                            list(self.cri_target.parameters()) + [self.alpha_log])
        self.last_update_info = dict(update_times=update_times)
        return out

    # ---------------------------------------------------------------------------------- checkpointing
    def save_or_load_agent(self, cwd: str, if_save: bool):
        """Whole-object ``th.save`` / ``th.load``, file names as the reference (AgentBase.py:280-297)."""
        for attr_name in sorted(self.save_attr_names):
            obj = getattr(self, attr_name)
            if obj is None:
                continue
            file_path = f"{cwd}/{attr_name}.pth"
            if if_save:
                th.save(obj, file_path)
            elif os.path.isfile(file_path):
                setattr(self, attr_name, th.load(file_path, map_location=self.device, weights_only=False))
        if not if_save:
            for opt_name, module in (("act_optimizer", self.act), ("cri_optimizer", self.cri)):
                loaded = getattr(self, opt_name)
                fresh = th.optim.Adam(module.parameters(), self.learning_rate)
                for p_old, p_new in zip(loaded.param_groups[0]["params"], fresh.param_groups[0]["params"]):
                    if p_old in loaded.state and len(loaded.state[p_old]):
                        fresh.state[p_new] = {k: (v.to(p_new.device) if k != "step" and th.is_tensor(v) else v)
                                              for k, v in loaded.state[p_old].items()}
                setattr(self, opt_name, fresh)

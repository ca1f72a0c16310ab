"""Drop-in ``AgentPPO`` whose hot path runs in ``libb200rl.so`` (hand-written sm_100a CUDA).

Boundary mirrored: the duck-typed Agent API the reference's training loops call --
``agent_class(net_dims, state_dim, action_dim, gpu_id, args)``, ``explore_env(env, horizon_len) -> 6 tensors``,
``update_net(buffer) -> 3 floats``, ``act`` / ``cri`` as picklable ``nn.Module`` s, ``last_state``, ``device``,
``explore_rate``, ``if_off_policy``, ``save_or_load_agent`` (reference ``elegantrl/train/run.py:47, 104, 125-126,
136``; agent ``elegantrl/agents/AgentPPO.py:12-232``; base ``elegantrl/agents/AgentBase.py:16-74, 239-297``).
The class name contains "PPO" because ``Config.get_if_off_policy`` classifies agents by name
(``elegantrl/train/config.py:108-111``).

What runs where
    explore_env   built-in Pendulum env -> one fused persistent kernel (csrc/rollout_tc.cu: tcgen05; csrc/rollout.cu:
                  FP32 pipe), which also produces V(s_t) and V(last_state); any other vec env -> one policy-step
                  kernel per step (csrc/forward.cu) around the env's own ``step``, optionally with the whole loop
                  captured in a CUDA graph (``cuda_graph_rollout``)
    update_net    values (only if not already produced by the fused rollout) -> GAE reverse scan (csrc/gae.cu)
                  -> ``update_times`` fused minibatch kernels (csrc/update.cu); one D2H copy of 3 floats
PyTorch here is plumbing only: device memory, streams, ``torch.distributed``.  There is no fallback: without
``libb200rl.so`` or without a CUDA device the agent raises.
"""
import ctypes as C
import functools
import os
from typing import Optional, Tuple

import numpy as np
import torch as th
from torch import nn

from .. import _lib
from .nets import ActorDiscretePPO, ActorPPO, CriticPPO

TEN = th.Tensor


def _on_device(method):
    """Run an engine entry point with ``self.device`` as the CURRENT CUDA device: the C-ABI launches on the stream handle it
    is given (the default stream's handle is 0 = "the current device's"), and ``cudaGetDevice`` / SM-count queries inside
    the library read the current device.  The reference's learners construct agents with ``gpu_id = k`` without ever
    calling ``set_device`` (``elegantrl/train/run.py:237``)."""
    @functools.wraps(method)
    def wrapped(self, *args, **kwargs):
        if self.device.type != "cuda" or th.cuda.current_device() == self.device.index:
            return method(self, *args, **kwargs)
        with th.cuda.device(self.device):
            return method(self, *args, **kwargs)
    return wrapped


def _linears(module: nn.Module):
    return [m for m in module.net if isinstance(m, nn.Linear)]


def _has_std(module: nn.Module) -> bool:
    """Gaussian actors train ``action_std_log``; the categorical actor only inherits it as dead weight (reference
    AgentPPO.py:393-397) and the critic has none."""
    return hasattr(module, "action_std_log") and not isinstance(module, ActorDiscretePPO)


def _trainable(module: nn.Module):
    """Trainable tensors in the engine's flat order: W0, b0, W1, b1, ..., action_std_log."""
    out = []
    for layer in _linears(module):
        out += [layer.weight, layer.bias]
    if _has_std(module):
        out.append(module.action_std_log)
    return out


class AgentPPO:
    """PPO + GAE with the reference's arithmetic (incl. its quirks), B200-native engine."""
    _categorical = False  # AgentDiscretePPO flips it: Categorical policy, int32 actions [H, N]

    def __init__(self, net_dims, state_dim: int, action_dim: int, gpu_id: int = 0, args=None):
        if args is None:
            from ..config import Config
            args = Config()
        # ---- fields of AgentBase.__init__ (reference AgentBase.py:27-68)
        self.if_discrete = getattr(args, "if_discrete", False)
        self.if_off_policy = False
        self.net_dims = list(net_dims)
        self.state_dim = int(state_dim)
        self.action_dim = int(action_dim)
        self.gamma = args.gamma
        self.max_step = getattr(args, "max_step", 12345)
        self.num_envs = getattr(args, "num_envs", None) or 1
        self.batch_size = int(args.batch_size)
        self.repeat_times = args.repeat_times
        self.reward_scale = args.reward_scale
        self.learning_rate = args.learning_rate
        self.clip_grad_norm = args.clip_grad_norm
        self.soft_update_tau = getattr(args, "soft_update_tau", 5e-3)
        self.state_value_tau = getattr(args, "state_value_tau", 0)
        self.explore_noise_std = getattr(args, "explore_noise_std", 0.05)
        self.explore_rate = getattr(args, "explore_rate", 1.0)  # read by run.py:126 for every agent
        self.last_state: Optional[TEN] = None
        self.device = th.device(f"cuda:{gpu_id}" if (th.cuda.is_available() and gpu_id >= 0) else "cpu")
        self.if_vec_env = self.num_envs > 1
        assert bool(self.if_discrete) == self._categorical, \
            "AgentPPO is the continuous-action agent, AgentDiscretePPO the discrete one (env_args['if_discrete'])"

        # ---- fields of AgentPPO.__init__ (reference AgentPPO.py:18-32)
        activation = getattr(args, "activation", "gelu")
        state_norm = getattr(args, "use_state_norm", True)
        actor_class = ActorDiscretePPO if self._categorical else ActorPPO
        self.act = actor_class(self.net_dims, self.state_dim, self.action_dim, activation, state_norm).to(self.device)
        self.cri = CriticPPO(self.net_dims, self.state_dim, self.action_dim, activation, state_norm).to(self.device)
        self.act_target = self.cri_target = None
        self.act_optimizer = th.optim.Adam(self.act.parameters(), self.learning_rate)
        self.cri_optimizer = th.optim.Adam(self.cri.parameters(), self.learning_rate)
        self.ratio_clip = getattr(args, "ratio_clip", 0.25)
        self.lambda_gae_adv = getattr(args, "lambda_gae_adv", 0.95)
        self.lambda_entropy = getattr(args, "lambda_entropy", 0.001)
        self.if_use_v_trace = getattr(args, "if_use_v_trace", True)
        self.save_attr_names = {"act", "act_target", "act_optimizer", "cri", "cri_target", "cri_optimizer"}

        # ---- engine state (never pickled with act/cri: raw pointers are rebuilt on every call)
        seed = getattr(args, "random_seed", None)
        self.seed = int(max(0, gpu_id) if seed is None else seed)
        self._update_draws = 0       # Philox offset of the minibatch index stream
        self._policy_steps = 0       # Philox offset of the per-step policy kernel (external envs)
        self._workspace: Optional[TEN] = None
        self._value_cache = None     # (key, values [H, N], last_value [N]) produced by the fused rollout
        self._rollout_id = 0         # generation counter of fused rollouts (part of the value-cache key)
        self._lib_handle = None      # ctypes handle (not pickled: see __getstate__)
        self._fused_pre = None       # ((n, h), output tensors) allocated ahead for the next fused rollout
        self._rollout_args = None    # reused b200rl_rollout_args block
        self._rollout_args_key = None
        self._rollout_args_keep = None
        self._desc_cache = {}        # id(module) -> (parameter pointers, descriptor)
        self._dist_group = None      # set by enable_data_parallel()
        self._px = None
        self._host_result = None     # pinned host block the update kernel writes the three logged scalars to
        self.pinned_result = True
        self._rank, self._world = 0, 1
        self.last_update_info = {}
        self.cuda_graph_rollout = bool(getattr(args, "cuda_graph_rollout", False))  # external envs: see _explore_vec_env_graphed
        self._rollout_graphs = {}
        self._ppo_flags = 0          # b200rl_ppo_hyper.flags: 0 = the reference's elegantrl arithmetic
        self._full_std = False       # advantage std over the whole buffer instead of the [::4, ::4] lattice

    # ------------------------------------------------------------------------------------ plumbing
    def __getstate__(self):
        """The agent may be pickled (multiprocessing): raw handles, pointer blocks and CUDA graphs are per process."""
        state = dict(self.__dict__)
        state.update(_lib_handle=None, _fused_pre=None, _rollout_args=None, _rollout_args_key=None, _rollout_args_keep=None, _desc_cache={}, _rollout_graphs={},
                     _workspace=None, _value_cache=None, _px=None, _dist_group=None, _host_result=None)
        return state

    def _require_engine(self):
        if self.device.type != "cuda":
            raise _lib.B200RLError("AgentPPO (B200 engine) needs a CUDA device; there is no CPU path "
                                   "(the CPU oracle lives in oracle/ and is test infrastructure only)")
        return _lib.load()

    @staticmethod
    def _check(t: TEN, name: str) -> TEN:
        assert t.dtype == th.float32 and t.is_contiguous() and t.is_cuda, f"{name}: need contiguous fp32 CUDA tensor"
        return t

    def _net_desc(self, module: nn.Module) -> _lib.Net:
        """C descriptor aliasing the module's parameter storages; cached per module and revalidated on every call by the
        storages' addresses (``load_state_dict`` keeps them, ``.to()`` does not; replacing the module rebuilds it, replacing a
        LAYER of a live module after the first call is not supported -- the reference never does)."""
        hit = self._desc_cache.get(id(module))
        if hit is not None and hit[2] is module:
            # fast path (this sits between a host synchronisation and the rollout launch): the Parameter / buffer OBJECTS seen
            # when the descriptor was built are kept alive here, so comparing their current addresses needs no module traversal
            ptrs, net, _, tensors = hit
            for t, p in zip(tensors, ptrs):
                if t.data_ptr() != p:
                    break
            else:
                return net
        tensors = tuple(module.parameters()) + tuple(module.buffers())
        ptrs = tuple(t.data_ptr() for t in tensors)
        net = self._build_net_desc(module)
        self._desc_cache[id(module)] = (ptrs, net, module, tensors)
        return net

    def _build_net_desc(self, module: nn.Module) -> _lib.Net:
        linears = _linears(module)
        assert 1 <= len(linears) <= _lib.MAX_LINEAR, "too many layers for the engine"
        net = _lib.Net()
        net.num_linear = len(linears)
        net.activation = _lib.ACTIVATION_CODES[getattr(module, "activation", "gelu")]
        net.dims[0] = linears[0].in_features
        for i, layer in enumerate(linears):
            net.dims[i + 1] = layer.out_features
            net.weight[i] = self._check(layer.weight.data, "weight").data_ptr()
            net.bias[i] = self._check(layer.bias.data, "bias").data_ptr()
        if getattr(module, "state_avg", None) is not None:
            net.state_avg = self._check(module.state_avg.data, "state_avg").data_ptr()
            net.state_std = self._check(module.state_std.data, "state_std").data_ptr()
        if _has_std(module):
            net.action_std_log = self._check(module.action_std_log.data, "action_std_log").data_ptr()
        return net

    def _adam_desc(self, optimizer: th.optim.Adam, module: nn.Module) -> _lib.Adam:
        """C descriptor aliasing the optimizer's exp_avg / exp_avg_sq tensors (created here if Adam has not
        stepped yet, in exactly the form torch.optim.Adam itself would create them)."""
        group = optimizer.param_groups[0]
        assert not group.get("amsgrad", False) and group.get("weight_decay", 0) == 0 and not group.get("maximize", False)
        adam = _lib.Adam()
        adam.lr, adam.eps = group["lr"], group["eps"]
        adam.beta1, adam.beta2 = group["betas"]
        step = None
        linears = _linears(module)
        for i, layer in enumerate(linears):
            for p, avg, sq in ((layer.weight, adam.exp_avg_w, adam.exp_avg_sq_w), (layer.bias, adam.exp_avg_b, adam.exp_avg_sq_b)):
                st = self._adam_state(optimizer, p)
                avg[i], sq[i] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                step = float(st["step"]) if step is None else step
        if _has_std(module):
            st = self._adam_state(optimizer, module.action_std_log)
            adam.exp_avg_std, adam.exp_avg_sq_std = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
        adam.step = int(step or 0)
        return adam

    def _adam_state(self, optimizer, p):
        st = optimizer.state[p]
        if len(st) == 0:
            st["step"] = th.tensor(0.0, dtype=th.float32)
            st["exp_avg"] = th.zeros_like(p, memory_format=th.preserve_format)
            st["exp_avg_sq"] = th.zeros_like(p, memory_format=th.preserve_format)
        self._check(st["exp_avg"], "exp_avg")
        self._check(st["exp_avg_sq"], "exp_avg_sq")
        return st

    @staticmethod
    def _set_adam_step(optimizer, module, step: int):
        for p in _trainable(module):
            optimizer.state[p]["step"] = th.tensor(float(step), dtype=th.float32)

    def _stream(self) -> int:
        return th.cuda.current_stream(self.device).cuda_stream

    def _get_workspace(self, act_desc, cri_desc) -> TEN:
        lib = _lib.load()
        need = lib.b200rl_workspace_bytes(C.byref(act_desc), C.byref(cri_desc))
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != self.device:
            self._workspace = th.zeros(need, dtype=th.uint8, device=self.device)
        return self._workspace

    def enable_data_parallel(self, group=None):
        """Shard envs over the ranks of a torch.distributed group: each rank rolls out / scans its own env slice;
        per cycle one all-reduce of the advantage sums, one all-gather of the packed minibatch records and one
        broadcast of parameters + moments (``sharded_mode = "gather"``, default), or one all-reduce of the flat
        gradient per minibatch (``"allreduce"``) -- SURVEY.md section 8(e).  Replaces the reference's host-pipe
        trajectory all-gather (run.py:305-320)."""
        import torch.distributed as dist
        self._dist_group = group if group is not None else dist.group.WORLD
        self._rank, self._world = dist.get_rank(self._dist_group), dist.get_world_size(self._dist_group)
        self._px = None  # peer-memory exchange (sharded_mode "peer"): allocated at the first sharded update

    def _peer_exchange(self, lib, act_desc, cri_desc, update_times: int):
        """Symmetric (peer-mapped) exchange buffer + flag array of the in-kernel exchange (``b200rl_ppo_update_sharded``).
        torch's symmetric-memory allocator is the plumbing: it allocates the same buffer on every rank of the group and
        maps every peer's copy into this process (NVLink P2P).  Allocated at the first sharded update (a collective: every
        rank gets here with the same ``update_times``), again only if a later schedule needs more room."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        floats = int(lib.b200rl_peer_exchange_floats(C.byref(act_desc), C.byref(cri_desc), self.batch_size // self._world,
                                                     update_times))
        if self._px is not None and self._px[1].numel() >= floats:
            return self._px[0]
        old = self._px
        data = symm.empty(floats, dtype=th.float32, device=self.device)
        flags = symm.empty(_lib.PX_FLAGS, dtype=th.int32, device=self.device)
        hd, hf = symm.rendezvous(data, self._dist_group), symm.rendezvous(flags, self._dist_group)
        data.zero_()
        flags.zero_()
        th.cuda.synchronize(self.device)
        dist.barrier(group=self._dist_group)   # every rank's flags are zero before anybody raises one
        px = _lib.PeerExchange(rank=self._rank, world=self._world, epoch=0 if old is None else old[0].epoch,
                               reserved=0 if old is None else old[0].reserved)
        for r in range(self._world):
            px.data[r], px.flags[r] = int(hd.buffer_ptrs[r]), int(hf.buffer_ptrs[r])
        self._px = (px, data, flags, hd, hf)
        return px

    def _peer_mode(self, lib, act_desc, cri_desc, hp, update_times: int) -> bool:
        """Use the in-kernel exchange?  ``sharded_mode``: "peer" (record gather once per cycle) / "peer_allreduce" (gradient
        all-reduce per minibatch), both inside the update kernel over peer memory; "auto" (default: "peer" when the nets have
        the tcgen05 kernel's shape and symmetric memory can be set up, else "gather"); "gather", "allreduce": NCCL."""
        mode = getattr(self, "sharded_mode", "auto")
        if mode not in ("auto", "peer", "peer_allreduce"):
            return False
        ok = (bool(lib.b200rl_update_tc_supported(C.byref(act_desc), C.byref(cri_desc), C.byref(hp)))
              and self.batch_size % self._world == 0 and self.batch_size // self._world <= 128 and self._world <= _lib.MAX_PEERS)
        if ok and mode != "peer_allreduce" and self.batch_size > 128:
            ok = False   # the record gather runs the whole minibatch in one 128-sample tile
        if ok:
            try:
                self._peer_exchange(lib, act_desc, cri_desc, update_times)
            except Exception as err:  # noqa: BLE001
                if mode != "auto":
                    raise
                import warnings
                warnings.warn(f"AgentPPO: peer-memory exchange unavailable ({err!r}); using the NCCL all-gather mode")
                self.sharded_mode = "gather"
                return False
        if not ok and mode != "auto":
            raise _lib.B200RLError(f"sharded_mode='{mode}' needs S -> 64 -> 64 -> OUT GELU nets and batch_size <= 128 (peer) / "
                                   "batch_size / world <= 128 (peer_allreduce)")
        return ok

    # ------------------------------------------------------------------------------------- rollout
    @_on_device
    def explore_env(self, env, horizon_len: int) -> Tuple[TEN, TEN, TEN, TEN, TEN, TEN]:
        """Reference AgentBase.explore_env dispatch (AgentBase.py:70-74) + AgentPPO._explore_vec_env /
        _explore_one_env (AgentPPO.py:34-129).  Returns (states, actions, logprobs, rewards, undones, unmasks)."""
        self._require_engine()
        if getattr(env, "env_kind", None) == "pendulum" and self._fused_rollout_ok(env):
            return self._explore_fused_pendulum(env, horizon_len)
        if self.if_vec_env:
            if self.cuda_graph_rollout and getattr(env, "device", None) == self.device:
                return self._explore_vec_env_graphed(env, horizon_len)
            return self._explore_vec_env(env, horizon_len)
        return self._explore_one_env(env, horizon_len)

    def _fused_rollout_ok(self, env) -> bool:
        return (env.device == self.device and self.state_dim == 3 and self.action_dim == 1
                and tuple(self.net_dims) in ((64, 64), (128, 64)) and env.num_envs == self.num_envs)

    def _fused_outputs(self, n: int, h: int):
        dev, f32 = self.device, th.float32
        return (th.empty((h, n, 3), dtype=f32, device=dev), th.empty((h, n, 1), dtype=f32, device=dev),
                th.empty((h, n), dtype=f32, device=dev), th.empty((h, n), dtype=f32, device=dev),
                th.empty((h, n), dtype=th.bool, device=dev), th.empty((h, n), dtype=th.bool, device=dev),
                th.empty((h, n), dtype=f32, device=dev), th.empty((n, 3), dtype=f32, device=dev),
                th.empty((n,), dtype=f32, device=dev))

    def _explore_fused_pendulum(self, env, horizon_len: int):
        """One launch.  The host work between the caller's last synchronisation and this launch is the end-to-end
        critical path (the GPU idles meanwhile), so it is kept minimal: the nine output tensors of THIS call were
        allocated at the end of the previous call (while the GPU was busy), the net descriptors are cached, the
        argument block is reused."""
        lib = self._lib_handle
        if lib is None:
            lib = self._lib_handle = _lib.load()
        n, h = env.num_envs, int(horizon_len)
        pre, self._fused_pre = self._fused_pre, None
        outs = pre[1] if (pre is not None and pre[0] == (n, h)) else self._fused_outputs(n, h)
        states, actions, logprobs, rewards, undones, unmasks, values, last_state, last_value = outs
        theta, theta_dot, cur_step = env.engine_state()
        act_desc, cri_desc = self._net_desc(self.act), self._net_desc(self.cri)
        eps = getattr(self, "_inject_eps", None)
        reset_noise = getattr(self, "_inject_reset_noise", None)
        args = self._rollout_args
        if args is None:
            args = self._rollout_args = _lib.RolloutArgs()
            self._rollout_args_key = None
        # the fields that do not change from cycle to cycle are written once per (env, nets, shape)
        key = (id(env), id(act_desc), id(cri_desc), n, h, env.max_step, id(theta), id(theta_dot), id(cur_step))
        if self._rollout_args_key != key:
            args.actor, args.critic = C.pointer(act_desc), C.pointer(cri_desc)
            args.num_envs, args.horizon_len, args.max_step = n, h, env.max_step
            args.theta, args.theta_dot, args.cur_step = theta.data_ptr(), theta_dot.data_ptr(), cur_step.data_ptr()
            self._rollout_args_key = key
            self._rollout_args_keep = (env, act_desc, cri_desc, theta, theta_dot, cur_step)   # ids above stay unique while these live
        args.reward_scale = float(self.reward_scale)
        args.states, args.actions, args.logprobs, args.rewards = states.data_ptr(), actions.data_ptr(), logprobs.data_ptr(), rewards.data_ptr()
        args.undones, args.unmasks, args.values = undones.data_ptr(), unmasks.data_ptr(), values.data_ptr()
        args.last_state, args.last_value = last_state.data_ptr(), last_value.data_ptr()
        args.eps, args.reset_noise = _lib.ptr(eps), _lib.ptr(reset_noise)
        args.seed, args.step_offset, args.env_offset = self.seed, env.global_step, self._rank * n
        _lib.check(lib.b200rl_rollout_pendulum(C.byref(args), self._stream()), "rollout_pendulum")
        env.global_step += h
        self.last_state = last_state
        self._rollout_id += 1
        states._b200rl_rollout_id = self._rollout_id   # the tag travels with THIS tensor object (a recycled address cannot alias it)
        self._value_cache = (self._cache_key(states), values, last_value)
        self._fused_pre = ((n, h), self._fused_outputs(n, h))   # the next call's outputs, allocated while the GPU is busy
        return states, actions, logprobs, rewards, undones, unmasks

    def _cache_key(self, states: TEN):
        """Identity of (this very rollout's states tensor, this very critic state): the rollout generation tag set on the
        tensor object, and the in-place version counter of every critic tensor (``load_state_dict``, an optimizer step or
        the engine's own update -- which bumps them explicitly -- invalidate the cached values)."""
        versions = tuple(p._version for p in self.cri.parameters()) + tuple(b._version for b in self.cri.buffers())
        return getattr(states, "_b200rl_rollout_id", None), tuple(states.shape), id(self.cri), versions

    def explore_action(self, state: TEN) -> Tuple[TEN, TEN]:
        """ActorPPO.get_action through the engine (reference AgentPPO.py:131-133): (action, logprob)."""
        action, logprob, _ = self._policy_step(state)
        return action, logprob

    @_on_device
    def _policy_step(self, state: TEN, eps: Optional[TEN] = None):
        """One engine exploration step: (action [rows, A] pre-tanh, logprob [rows], env_action = tanh(action));
        categorical policy: (action int32 [rows], logprob [rows], env_action = action.long())."""
        lib = self._require_engine()
        state = self._check(state.to(self.device, th.float32).contiguous(), "state")
        rows = state.shape[0]
        if self._categorical:
            action = th.empty((rows,), dtype=th.int32, device=self.device)
            logprob = th.empty((rows,), dtype=th.float32, device=self.device)
            act_desc = self._net_desc(self.act)
            _lib.check(lib.b200rl_policy_step_discrete(C.byref(act_desc), None, _lib.ptr(state), rows, _lib.ptr(eps), self.seed,
                                                       self._policy_steps, self._rank * rows, _lib.ptr(action),
                                                       _lib.ptr(logprob), None, self._stream()), "policy_step_discrete")
            self._policy_steps += 1
            return action, logprob, action.long()
        action = th.empty((rows, self.action_dim), dtype=th.float32, device=self.device)
        env_action = th.empty_like(action)
        logprob = th.empty((rows,), dtype=th.float32, device=self.device)
        act_desc = self._net_desc(self.act)
        _lib.check(lib.b200rl_policy_step(C.byref(act_desc), None, _lib.ptr(state), rows, _lib.ptr(eps), self.seed,
                                          self._policy_steps, self._rank * rows, _lib.ptr(action), _lib.ptr(logprob),
                                          _lib.ptr(env_action), None, self._stream()), "policy_step")
        self._policy_steps += 1
        return action, logprob, env_action

    def _explore_vec_env(self, env, horizon_len: int):
        """External tensor vec env: engine policy step + the env's own step(), per time step."""
        n, h, dev = self.num_envs, int(horizon_len), self.device
        states = th.empty((h, n, self.state_dim), dtype=th.float32, device=dev)
        actions = self._new_actions(h, n)
        logprobs = th.empty((h, n), dtype=th.float32, device=dev)
        rewards = th.empty((h, n), dtype=th.float32, device=dev)
        terminals = th.empty((h, n), dtype=th.bool, device=dev)
        truncates = th.empty((h, n), dtype=th.bool, device=dev)
        state = self.last_state.to(dev)
        noise = getattr(self, "_inject_eps", None)  # parity tests: [H, N, A] N(0,1) (Gaussian) / Exp(1) (categorical)
        for t in range(h):
            action, logprob, env_action = self._policy_step(state, None if noise is None else noise[t].contiguous())
            states[t], actions[t], logprobs[t] = state, action, logprob
            state, reward, terminal, truncate, _ = env.step(env_action)
            state = state.to(dev)
            rewards[t], terminals[t], truncates[t] = reward, terminal, truncate
        self.last_state = state
        rewards *= self.reward_scale
        self._value_cache = None
        return states, actions, logprobs, rewards, th.logical_not(terminals), th.logical_not(truncates)

    def _explore_vec_env_graphed(self, env, horizon_len: int):
        """``_explore_vec_env`` with the whole H-step loop -- policy-step kernels, the env's own torch ops, the stores --
        captured ONCE in a CUDA graph and replayed every cycle (opt-in: ``agent.cuda_graph_rollout = True``).  An eager
        torch vec env costs dozens of launches per step; the replay removes the launch and Python overhead.
        Requirements on the env: tensors on this device, ``step`` free of host synchronisation (no ``.item()``, no
        data-dependent shapes), state kept in tensor attributes of the env object (attributes that ``step`` rebinds are
        copied back into their original storage at the end of the graph), randomness from ``torch.Generator`` attributes
        (registered with the graph so that replays draw fresh numbers).  The policy's Philox step counter lives in device
        memory and is advanced inside the graph.  The returned tensors are STATIC: the next call overwrites them (the
        reference's loops consume a buffer before they roll out again, run.py:104-130)."""
        lib = _lib.load()
        n, h, dev = self.num_envs, int(horizon_len), self.device
        key = (id(env), h, n)
        rec = self._rollout_graphs.get(key)
        if rec is None:
            rec = dict(state=self.last_state.to(dev, th.float32).clone(), step_base=th.zeros(1, dtype=th.int64, device=dev),
                       states=th.empty((h, n, self.state_dim), dtype=th.float32, device=dev), actions=self._new_actions(h, n),
                       logprobs=th.empty((h, n), dtype=th.float32, device=dev), rewards=th.empty((h, n), dtype=th.float32, device=dev),
                       terminals=th.empty((h, n), dtype=th.bool, device=dev), truncates=th.empty((h, n), dtype=th.bool, device=dev),
                       undones=th.empty((h, n), dtype=th.bool, device=dev), unmasks=th.empty((h, n), dtype=th.bool, device=dev))
            graph = th.cuda.CUDAGraph()
            for value in vars(env).values():
                if isinstance(value, th.Generator) and value.device.type == "cuda":
                    graph.register_generator_state(value)
            before = {k: v for k, v in vars(env).items() if th.is_tensor(v) and v.is_cuda}
            lib.b200rl_set_policy_step_base(rec["step_base"].data_ptr())
            try:
                with th.cuda.graph(graph):
                    state = rec["state"]
                    for t in range(h):
                        action, logprob, env_action = self._policy_step(state)
                        rec["states"][t].copy_(state)
                        rec["actions"][t].copy_(action)
                        rec["logprobs"][t].copy_(logprob)
                        state, reward, terminal, truncate, _ = env.step(env_action)
                        rec["rewards"][t].copy_(reward)
                        rec["terminals"][t].copy_(terminal)
                        rec["truncates"][t].copy_(truncate)
                    rec["rewards"].mul_(self.reward_scale)
                    th.logical_not(rec["terminals"], out=rec["undones"])
                    th.logical_not(rec["truncates"], out=rec["unmasks"])
                    rec["state"].copy_(state)
                    for name, old in before.items():  # state the env rebound to fresh tensors goes back to its storage
                        new = getattr(env, name)
                        if new is not old and th.is_tensor(new) and new.shape == old.shape and new.dtype == old.dtype:
                            old.copy_(new)
                            setattr(env, name, old)
                    rec["step_base"].add_(h)
            finally:
                lib.b200rl_set_policy_step_base(None)
            rec["graph"] = graph
            self._rollout_graphs[key] = rec
        elif self.last_state is not rec["state"]:
            rec["state"].copy_(self.last_state)  # the caller reset the env / replaced the observation
        rec["graph"].replay()
        self.last_state = rec["state"]
        self._value_cache = None
        return rec["states"], rec["actions"], rec["logprobs"], rec["rewards"], rec["undones"], rec["unmasks"]

    def _new_actions(self, h: int, n: int) -> TEN:
        """fp32 [H, N, A] raw actions, or int32 [H, N] indices for the discrete agent (reference AgentPPO.py:102-104)."""
        if self._categorical:
            return th.empty((h, n), dtype=th.int32, device=self.device)
        return th.empty((h, n, self.action_dim), dtype=th.float32, device=self.device)

    def _explore_one_env(self, env, horizon_len: int):
        """Single gym-style env with numpy I/O (reference AgentPPO.py:34-85); outputs shaped [H, 1, ...]."""
        h, dev = int(horizon_len), self.device
        states = th.empty((h, 1, self.state_dim), dtype=th.float32, device=dev)
        actions = self._new_actions(h, 1)
        logprobs = th.empty((h, 1), dtype=th.float32, device=dev)
        rewards = th.zeros((h, 1), dtype=th.float32)
        terminals = th.zeros((h, 1), dtype=th.bool)
        truncates = th.zeros((h, 1), dtype=th.bool)
        state = self.last_state.to(dev)
        noise = getattr(self, "_inject_eps", None)  # parity tests: [H, 1, A]
        for t in range(h):
            action, logprob, env_action = self._policy_step(state, None if noise is None else noise[t].contiguous())
            states[t], actions[t], logprobs[t] = state, action, logprob
            ary_state, reward, terminal, truncate, _ = env.step(env_action[0].cpu().numpy())
            if terminal or truncate:
                ary_state, _ = env.reset()
            state = th.as_tensor(np.asarray(ary_state), dtype=th.float32, device=dev).reshape(1, self.state_dim)
            rewards[t], terminals[t], truncates[t] = float(reward), bool(terminal), bool(truncate)
        self.last_state = state
        self._value_cache = None
        return (states, actions, logprobs, (rewards * self.reward_scale).to(dev),
                th.logical_not(terminals).to(dev), th.logical_not(truncates).to(dev))

    # -------------------------------------------------------------------------------------- update
    @_on_device
    def get_values(self, states: TEN) -> TEN:
        """critic(states) for [..., S] -> [...] (reference update_net values pass, AgentPPO.py:141-143)."""
        lib = self._require_engine()
        flat = self._check(states.reshape(-1, self.state_dim), "states")
        out = th.empty((flat.shape[0],), dtype=th.float32, device=self.device)
        cri_desc = self._net_desc(self.cri)
        _lib.check(lib.b200rl_mlp_forward(C.byref(cri_desc), _lib.ptr(flat), flat.shape[0], _lib.ptr(out), 0, self._stream()),
                   "mlp_forward")
        return out.reshape(states.shape[:-1])

    @_on_device
    def get_advantages(self, states: TEN, rewards: TEN, undones: TEN, unmasks: TEN, values: TEN,
                       last_value: Optional[TEN] = None):
        """Reference AgentPPO.get_advantages (AgentPPO.py:207-232); mutates rewards / undones in place like it.
        Returns (advantages, reward_sums, stat_sums) -- the reduction inputs of the normalisation come for free."""
        lib = self._require_engine()
        h, n = rewards.shape
        if last_value is None:
            last_value = self.get_values(self.last_state.to(self.device))
        advantages = th.empty_like(values)
        reward_sums = th.empty_like(values)
        stat_sums = th.empty(4, dtype=th.float64, device=self.device)
        assert undones.dtype == th.bool and unmasks.dtype == th.bool
        _lib.check(lib.b200rl_gae(_lib.ptr(self._check(rewards, "rewards")), _lib.ptr(undones), _lib.ptr(unmasks),
                                  _lib.ptr(self._check(values, "values")), _lib.ptr(self._check(last_value, "last_value")),
                                  h, n, float(self.gamma), float(self.lambda_gae_adv), int(bool(self.if_use_v_trace)),
                                  self._rank * n, _lib.ptr(advantages), _lib.ptr(reward_sums), _lib.ptr(stat_sums),
                                  self._stream()), "gae")
        return advantages, reward_sums, stat_sums

    def update_net(self, buffer) -> Tuple[float, float, float]:
        """Reference AgentPPO.update_net (AgentPPO.py:135-171): returns (obj_critic, obj_actor, obj_entropy)
        averaged over ``update_times = int(H * repeat_times / batch_size)`` minibatch updates."""
        # The three scalars are the one device -> host read of the cycle.  The update kernel stores them straight into pinned
        # host memory (unified addressing: a cudaHostAlloc'ed block is device-accessible at the same address), so the host
        # only waits for the stream -- no copy-engine round trip behind the last kernel.
        self._require_engine()
        host = self._host_result if self.pinned_result else None
        if host is None and self.pinned_result:
            host = self._host_result = th.empty(3, dtype=th.float32).pin_memory()
        if host is not None:
            self.update_net_device(buffer, _out=host)
            th.cuda.current_stream(self.device).synchronize()
            obj_critic, obj_actor, obj_entropy = host.tolist()
        else:
            obj_critic, obj_actor, obj_entropy = self.update_net_device(buffer).tolist()
        if obj_critic != obj_critic and getattr(self, "_px", None) is not None:   # NaN: did a peer fail to answer?
            off = int(_lib.load().b200rl_workspace_error_offset())
            if int(self._workspace[off:off + 4].view(th.int32).item()) != 0:
                raise _lib.B200RLError("env-sharded update: a peer rank did not join the gradient exchange within 2 s")
        return obj_critic, obj_actor, obj_entropy

    @_on_device
    def update_net_device(self, buffer, _out: Optional[TEN] = None) -> TEN:
        """``update_net`` without the host synchronisation: the three scalars stay in a device tensor (or go to ``_out``,
        three floats of device-accessible memory)."""
        lib = self._require_engine()
        states, actions, logprobs, rewards, undones, unmasks = buffer
        h, n = states.shape[0], states.shape[1]
        dev = self.device
        states = self._check(states, "states")
        if self._categorical:
            actions = actions.reshape(h, n)
            assert actions.dtype == th.int32 and actions.is_contiguous() and actions.is_cuda, "actions: int32 indices [H, N]"
        else:
            actions = self._check(actions.reshape(h, n, self.action_dim), "actions")

        # values: reuse what the fused rollout already computed with this very critic
        cache = self._value_cache
        if cache is not None and cache[0] == self._cache_key(states):
            values, last_value = cache[1], cache[2]
        else:
            values = self.get_values(states)
            last_value = None
        self._value_cache = None
        advantages, reward_sums, stat_sums = self.get_advantages(states, rewards, undones, unmasks, values, last_value)

        n_global = n * self._world
        count_lattice = 0 if self._full_std else ((h + 3) // 4) * ((n_global + 3) // 4)
        update_times = int(h * self.repeat_times / self.batch_size)
        assert update_times >= 1
        act_desc, cri_desc = self._net_desc(self.act), self._net_desc(self.cri)
        act_adam, cri_adam = self._adam_desc(self.act_optimizer, self.act), self._adam_desc(self.cri_optimizer, self.cri)
        workspace = self._get_workspace(act_desc, cri_desc)
        hp = _lib.PPOHyper(ratio_clip=float(self.ratio_clip), lambda_entropy=float(self.lambda_entropy),
                           clip_grad_norm=float(self.clip_grad_norm or 0.0), flags=int(self._ppo_flags))
        stats = th.empty(4, dtype=th.float32, device=dev)
        peer = self._world > 1 and self._peer_mode(lib, act_desc, cri_desc, hp, update_times)
        if not peer:
            if self._world > 1:
                import torch.distributed as dist
                dist.all_reduce(stat_sums, group=self._dist_group)
            _lib.check(lib.b200rl_adv_stats(_lib.ptr(stat_sums), h * n_global, count_lattice, _lib.ptr(stats), self._stream()),
                       "adv_stats")
        tb = _lib.TrainBuffer(states=_lib.ptr(states), actions=_lib.ptr(actions), unmasks=_lib.ptr(unmasks),
                              logprobs=_lib.ptr(self._check(logprobs, "logprobs")), advantages=_lib.ptr(advantages),
                              reward_sums=_lib.ptr(reward_sums), adv_stats=_lib.ptr(stats), horizon_len=h, num_envs=n,
                              discrete_actions=int(self._categorical))
        out = th.empty(3, dtype=th.float32, device=dev) if _out is None else _out
        ids = getattr(self, "_inject_ids", None)
        if self._world == 1:
            _lib.check(lib.b200rl_ppo_update(C.byref(act_desc), C.byref(cri_desc), C.byref(act_adam), C.byref(cri_adam),
                                             C.byref(tb), C.byref(hp), self.batch_size, update_times, _lib.ptr(ids),
                                             self.seed, self._update_draws, _lib.ptr(out), _lib.ptr(workspace),
                                             workspace.numel(), self._stream()), "ppo_update")
        elif peer:
            # zero host-issued collectives: statistics and gradients are exchanged inside the update kernel (NVLink)
            px = self._px[0]
            seed = self.seed + 0x9E3779B9 * (self._rank + 1)  # independent index streams per shard
            _lib.check(lib.b200rl_ppo_update_sharded(C.byref(act_desc), C.byref(cri_desc), C.byref(act_adam), C.byref(cri_adam),
                                                     C.byref(tb), C.byref(hp), self.batch_size, update_times, _lib.ptr(ids), seed,
                                                     self._update_draws, _lib.ptr(stat_sums), h * n_global, count_lattice,
                                                     _lib.ptr(stats), _lib.ptr(out), _lib.ptr(workspace), workspace.numel(),
                                                     C.byref(px), 0 if getattr(self, "sharded_mode", "auto") == "peer_allreduce" else 1,
                                                     self._stream()), "ppo_update_sharded")
            px.epoch += update_times
            px.reserved += 1
        else:
            self._update_sharded(lib, act_desc, cri_desc, act_adam, cri_adam, tb, hp, update_times, ids, out, workspace)
        self._update_draws += update_times
        # the kernels wrote the parameters behind autograd's back: bump the in-place version counters (value-cache key,
        # and anything else of torch's that watches them)
        th.autograd.graph.increment_version(list(self.act.parameters()) + list(self.cri.parameters()))
        self._set_adam_step(self.act_optimizer, self.act, act_adam.step)
        self._set_adam_step(self.cri_optimizer, self.cri, cri_adam.step)
        self.last_update_info = dict(update_times=update_times, advantages=advantages, reward_sums=reward_sums,
                                     values=values, adv_stats=stats)
        return out

    def _update_sharded(self, lib, act_desc, cri_desc, act_adam, cri_adam, tb, hp, update_times, ids, out, workspace):
        """Env-sharded update.  The minibatch indices do not depend on the parameters, so every rank packs ITS share
        of ALL the minibatches of this update_net (update_times x batch_size/world sampled transitions, ~30 KB) in
        one kernel, ONE all-gather exchanges them, and every rank then runs the same single-launch update on the
        same global minibatches; a broadcast of the (36 KB) parameters + moments from rank 0 keeps the replicas
        bit-identical (the gradient reduction uses floating-point RED.ADD, whose order is not deterministic).
        2 latency-bound collectives per cycle instead of one all-reduce per minibatch."""
        import torch.distributed as dist
        assert self.batch_size % self._world == 0, "batch_size must divide over the ranks"
        local_batch = self.batch_size // self._world
        if getattr(self, "sharded_mode", "gather") == "gather":
            rec = ((self.state_dim + self.action_dim + 3) & ~3) + 4
            send = th.empty((update_times * local_batch, rec), dtype=th.float32, device=self.device)
            seed = self.seed + 0x9E3779B9 * (self._rank + 1)  # independent index streams per shard
            _lib.check(lib.b200rl_pack_minibatches(C.byref(tb), self.state_dim, self.action_dim, local_batch, update_times,
                                                   _lib.ptr(ids), seed, self._update_draws, _lib.ptr(send), self._stream()),
                       "pack_minibatches")
            recv = th.empty((self._world * update_times * local_batch, rec), dtype=th.float32, device=self.device)
            dist.all_gather_into_tensor(recv, send, group=self._dist_group)
            key = (self._world, update_times, local_batch)
            if getattr(self, "_packed_ids_key", None) != key:  # minibatch u = records (r, u, :) of every rank r
                r = th.arange(self._world, device=self.device).view(1, -1, 1) * (update_times * local_batch)
                u = th.arange(update_times, device=self.device).view(-1, 1, 1) * local_batch
                j = th.arange(local_batch, device=self.device).view(1, 1, -1)
                self._packed_ids = (r + u + j).reshape(update_times, -1).contiguous()
                self._packed_ids_key = key
            packed = _lib.TrainBuffer(states=_lib.ptr(recv), horizon_len=0, num_envs=recv.shape[0],
                                      discrete_actions=int(self._categorical))
            _lib.check(lib.b200rl_ppo_update(C.byref(act_desc), C.byref(cri_desc), C.byref(act_adam), C.byref(cri_adam),
                                             C.byref(packed), C.byref(hp), self.batch_size, update_times,
                                             _lib.ptr(self._packed_ids), self.seed, self._update_draws, _lib.ptr(out),
                                             _lib.ptr(workspace), workspace.numel(), self._stream()), "ppo_update")
            tensors = [p.data for p in _trainable(self.act) + _trainable(self.cri)]
            for opt, module in ((self.act_optimizer, self.act), (self.cri_optimizer, self.cri)):
                for p in _trainable(module):
                    tensors += [opt.state[p]["exp_avg"], opt.state[p]["exp_avg_sq"]]
            flat = th.cat([t.reshape(-1) for t in tensors])
            dist.broadcast(flat, src=dist.get_global_rank(self._dist_group, 0), group=self._dist_group)
            th._foreach_copy_(tensors, [c.view_as(t) for c, t in zip(flat.split([t.numel() for t in tensors]), tensors)])
            return
        # ---- alternative: per-minibatch gradient all-reduce (b200rl_ppo_grads -> NCCL -> b200rl_ppo_apply)
        grad_numel = lib.b200rl_grad_numel(C.byref(act_desc), C.byref(cri_desc))
        grad_off = lib.b200rl_workspace_grad_offset()
        flat_grads = workspace[grad_off:grad_off + 4 * grad_numel].view(th.float32)
        loss_sums = th.zeros(4, dtype=th.float64, device=self.device)
        seed = self.seed + 0x9E3779B9 * (self._rank + 1)  # independent index streams per shard
        for u in range(update_times):
            ids_u = ids[u] if ids is not None else None
            _lib.check(lib.b200rl_ppo_grads(C.byref(act_desc), C.byref(cri_desc), C.byref(tb), C.byref(hp), local_batch,
                                            self.batch_size, _lib.ptr(ids_u), seed, self._update_draws + u,
                                            _lib.ptr(loss_sums), _lib.ptr(workspace), workspace.numel(), self._stream()),
                       "ppo_grads")
            dist.all_reduce(flat_grads, group=self._dist_group)
            _lib.check(lib.b200rl_ppo_apply(C.byref(act_desc), C.byref(cri_desc), C.byref(act_adam), C.byref(cri_adam),
                                            C.byref(hp), _lib.ptr(workspace), workspace.numel(), self._stream()), "ppo_apply")
        dist.all_reduce(loss_sums, group=self._dist_group)
        _lib.check(lib.b200rl_loss_means(_lib.ptr(loss_sums), update_times, _lib.ptr(out), self._stream()), "loss_means")

    # ------------------------------------------------------------------------------ checkpointing
    def save_or_load_agent(self, cwd: str, if_save: bool):
        """Whole-object ``th.save`` / ``th.load`` of act, cri and their optimizers, file names as the reference
        (AgentBase.py:280-297)."""
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
            # a separately pickled optimizer refers to its own parameter copies (reference quirk): re-attach its
            # state, by position, to an optimizer over the loaded modules so that training continues correctly
            for opt_name, module in (("act_optimizer", self.act), ("cri_optimizer", self.cri)):
                loaded = getattr(self, opt_name)
                fresh = th.optim.Adam(module.parameters(), self.learning_rate)
                for p_old, p_new in zip(loaded.param_groups[0]["params"], fresh.param_groups[0]["params"]):
                    if p_old in loaded.state and len(loaded.state[p_old]):
                        fresh.state[p_new] = {k: (v.to(p_new.device) if k != "step" and th.is_tensor(v) else v)
                                              for k, v in loaded.state[p_old].items()}
                setattr(self, opt_name, fresh)
        self._value_cache = None


class AgentDiscretePPO(AgentPPO):
    """Drop-in for the reference's ``AgentDiscretePPO`` (``elegantrl/agents/AgentPPO.py:252-270``): the same cycle with a
    Categorical policy (``ActorDiscretePPO`` :393-425).  ``explore_env`` returns ``actions`` as int32 ``[H, N]`` (:103-104)
    and hands ``action.long()`` to ``env.step`` (:423-425); ``lambda_entropy`` defaults to 0.01 (:263)."""
    _categorical = True

    def __init__(self, net_dims, state_dim: int, action_dim: int, gpu_id: int = 0, args=None):
        if args is None:
            from ..config import Config
            args = Config()
            args.if_discrete = True
        super().__init__(net_dims, state_dim, action_dim, gpu_id, args)
        self.lambda_entropy = getattr(args, "lambda_entropy", 0.01)


class AgentA2C(AgentPPO):
    """Drop-in for the reference's ``AgentA2C`` (``elegantrl/agents/AgentPPO.py:252-311``): PPO's rollout / GAE pass with the
    plain policy-gradient actor objective ``(advantage * new_logprob).mean()`` (no ratio, no clip, no entropy term, not
    masked) and ``update_net`` returning ``(obj_critic, obj_actor, 0)``.  The reference's ``update_objectives`` indexes the
    time axis only and is coherent for single-env buffers ``[H, 1, ...]``, so this class asserts
    ``num_envs == 1`` in ``update_net``; use ``AgentPPO`` for vec envs."""

    def __init__(self, net_dims, state_dim: int, action_dim: int, gpu_id: int = 0, args=None):
        super().__init__(net_dims, state_dim, action_dim, gpu_id, args)
        self._ppo_flags = _lib.PPO_A2C | _lib.PPO_ACTOR_UNMASKED
        self.pinned_result = False   # the third slot is patched on the device below (stream-ordered)

    def update_net_device(self, buffer, _out: Optional[TEN] = None) -> TEN:
        assert _out is None
        assert buffer[0].shape[1] == 1, "AgentA2C follows the reference: single-env buffers [H, 1, ...] only"
        lambda_entropy, self.lambda_entropy = self.lambda_entropy, 0.0  # the A2C objective has no entropy term
        try:
            out = super().update_net_device(buffer)
        finally:
            self.lambda_entropy = lambda_entropy
        out[2] = 0.0  # reference returns a literal 0 in the third slot (AgentPPO.py:292)
        return out


class AgentDiscreteA2C(AgentDiscretePPO):
    """The reference's ``AgentDiscreteA2C`` (``AgentPPO.py:330-342``) derives from ``AgentDiscretePPO`` and inherits its
    ``update_objectives`` unchanged: it IS discrete PPO under another name."""

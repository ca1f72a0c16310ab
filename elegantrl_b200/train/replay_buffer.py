"""Drop-in for the reference's ``ReplayBuffer`` (``elegantrl/train/replay_buffer.py:11-134``; no PER) on the B200 engine:
same constructor arguments, attributes (``p, if_full, cur_size, add_size, max_size, num_seqs, states, actions, rewards,
undones, unmasks``) and methods ``update(items)`` / ``sample(batch_size)``; the five rings are the same time-major
``[max_size, num_seqs, ...]`` float32 tensors.  ``update`` is one ring-write kernel (``b200rl_replay_append``) instead of ten
slice assignments; the pointer arithmetic is the reference's (:84-118).  ``AgentSAC.update_net`` does not call ``sample``:
its gathers are fused into the update kernels (``b200rl_sac_update``); ``sample`` exists for API parity."""
from typing import Tuple

import torch as th

from .. import _lib

TEN = th.Tensor


class ReplayBuffer:
    def __init__(self, max_size: int, state_dim: int, action_dim: int, gpu_id: int = 0, num_seqs: int = 1,
                 if_use_per: bool = False, if_discrete: bool = False, args=None):
        assert not if_use_per, "prioritised replay is out of scope of the B200 engine (SURVEY section 2)"
        assert not if_discrete, "the engine's off-policy path is the continuous-action one (AgentSAC)"
        self.p, self.if_full, self.cur_size, self.add_size = 0, False, 0, 0
        self.max_size, self.num_seqs = int(max_size), int(num_seqs)
        self.state_dim, self.action_dim = int(state_dim), int(action_dim)
        self.device = th.device(f"cuda:{gpu_id}" if (th.cuda.is_available() and gpu_id >= 0) else "cpu")
        f32 = dict(dtype=th.float32, device=self.device)
        self.states = th.empty((max_size, num_seqs, state_dim), **f32)
        self.actions = th.empty((max_size, num_seqs, action_dim), **f32)
        self.rewards = th.empty((max_size, num_seqs), **f32)
        self.undones = th.empty((max_size, num_seqs), **f32)
        self.unmasks = th.empty((max_size, num_seqs), **f32)
        self.ids0 = th.tensor((), dtype=th.long, device=self.device)
        self.ids1 = th.tensor((), dtype=th.long, device=self.device)
        self.if_use_per = False

    def descriptor(self) -> "_lib.ReplayBufferDesc":
        return _lib.ReplayBufferDesc(states=self.states.data_ptr(), actions=self.actions.data_ptr(), rewards=self.rewards.data_ptr(),
                                     undones=self.undones.data_ptr(), unmasks=self.unmasks.data_ptr(), max_size=self.max_size,
                                     num_seqs=self.num_seqs, state_dim=self.state_dim, action_dim=self.action_dim)

    def update(self, items: Tuple[TEN, ...]):
        """replay_buffer.py:78-118: append ``add_size`` time rows at the pointer, wrapping around."""
        import ctypes as C
        if self.device.type != "cuda":
            raise _lib.B200RLError("ReplayBuffer (B200 engine) needs a CUDA device; there is no CPU path")
        states, actions, rewards, undones, unmasks = items
        self.add_size = rewards.shape[0]
        assert self.add_size <= self.max_size
        if undones.dtype != th.bool:   # explore_env returns torch.bool masks; the rings hold them as float32 (reference :57-58)
            undones = undones != 0
        if unmasks.dtype != th.bool:
            unmasks = unmasks != 0
        tensors = [t.contiguous() for t in (states.to(th.float32), actions.to(th.float32), rewards.to(th.float32), undones, unmasks)]
        desc = self.descriptor()
        with th.cuda.device(self.device):
            _lib.check(_lib.load().b200rl_replay_append(C.byref(desc), self.p, self.add_size, *(t.data_ptr() for t in tensors),
                                                        th.cuda.current_stream(self.device).cuda_stream), "replay_append")
        p = self.p + self.add_size
        if p > self.max_size:
            self.if_full = True
            p = p - self.max_size
        self.p = p
        self.cur_size = self.max_size if self.if_full else self.p

    def sample(self, batch_size: int) -> Tuple[TEN, TEN, TEN, TEN, TEN, TEN]:
        """replay_buffer.py:120-134 (torch indexing; the engine's update kernels gather by themselves)."""
        sample_len = self.cur_size - 1
        ids = th.randint(sample_len * self.num_seqs, size=(batch_size,), requires_grad=False, device=self.device)
        self.ids0 = ids0 = th.fmod(ids, sample_len)
        self.ids1 = ids1 = th.div(ids, sample_len, rounding_mode='floor')
        return (self.states[ids0, ids1], self.actions[ids0, ids1], self.rewards[ids0, ids1], self.undones[ids0, ids1],
                self.unmasks[ids0, ids1], self.states[ids0 + 1, ids1])

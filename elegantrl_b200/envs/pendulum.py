"""Pendulum-v1 as a tensor vec env obeying the ElegantRL vec-env contract.

The contract (what ``AgentPPO._explore_vec_env`` calls, reference
``elegantrl/agents/AgentPPO.py:119`` and ``elegantrl/train/config.py:243-247, 291-302``):

* ``reset() -> (state [N, S] float32, info_dict)``
* ``step(action [N, A] float32 in (-1, 1)) -> (state, reward [N], terminal [N] bool, truncate [N] bool, info)``
* sub-envs auto-reset inside ``step`` and return the post-reset observation
* attributes ``env_name, num_envs, max_step, state_dim, action_dim, if_discrete, device``

Physics is gymnasium's ``Pendulum-v1`` (g=10, m=l=1, dt=0.05, max_speed 8, max_torque 2) with the scaling of
ElegantRL's wrapper (reference ``elegantrl/envs/CustomGymEnv.py:39-44``): torque = 2*action, reward*0.5.
gymnasium is not vendored by the reference and is absent here, so this file *is* the definition of the env the
B200 engine's fused rollout kernel (``csrc/rollout.cu``) must reproduce op for op; ``step`` below is the plain
PyTorch statement of it and is what the reference agent is run against when golden vectors are minted.

The fused kernel reads and writes ``theta / theta_dot / cur_step`` in place (see ``engine_state``), so torch
``step`` calls and fused rollouts can be interleaved on the same env object.
"""
import math
from typing import Optional, Tuple

import torch as th

TEN = th.Tensor

GRAVITY = 10.0
MASS = 1.0
LENGTH = 1.0
DT = 0.05
MAX_SPEED = 8.0
MAX_TORQUE = 2.0
ACTION_SCALE = 2.0  # CustomGymEnv.py:42  env.step(action * 2)
REWARD_SCALE = 0.5  # CustomGymEnv.py:44  float(reward) * 0.5
PI = math.pi
TWO_PI = 2.0 * math.pi


class PendulumVecEnv:
    """N independent pendulums stepped as tensors. ``env_kind`` tells the engine a fused kernel exists."""
    env_kind = "pendulum"

    def __init__(self, num_envs: int = 8, gpu_id: int = -1, max_step: int = 200, seed: int = 0, **_kwargs):
        self.env_name = "Pendulum-v1"
        self.num_envs = int(num_envs)
        self.max_step = int(max_step)
        self.state_dim = 3
        self.action_dim = 1
        self.if_discrete = False
        self.device = th.device(f"cuda:{gpu_id}" if (th.cuda.is_available() and gpu_id >= 0) else "cpu")
        self.seed = int(seed)

        self.generator = th.Generator(device=self.device)
        self.generator.manual_seed(self.seed)
        # ONE [3, N] block (theta, theta_dot, cur_step as int32 bits): a host loop moves the whole env state with a single copy
        self._block = th.zeros((3, self.num_envs), dtype=th.float32, device=self.device)
        self._rows = (self._block[0], self._block[1], self._block[2].view(th.int32))
        self.theta, self.theta_dot, self.cur_step = self._rows
        self.global_step = 0  # counts step() calls + fused steps: Philox offset of the fused kernel
        self.reset_noise: Optional[TEN] = None  # injected U[0,1) noise [T, N, 2] for parity tests
        self._reset_noise_row = 0

    '''contract'''

    def reset(self, **_kwargs) -> Tuple[TEN, dict]:
        u = self._draw_uniform()
        self.theta[:] = (u[:, 0] * 2.0 - 1.0) * PI
        self.theta_dot[:] = u[:, 1] * 2.0 - 1.0
        self.cur_step.zero_()
        return self.get_state(), dict()

    def step(self, action: TEN) -> Tuple[TEN, TEN, TEN, TEN, dict]:
        theta, theta_dot = self.theta, self.theta_dot
        torque = (action.reshape(self.num_envs, -1)[:, 0].to(th.float32) * ACTION_SCALE).clamp(-MAX_TORQUE, MAX_TORQUE)

        theta_norm = th.remainder(theta + PI, TWO_PI) - PI
        cost = theta_norm * theta_norm + 0.1 * (theta_dot * theta_dot) + 0.001 * (torque * torque)
        reward = cost * (-REWARD_SCALE)

        accel = (3.0 * GRAVITY / (2.0 * LENGTH)) * th.sin(theta) + (3.0 / (MASS * LENGTH * LENGTH)) * torque
        new_theta_dot = (theta_dot + accel * DT).clamp(-MAX_SPEED, MAX_SPEED)
        new_theta = theta + new_theta_dot * DT

        self.cur_step += 1
        truncate = self.cur_step >= self.max_step
        terminal = th.zeros_like(truncate)

        u = self._draw_uniform()
        self.theta = th.where(truncate, (u[:, 0] * 2.0 - 1.0) * PI, new_theta)
        self.theta_dot = th.where(truncate, u[:, 1] * 2.0 - 1.0, new_theta_dot)
        self.cur_step = th.where(truncate, th.zeros_like(self.cur_step), self.cur_step)
        self.global_step += 1
        return self.get_state(), reward, terminal, truncate, dict()

    def close(self):
        pass

    '''helpers'''

    def get_state(self) -> TEN:
        return th.stack((th.cos(self.theta), th.sin(self.theta), self.theta_dot), dim=1)

    def _draw_uniform(self) -> TEN:
        if self.reset_noise is not None:
            u = self.reset_noise[self._reset_noise_row].to(self.device)
            self._reset_noise_row += 1
            return u
        return th.rand((self.num_envs, 2), dtype=th.float32, device=self.device, generator=self.generator)

    def inject_reset_noise(self, noise: Optional[TEN]):
        """noise[r] is consumed by the r-th call of reset()/step(); row 0 by the first call after injection."""
        self.reset_noise = noise
        self._reset_noise_row = 0

    def engine_state(self) -> Tuple[TEN, TEN, TEN]:
        """Tensors the fused rollout kernel updates in place: rows of the [3, N] state block.  ``step`` / a caller may have
        rebound the attributes to fresh tensors; those are copied back into the block first."""
        rows = self._rows   # the block's row views, created once: an identity test per call instead of three indexing ops
        if rows is None:
            blk = self._block
            rows = self._rows = (blk[0], blk[1], blk[2].view(th.int32))
        if self.theta is not rows[0]:
            rows[0].copy_(self.theta)
            self.theta = rows[0]
        if self.theta_dot is not rows[1]:
            rows[1].copy_(self.theta_dot)
            self.theta_dot = rows[1]
        if self.cur_step is not rows[2]:
            rows[2].copy_(self.cur_step)
            self.cur_step = rows[2]
        return self.theta, self.theta_dot, self.cur_step

    def engine_state_block(self) -> TEN:
        """The [3, N] fp32 block behind ``engine_state()`` (row 2 = ``cur_step`` int32 bit patterns): one H2D / D2H copy
        moves the whole env state."""
        self.engine_state()
        return self._block


class PendulumEnv:
    """ONE pendulum with the gym-style numpy contract of the reference's single-env path (what ``AgentPPO._explore_one_env``
    calls, reference ``elegantrl/agents/AgentPPO.py:63-70``; ``envs/CustomGymEnv.py:24-44``): ``reset() -> (state [S]
    ndarray, info)``, ``step(action [A] ndarray) -> (state, reward float, terminal bool, truncate bool, info)``, NO
    auto-reset -- the caller resets after a terminal / truncated step.  Same physics as ``PendulumVecEnv`` (it wraps one
    with ``num_envs = 1`` on the CPU; the vec env's internal auto-reset is undone by the caller's ``reset()``)."""

    def __init__(self, max_step: int = 200, seed: int = 0, **_kwargs):
        self.inner = PendulumVecEnv(num_envs=1, gpu_id=-1, max_step=max_step, seed=seed)
        self.env_name, self.num_envs, self.max_step = "Pendulum-v1", 1, int(max_step)
        self.state_dim, self.action_dim, self.if_discrete = 3, 1, False

    def reset(self, **_kwargs):
        return self.inner.reset()[0][0].numpy().copy(), dict()

    def step(self, action):
        state, reward, terminal, truncate, _ = self.inner.step(th.as_tensor(action, dtype=th.float32).reshape(1, 1))
        return state[0].numpy().copy(), float(reward[0]), bool(terminal[0]), bool(truncate[0]), dict()

    def close(self):
        pass

"""CartPole as a tensor vec env with a DISCRETE action space, obeying the ElegantRL vec-env contract
(see ``envs/pendulum.py`` for the contract; reference ``elegantrl/agents/AgentPPO.py:119``,
``elegantrl/train/config.py:243-247, 291-302``).

The reference's discrete demo (``examples/demo_A2C_PPO_discrete.py``) trains ``AgentDiscretePPO`` on gymnasium's
``CartPole-v1`` through the subprocess ``VecEnv``; gymnasium is not vendored by the reference and is absent here, so the
classic cart-pole dynamics (Barto, Sutton & Anderson; Euler integration, tau = 0.02, force 10 N, terminal beyond
+-2.4 m or +-12 degrees, reward 1 per step) are restated in torch.  Unlike Pendulum it produces real ``terminal``
flags, so it exercises the ``undones`` path of the GAE kernel through the external-env (per-step) engine path.
``step`` receives ``action [N]`` integer indices, exactly what ``ActorDiscretePPO.convert_action_for_env``
(reference ``AgentPPO.py:423-425``) passes.
"""
import math
from typing import Optional, Tuple

import torch as th

TEN = th.Tensor

GRAVITY = 9.8
MASS_CART = 1.0
MASS_POLE = 0.1
TOTAL_MASS = MASS_CART + MASS_POLE
HALF_LENGTH = 0.5
POLE_MASS_LENGTH = MASS_POLE * HALF_LENGTH
FORCE_MAG = 10.0
TAU = 0.02
THETA_LIMIT = 12 * 2 * math.pi / 360
X_LIMIT = 2.4


class CartPoleVecEnv:
    """N independent cart-poles stepped as tensors; no fused kernel exists for it (``env_kind`` is absent), so the
    agent drives it through ``b200rl_policy_step_discrete`` once per step."""

    def __init__(self, num_envs: int = 8, gpu_id: int = -1, max_step: int = 500, seed: int = 0, **_kwargs):
        self.env_name = "CartPole-v1"
        self.num_envs = int(num_envs)
        self.max_step = int(max_step)
        self.state_dim = 4
        self.action_dim = 2
        self.if_discrete = True
        self.device = th.device(f"cuda:{gpu_id}" if (th.cuda.is_available() and gpu_id >= 0) else "cpu")
        self.generator = th.Generator(device=self.device)
        self.generator.manual_seed(int(seed))
        self.state = th.zeros((self.num_envs, 4), dtype=th.float32, device=self.device)
        self.cur_step = th.zeros(self.num_envs, dtype=th.int32, device=self.device)
        self.reset_noise: Optional[TEN] = None  # injected U[0,1) noise [T, N, 4] for parity tests
        self._reset_noise_row = 0

    def reset(self, **_kwargs) -> Tuple[TEN, dict]:
        self.state = self._draw_uniform() * 0.1 - 0.05
        self.cur_step.zero_()
        return self.state.clone(), dict()

    def step(self, action: TEN) -> Tuple[TEN, TEN, TEN, TEN, dict]:
        x, x_dot, theta, theta_dot = self.state.unbind(dim=1)
        force = th.where(action.reshape(self.num_envs) > 0, FORCE_MAG, -FORCE_MAG).to(th.float32)
        cos_t, sin_t = th.cos(theta), th.sin(theta)
        temp = (force + POLE_MASS_LENGTH * (theta_dot * theta_dot) * sin_t) / TOTAL_MASS
        theta_acc = (GRAVITY * sin_t - cos_t * temp) / (HALF_LENGTH * (4.0 / 3.0 - MASS_POLE * (cos_t * cos_t) / TOTAL_MASS))
        x_acc = temp - POLE_MASS_LENGTH * theta_acc * cos_t / TOTAL_MASS
        x = x + TAU * x_dot
        x_dot = x_dot + TAU * x_acc
        theta = theta + TAU * theta_dot
        theta_dot = theta_dot + TAU * theta_acc
        new_state = th.stack((x, x_dot, theta, theta_dot), dim=1)

        self.cur_step += 1
        terminal = (x.abs() > X_LIMIT) | (theta.abs() > THETA_LIMIT)
        truncate = (self.cur_step >= self.max_step) & ~terminal
        reward = th.ones(self.num_envs, dtype=th.float32, device=self.device)
        done = terminal | truncate
        fresh = self._draw_uniform() * 0.1 - 0.05
        self.state = th.where(done[:, None], fresh, new_state)
        self.cur_step = th.where(done, th.zeros_like(self.cur_step), self.cur_step)
        return self.state.clone(), reward, terminal, truncate, dict()

    def close(self):
        pass

    def _draw_uniform(self) -> TEN:
        if self.reset_noise is not None:
            u = self.reset_noise[self._reset_noise_row].to(self.device)
            self._reset_noise_row += 1
            return u
        return th.rand((self.num_envs, 4), dtype=th.float32, device=self.device, generator=self.generator)

    def inject_reset_noise(self, noise: Optional[TEN]):
        """noise[r] is consumed by the r-th call of reset()/step(); row 0 by the first call after injection."""
        self.reset_noise = noise
        self._reset_noise_row = 0

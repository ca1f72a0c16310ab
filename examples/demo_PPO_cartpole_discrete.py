"""Discrete-action PPO on CartPole with the B200 engine -- the recipe of the reference's
``examples/demo_A2C_PPO_discrete.py:15-47`` (train_discrete_ppo_a2c_for_cartpole) with ``AgentDiscretePPO`` swapped for
the engine's and gymnasium's CartPole-v1 (absent here) replaced by the tensor ``CartPoleVecEnv``.  No fused rollout kernel
exists for this env: every step is one policy-step kernel around the env's own torch ops, and the whole H-step loop is
captured in a CUDA graph (``cuda_graph_rollout``).

    python examples/demo_PPO_cartpole_discrete.py [gpu_id] [num_envs]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from elegantrl_b200 import Config  # noqa: E402
from elegantrl_b200.agents import AgentDiscretePPO  # noqa: E402
from elegantrl_b200.envs import CartPoleVecEnv  # noqa: E402
from elegantrl_b200.train import train_agent  # noqa: E402


def main():
    gpu_id = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    num_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    env_args = {'env_name': 'CartPole-v1', 'max_step': 500, 'state_dim': 4, 'action_dim': 2, 'if_discrete': True,
                'num_envs': num_envs}
    args = Config(AgentDiscretePPO, CartPoleVecEnv, env_args)
    args.net_dims = [256, 128]
    args.gamma = 0.97
    args.reward_scale = 2 ** -2
    args.horizon_len = 64
    args.batch_size = 2048
    args.repeat_times = 256            # update_times = int(H * repeat_times / batch_size) = 8 minibatches of 2048
    args.learning_rate = 2e-4
    args.lambda_gae_adv = 0.75
    args.lambda_entropy = 0.0001
    args.gpu_id = gpu_id
    args.random_seed = 0
    args.cuda_graph_rollout = os.environ.get("CUDA_GRAPH", "1") == "1"
    train_agent(args, max_cycles=int(os.environ.get("CYCLES", 100)), eval_every=10)


if __name__ == "__main__":
    main()

"""PPO on Pendulum-v1 with the B200 engine -- the vec-env recipe of the reference's
``examples/demo_A2C_PPO.py:57-88`` (train_ppo_a2c_for_pendulum_vec_env) with the agent / env classes swapped.

    python examples/demo_PPO_pendulum_vec_env.py [gpu_id] [num_envs]

If the reference package is importable, ``USE_REFERENCE_LOOP=1`` runs the very same agent under the reference's own
``train_agent(args, if_single_process=True)`` instead of this repo's minimal loop.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from elegantrl_b200.agents import AgentPPO  # noqa: E402
from elegantrl_b200.envs import PendulumVecEnv  # noqa: E402


def main():
    gpu_id = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    num_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    env_args = {'env_name': 'Pendulum', 'max_step': 200, 'state_dim': 3, 'action_dim': 1, 'if_discrete': False,
                'num_envs': num_envs}
    if os.environ.get("USE_REFERENCE_LOOP"):
        from elegantrl import Config, train_agent            # the reference's own Config and loop
    else:
        from elegantrl_b200 import Config
        from elegantrl_b200.train import train_agent
    args = Config(AgentPPO, PendulumVecEnv, env_args)
    args.net_dims = [64, 64]
    args.gamma = 0.97
    args.reward_scale = 2 ** -2
    args.horizon_len = 64
    args.batch_size = 2048
    args.repeat_times = 256            # update_times = int(H * repeat_times / batch_size) = 8 minibatches of 2048
    args.learning_rate = 4e-4
    args.lambda_entropy = 0.001
    args.gpu_id = gpu_id
    args.random_seed = 0
    if os.environ.get("USE_REFERENCE_LOOP"):
        args.break_step = int(2e4)
        train_agent(args, if_single_process=True)
    else:
        train_agent(args, max_cycles=int(os.environ.get("CYCLES", 150)), eval_every=10)


if __name__ == "__main__":
    main()

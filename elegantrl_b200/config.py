"""Attribute bag with the hyper-parameter names the on-policy agent reads.

Mirror of the fields of the reference ``Config`` (``elegantrl/train/config.py:10-83``) that
``AgentBase.__init__`` / ``AgentPPO.__init__`` consume (``elegantrl/agents/AgentBase.py:27-68``,
``elegantrl/agents/AgentPPO.py:18-32``).  The B200 agent is duck-typed on ``args``: the reference's own
``Config`` object works unchanged; this class exists so the repo runs where the reference is not installed
(the GPU box) and carries only what the hot path needs -- no env factory, no evaluator options.
"""
from typing import Optional


class Config:
    def __init__(self, agent_class=None, env_class=None, env_args: Optional[dict] = None):
        self.agent_class = agent_class
        self.env_class = env_class
        if env_args is None:
            env_args = {'env_name': None, 'num_envs': 1, 'max_step': 12345,
                        'state_dim': None, 'action_dim': None, 'if_discrete': None}
        env_args.setdefault('num_envs', 1)
        env_args.setdefault('max_step', 12345)
        self.env_args = env_args
        self.env_name = env_args['env_name']
        self.num_envs = env_args['num_envs']
        self.max_step = env_args['max_step']
        self.state_dim = env_args['state_dim']
        self.action_dim = env_args['action_dim']
        self.if_discrete = env_args['if_discrete']
        self.if_off_policy = False  # this package only carries the on-policy path

        # reward shaping
        self.gamma = 0.99
        self.reward_scale = 2 ** 0
        # training (on-policy defaults, config.py:52-58)
        self.net_dims = [128, 128]
        self.learning_rate = 6e-5
        self.clip_grad_norm = 3.0
        self.state_value_tau = 0
        self.soft_update_tau = 5e-3
        self.continue_train = False
        self.batch_size = 128
        self.horizon_len = 2048
        self.buffer_size = None
        self.repeat_times = 8.0
        self.if_use_vtrace = True  # sic: the agent reads `if_use_v_trace`
        self.buffer_init_size = None
        # device
        self.gpu_id = 0
        self.num_workers = 2
        self.num_threads = 8
        self.random_seed = None
        self.learner_gpu_ids = ()
        # evaluation / bookkeeping fields the training loop reads
        self.cwd = None
        self.if_remove = True
        self.break_step = float('inf')
        self.break_score = float('inf')
        self.eval_times = 3
        self.eval_per_step = int(2e4)

"""ctypes binding of ``libb200rl.so`` (C ABI: ``include/b200rl.h``).

There is no CPU fallback: if the shared library is missing, ``load()`` raises.  The library is built in-tree
by ``__graft_entry__.build()`` / ``python -m elegantrl_b200._build``.
"""
import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200rl.so")

MAX_LINEAR = 8
ACT_GELU, ACT_RELU = 0, 1
ACTIVATION_CODES = {"gelu": ACT_GELU, "relu": ACT_RELU}

c_float_p = C.POINTER(C.c_float)


class Net(C.Structure):
    _fields_ = [("num_linear", C.c_int32), ("activation", C.c_int32), ("dims", C.c_int32 * (MAX_LINEAR + 1)),
                ("reserved", C.c_int32), ("weight", C.c_void_p * MAX_LINEAR), ("bias", C.c_void_p * MAX_LINEAR),
                ("state_avg", C.c_void_p), ("state_std", C.c_void_p), ("action_std_log", C.c_void_p)]


class Adam(C.Structure):
    _fields_ = [("exp_avg_w", C.c_void_p * MAX_LINEAR), ("exp_avg_b", C.c_void_p * MAX_LINEAR),
                ("exp_avg_sq_w", C.c_void_p * MAX_LINEAR), ("exp_avg_sq_b", C.c_void_p * MAX_LINEAR),
                ("exp_avg_std", C.c_void_p), ("exp_avg_sq_std", C.c_void_p),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("step", C.c_int64)]


class PPOHyper(C.Structure):
    _fields_ = [("ratio_clip", C.c_float), ("lambda_entropy", C.c_float), ("clip_grad_norm", C.c_float),
                ("flags", C.c_int32)]


PPO_SMOOTH_L1, PPO_MIN_CLIP, PPO_ENTROPY_BONUS, PPO_ACTOR_UNMASKED, PPO_CRITIC_MASK_MEAN = 1, 2, 4, 8, 16
PPO_HELLOWORLD = PPO_SMOOTH_L1 | PPO_MIN_CLIP | PPO_ENTROPY_BONUS | PPO_ACTOR_UNMASKED | PPO_CRITIC_MASK_MEAN
PPO_A2C = 32  # AgentA2C's actor objective (no ratio / clip); used together with PPO_ACTOR_UNMASKED and lambda_entropy = 0


class TrainBuffer(C.Structure):
    _fields_ = [("states", C.c_void_p), ("actions", C.c_void_p), ("unmasks", C.c_void_p), ("logprobs", C.c_void_p),
                ("advantages", C.c_void_p), ("reward_sums", C.c_void_p), ("adv_stats", C.c_void_p),
                ("horizon_len", C.c_int32), ("num_envs", C.c_int32), ("discrete_actions", C.c_int32), ("reserved", C.c_int32)]


MAX_PEERS, PX_FLAGS = 8, 64


class PeerExchange(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("data", C.c_void_p * MAX_PEERS), ("flags", C.c_void_p * MAX_PEERS),
                ("epoch", C.c_uint32), ("reserved", C.c_uint32)]


SAC_MAX_ENSEMBLES, MAX_GROUP_TENSORS = 8, 64


class SacActor(C.Structure):
    _fields_ = [("net_s", Net), ("net_a", Net)]


class SacCritic(C.Structure):
    _fields_ = [("encoder", Net), ("num_ensembles", C.c_int32), ("reserved", C.c_int32), ("decoder", Net * SAC_MAX_ENSEMBLES)]


class ParamGroup(C.Structure):
    _fields_ = [("num_tensors", C.c_int32), ("reserved", C.c_int32), ("param", C.c_void_p * MAX_GROUP_TENSORS),
                ("exp_avg", C.c_void_p * MAX_GROUP_TENSORS), ("exp_avg_sq", C.c_void_p * MAX_GROUP_TENSORS),
                ("numel", C.c_int32 * MAX_GROUP_TENSORS), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("step", C.c_int64)]


class ReplayBufferDesc(C.Structure):
    _fields_ = [("states", C.c_void_p), ("actions", C.c_void_p), ("rewards", C.c_void_p), ("undones", C.c_void_p),
                ("unmasks", C.c_void_p), ("max_size", C.c_int32), ("num_seqs", C.c_int32), ("state_dim", C.c_int32),
                ("action_dim", C.c_int32)]


class SacHyper(C.Structure):
    _fields_ = [("gamma", C.c_float), ("soft_update_tau", C.c_float), ("clip_grad_norm", C.c_float), ("target_entropy", C.c_float)]


class RolloutArgs(C.Structure):
    _fields_ = [("actor", C.POINTER(Net)), ("critic", C.POINTER(Net)),
                ("num_envs", C.c_int32), ("horizon_len", C.c_int32), ("max_step", C.c_int32),
                ("reward_scale", C.c_float),
                ("theta", C.c_void_p), ("theta_dot", C.c_void_p), ("cur_step", C.c_void_p),
                ("states", C.c_void_p), ("actions", C.c_void_p), ("logprobs", C.c_void_p), ("rewards", C.c_void_p),
                ("undones", C.c_void_p), ("unmasks", C.c_void_p), ("values", C.c_void_p),
                ("last_state", C.c_void_p), ("last_value", C.c_void_p),
                ("eps", C.c_void_p), ("reset_noise", C.c_void_p),
                ("seed", C.c_uint64), ("step_offset", C.c_uint64), ("env_offset", C.c_int64),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


ROLLOUT_DETERMINISTIC = 1


# symbol -> (restype, argtypes); every symbol include/b200rl.h declares (tests/test_abi.py checks the two agree)
SIGNATURES = {
    "b200rl_version": (C.c_char_p, []),
    "b200rl_last_error": (C.c_char_p, []),
    "b200rl_launch_count": (C.c_int64, []),
    "b200rl_workspace_bytes": (C.c_int64, [C.POINTER(Net), C.POINTER(Net)]),
    "b200rl_workspace_grad_offset": (C.c_int64, []),
    "b200rl_grad_numel": (C.c_int64, [C.POINTER(Net), C.POINTER(Net)]),
    "b200rl_mlp_forward": (C.c_int, [C.POINTER(Net), C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "b200rl_policy_step": (C.c_int, [C.POINTER(Net), C.POINTER(Net), C.c_void_p, C.c_int64, C.c_void_p, C.c_uint64,
                                     C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200rl_policy_step_discrete": (C.c_int, [C.POINTER(Net), C.POINTER(Net), C.c_void_p, C.c_int64, C.c_void_p, C.c_uint64,
                                              C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200rl_set_policy_step_base": (None, [C.c_void_p]),
    "b200rl_rollout_pendulum": (C.c_int, [C.POINTER(RolloutArgs), C.c_void_p]),
    "b200rl_gae": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                             C.c_float, C.c_float, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200rl_adv_stats": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "b200rl_normalize_adv": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "b200rl_ppo_update": (C.c_int, [C.POINTER(Net), C.POINTER(Net), C.POINTER(Adam), C.POINTER(Adam),
                                    C.POINTER(TrainBuffer), C.POINTER(PPOHyper), C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "b200rl_pack_minibatches": (C.c_int, [C.POINTER(TrainBuffer), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                          C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "b200rl_ppo_grads": (C.c_int, [C.POINTER(Net), C.POINTER(Net), C.POINTER(TrainBuffer), C.POINTER(PPOHyper),
                                   C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_void_p]),
    "b200rl_ppo_apply": (C.c_int, [C.POINTER(Net), C.POINTER(Net), C.POINTER(Adam), C.POINTER(Adam),
                                   C.POINTER(PPOHyper), C.c_void_p, C.c_int64, C.c_void_p]),
    "b200rl_loss_means": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "b200rl_replay_append": (C.c_int, [C.POINTER(ReplayBufferDesc), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200rl_sac_policy_step": (C.c_int, [C.POINTER(SacActor), C.c_void_p, C.c_int64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64,
                                         C.c_void_p, C.c_void_p]),
    "b200rl_sac_workspace_bytes": (C.c_int64, [C.POINTER(SacActor), C.POINTER(SacCritic), C.c_int32]),
    "b200rl_sac_update": (C.c_int, [C.POINTER(SacActor), C.POINTER(SacCritic), C.POINTER(SacCritic), C.POINTER(ParamGroup),
                                    C.POINTER(ParamGroup), C.POINTER(ParamGroup), C.POINTER(ReplayBufferDesc), C.c_int32,
                                    C.POINTER(SacHyper), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                    C.c_uint64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "b200rl_workspace_error_offset": (C.c_int64, []),
    "b200rl_update_tc_supported": (C.c_int32, [C.POINTER(Net), C.POINTER(Net), C.POINTER(PPOHyper)]),
    "b200rl_peer_exchange_floats": (C.c_int64, [C.POINTER(Net), C.POINTER(Net), C.c_int32, C.c_int32]),
    "b200rl_ppo_update_sharded": (C.c_int, [C.POINTER(Net), C.POINTER(Net), C.POINTER(Adam), C.POINTER(Adam),
                                            C.POINTER(TrainBuffer), C.POINTER(PPOHyper), C.c_int32, C.c_int32, C.c_void_p,
                                            C.c_uint64, C.c_uint64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int64, C.POINTER(PeerExchange), C.c_int32, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


class B200RLError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libb200rl.so and bind every declared symbol.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200RLError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().b200rl_last_error().decode("utf-8", "replace")
        raise B200RLError(f"{what or 'libb200rl'} failed (rc={rc}): {msg}")


def ptr(t) -> Optional[int]:
    """Device pointer of a contiguous torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "libb200rl needs contiguous tensors"
    return t.data_ptr()

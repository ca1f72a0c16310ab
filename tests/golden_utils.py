"""Load the committed golden vectors (minted from the Python reference by oracle/make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SYNTH_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "synth_*.npz")))
ROLLOUT_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "rollout_*.npz")))


def load(name):
    with np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")) as z:
        return {k: z[k] for k in z.files}


def net_of(g, prefix, dtype=np.float32, activation="gelu"):
    n_layers = int(g[f"{prefix}.n_layers"])
    net = {"W": [g[f"{prefix}.W{i}"].astype(dtype).copy() for i in range(n_layers)],
           "b": [g[f"{prefix}.b{i}"].astype(dtype).copy() for i in range(n_layers)],
           "state_avg": g[f"{prefix}.state_avg"].astype(dtype).copy(),
           "state_std": g[f"{prefix}.state_std"].astype(dtype).copy(),
           "activation": activation}
    key = f"{prefix}.action_std_log"
    if key in g:
        net["action_std_log"] = g[key].astype(dtype).copy()
    return net


def hyper_of(g):
    return dict(gamma=float(g["hp.gamma"]), lambda_gae_adv=float(g["hp.lambda_gae_adv"]),
                ratio_clip=float(g["hp.ratio_clip"]), lambda_entropy=float(g["hp.lambda_entropy"]),
                clip_grad_norm=float(g["hp.clip_grad_norm"]), learning_rate=float(g["hp.learning_rate"]),
                reward_scale=float(g["hp.reward_scale"]), batch_size=int(g["hp.batch_size"]),
                repeat_times=float(g["hp.repeat_times"]), if_use_v_trace=bool(int(g["hp.if_use_v_trace"])))


def flat_params(net):
    out = [p for pair in zip(net["W"], net["b"]) for p in pair]
    if "action_std_log" in net:
        out.append(net["action_std_log"])
    return out


HELLOWORLD_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "helloworld_*.npz")))


def plain_net_of(g, prefix, dtype=np.float32):
    """helloworld nets: ReLU, no state_norm (reference helloworld/helloworld_PPO_single_file.py:172-212)."""
    n_layers = int(g[f"{prefix}.n_layers"])
    net = {"W": [g[f"{prefix}.W{i}"].astype(dtype).copy() for i in range(n_layers)],
           "b": [g[f"{prefix}.b{i}"].astype(dtype).copy() for i in range(n_layers)],
           "state_avg": None, "state_std": None, "activation": "relu"}
    if f"{prefix}.action_std_log" in g:
        net["action_std_log"] = g[f"{prefix}.action_std_log"].astype(dtype).copy()
    return net


def helloworld_hyper_of(g):
    return dict(gamma=float(g["hp.gamma"]), lambda_gae_adv=float(g["hp.lambda_gae_adv"]), ratio_clip=float(g["hp.ratio_clip"]),
                lambda_entropy=float(g["hp.lambda_entropy"]), learning_rate=float(g["hp.learning_rate"]),
                batch_size=int(g["hp.batch_size"]), repeat_times=float(g["hp.repeat_times"]))


DISCRETE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "discrete_*.npz")))
CARTPOLE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "cartpole_*.npz")))


def discrete_net_of(g, prefix, dtype=np.float32):
    """ActorDiscretePPO: the inherited action_std_log is dead weight (reference AgentPPO.py:393-397), drop it."""
    net = net_of(g, prefix, dtype)
    net.pop("action_std_log", None)
    return net
A2C_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "a2c_*.npz")))

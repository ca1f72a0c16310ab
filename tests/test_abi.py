"""CPU checks of the C-ABI boundary: the library loads (no GPU needed: cudart is linked statically), exports
every symbol include/b200rl.h declares, the ctypes mirror agrees with the C struct layouts, and argument
validation fails loudly (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from elegantrl_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "b200rl.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"B200RL_API[^;(]*?\b(b200rl_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/b200rl.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"
    assert lib.b200rl_version().startswith(b"b200rl")


def test_struct_layouts_match_c():
    """Compile a probe with gcc against the header and compare sizeof / offsetof with the ctypes mirror."""
    probe = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "b200rl.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu\n", sizeof(b200rl_net), sizeof(b200rl_adam), sizeof(b200rl_ppo_hyper),
               sizeof(b200rl_train_buffer), sizeof(b200rl_rollout_args), sizeof(b200rl_peer_exchange));
        printf("%zu %zu %zu %zu %zu %zu %zu\n", offsetof(b200rl_net, weight), offsetof(b200rl_net, action_std_log),
               offsetof(b200rl_adam, lr), offsetof(b200rl_adam, step), offsetof(b200rl_rollout_args, seed),
               offsetof(b200rl_rollout_args, flags), offsetof(b200rl_peer_exchange, epoch));
        return 0;
    }'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "probe.c"), os.path.join(d, "probe")
        open(src, "w").write(probe)
        subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out[:6]]
    offs = [int(x) for x in out[6:]]
    assert sizes == [C.sizeof(_lib.Net), C.sizeof(_lib.Adam), C.sizeof(_lib.PPOHyper), C.sizeof(_lib.TrainBuffer),
                     C.sizeof(_lib.RolloutArgs), C.sizeof(_lib.PeerExchange)]
    assert offs == [_lib.Net.weight.offset, _lib.Net.action_std_log.offset, _lib.Adam.lr.offset, _lib.Adam.step.offset,
                    _lib.RolloutArgs.seed.offset, _lib.RolloutArgs.flags.offset, _lib.PeerExchange.epoch.offset]


def test_argument_validation_reports_errors():
    lib = _lib.load()
    net = _lib.Net()  # num_linear = 0 -> invalid
    rc = lib.b200rl_mlp_forward(C.byref(net), None, 4, None, 0, None)
    assert rc != 0 and b"num_linear" in lib.b200rl_last_error()
    with pytest.raises(_lib.B200RLError):
        _lib.check(rc, "mlp_forward")
    assert lib.b200rl_workspace_bytes(None, None) == -1
    assert lib.b200rl_workspace_grad_offset() == 256


def test_agent_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from elegantrl_b200 import Config
    from elegantrl_b200.agents import AgentPPO
    agent = AgentPPO([64, 64], 3, 1, gpu_id=-1, args=Config())
    with pytest.raises(_lib.B200RLError):
        agent.update_net([None] * 6)
    with pytest.raises(_lib.B200RLError):
        agent.explore_env(None, 8)


def test_product_path_never_imports_oracle():
    """The shipped package must not reach into oracle/ (parity claims are void otherwise)."""
    pkg = os.path.join(REPO, "elegantrl_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports oracle"

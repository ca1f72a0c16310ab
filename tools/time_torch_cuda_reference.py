"""Diagnostic: the reference's own op sequence (oracle/cpu_port.py, pinned to the goldens) run as EAGER PyTorch on the
same B200 -- the "stronger baseline" of SURVEY.md 8(d).  Same workload as bench.py: 65 536 envs x 128 steps + update."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.cpu_port import time_cpu_cycles

r = time_cpu_cycles(65536, 128, (64, 64), warmup=2, cycles=5, threads=8, device="cuda:0")
print(json.dumps(dict(impl="reference op sequence, eager PyTorch CUDA", env_steps_per_sec=r["env_steps_per_sec"],
                      explore_ms=[round(1e3 * x, 2) for x in r["explore_s"]], update_ms=[round(1e3 * x, 2) for x in r["update_s"]])))

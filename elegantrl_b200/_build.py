"""Compile libb200rl.so (sm_100a) in-tree with nvcc.  Called by ``__graft_entry__.build()``.

The library is built next to this file (``elegantrl_b200/libb200rl.so``) so that it travels with the repo
snapshot to the GPU box; there is no JIT cache and no fallback: importing ``elegantrl_b200._lib`` without the
library raises.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libb200rl.so")
SOURCES = ["api.cu", "forward.cu", "forward_tc.cu", "gae.cu", "update.cu", "update_tc.cu", "sac.cu", "rollout.cu", "rollout_tc.cu", "rollout_ts.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libb200rl.so cannot be built")
    return nvcc


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "b200rl.h"))
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    extra = (["-Xptxas", "-v"] if verbose else []) + os.environ.get("B200RL_EXTRA_NVCC_FLAGS", "").split()
    sources = list(SOURCES)

    def compile_one(src):
        src_path = os.path.join(CSRC, src)
        obj = os.path.join(obj_dir, os.path.basename(src).replace(".cu", ".o"))
        if force or _stale(obj, [src_path] + headers):
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src_path, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
            if verbose:
                sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as pool:
        objs = list(pool.map(compile_one, sources))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

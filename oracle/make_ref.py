"""Recipe that places the UNMODIFIED reference package next to the oracle: ``oracle/_ref/elegantrl/``.
TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by the product package.

    python oracle/make_ref.py            # needs /root/reference (the build container); idempotent

The reference is pure Python (no native sources), so "building" it is a byte-for-byte copy of
``/root/reference/elegantrl/**/*.py``.  ``oracle/_ref/`` is git-ignored (reference sources never enter the history) but
NOT gpurun-ignored: it travels to the GPU box with the snapshot, where ``bench.py --impl reference`` and the
``cpu_baseline`` leg time the reference's own ``AgentPPO`` (``gpu_id=-1``) and the ``-m gpu`` drop-in test runs the
reference's own ``train_agent`` around the B200 agent.  ``MANIFEST.json`` records the sha256 of every copied file and of
its source, which is how a reader checks that nothing was edited on the way.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("ELEGANTRL_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make_ref(verbose=True):
    src_pkg = os.path.join(REFERENCE, "elegantrl")
    if not os.path.isdir(src_pkg):
        if verbose:
            print(f"| make_ref: {src_pkg} not found (GPU box?): keeping whatever is in {DEST}")
        return os.path.isdir(os.path.join(DEST, "elegantrl"))
    dst_pkg = os.path.join(DEST, "elegantrl")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    manifest = {}
    for root, _, files in os.walk(src_pkg):
        for name in sorted(files):
            if not name.endswith(".py"):
                continue
            src = os.path.join(root, name)
            rel = os.path.relpath(src, REFERENCE)
            dst = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            assert _sha(src) == _sha(dst)
            manifest[rel] = _sha(dst)
    head = ""
    try:
        import subprocess
        head = subprocess.run(["git", "-C", REFERENCE, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:  # noqa: BLE001
        pass
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REFERENCE, "commit": head, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"| make_ref: copied {len(manifest)} files of {src_pkg} -> {dst_pkg}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)

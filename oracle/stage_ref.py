"""ORACLE — TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Stages the reference's own hot-path module for the GPU box.

    python -m oracle.stage_ref            (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python, so "building" it is making its unmodified source importable where /root/reference
does not exist (the GPU box): the three files `import easywakeword.wakeword` needs are copied byte for byte into
oracle/_ref/easywakeword/ — a directory that is git-ignored (never part of the repo's history) but travels with
gpurun snapshots, exactly like a compiled oracle/_ref binary would.  A manifest with their sha256 is written beside
them and checked against tests/golden/MANIFEST.json (the hash of the file the goldens were generated from).

Consumers: bench.py `--impl reference` and the `cpu_baseline` leg (the CPU arm then reports kind "reference": the
reference's SoundBuffer / WordMatcher / WakeWord._detect_word under the fake clock of oracle/ref_harness.py instead
of the oracle port).  No product code and no test under -m gpu reads oracle/_ref.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
SRC_ROOT = "/root/reference"
DST_ROOT = os.path.join(HERE, "_ref")
FILES = ["easywakeword/__init__.py", "easywakeword/wakeword.py", "easywakeword/transcriber.py"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def staged() -> bool:
    return all(os.path.isfile(os.path.join(DST_ROOT, f)) for f in FILES)


def stage(src_root: str = SRC_ROOT) -> dict | None:
    """Copies FILES from src_root into oracle/_ref/.  Returns the manifest, or None when src_root is absent."""
    if not all(os.path.isfile(os.path.join(src_root, f)) for f in FILES):
        return None
    man = {"source": src_root, "files": {}}
    for f in FILES:
        dst = os.path.join(DST_ROOT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, f), dst)
        man["files"][f] = _sha(dst)
    gm = os.path.join(REPO, "tests", "golden", "MANIFEST.json")
    if os.path.exists(gm):
        want = json.load(open(gm)).get("reference_wakeword_py_sha256")
        man["matches_golden_manifest"] = (want == man["files"]["easywakeword/wakeword.py"]) if want else None
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as f:
        json.dump(man, f, indent=1)
    return man


if __name__ == "__main__":
    m = stage()
    print(json.dumps(m, indent=1) if m else f"{SRC_ROOT} not present: nothing staged")

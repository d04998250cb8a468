"""TEST INFRASTRUCTURE ONLY -- recipe that makes the UNMODIFIED reference travel to the GPU box.

    python -m oracle.build_ref            # in the build container (needs /root/reference)

The reference is pure Python (its arithmetic core is third-party NATTEN, see oracle/natten_stub.py),
so "building" it means copying its `src/` package, byte for byte, from where it lies under
/root/reference into `oracle/_ref/` and recording a SHA-256 manifest.  `oracle/_ref/` is listed in
.gitignore (the reference's sources never enter this repository's history) but NOT in .gpurunignore,
so the copy travels to the GPU box with the snapshot like our own built `.so` files.  There,
`oracle/reference_runner.py` imports it with the NATTEN stub installed, and `bench.py --impl
reference` / `cpu_baseline` time the reference's own modules on the host cores (`kind: "reference"`).

Only `__graft_entry__.build()` (when /root/reference is present) and a developer call this script.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("NAF_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")
#: what the hot path imports: the whole `src` package (src/model/__init__.py imports every model)
TREES = ["src"]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "src", "model", "naf.py"))


def build() -> str:
    """Copy the reference's package into oracle/_ref/ (idempotent); returns the destination."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_SRC}")
    manifest = {}
    for tree in TREES:
        src_root = os.path.join(REF_SRC, tree)
        for base, dirs, files in os.walk(src_root):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            for f in files:
                if not f.endswith(".py"):
                    continue
                src = os.path.join(base, f)
                rel = os.path.relpath(src, REF_SRC)
                dst = os.path.join(DEST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                with open(dst, "rb") as fh:
                    manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    return DEST


def verify() -> bool:
    """True if oracle/_ref/ exists and every file still has the recorded hash (unmodified copy)."""
    path = os.path.join(DEST, "MANIFEST.json")
    if not os.path.isfile(path):
        return False
    with open(path) as fh:
        files = json.load(fh)["files"]
    for rel, digest in files.items():
        p = os.path.join(DEST, rel)
        if not os.path.isfile(p):
            return False
        with open(p, "rb") as fh:
            if hashlib.sha256(fh.read()).hexdigest() != digest:
                return False
    return bool(files)


if __name__ == "__main__":
    print(build())

"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference modules of the SSD3D path, staged next to the oracle.

TEST / BASELINE INFRASTRUCTURE ONLY (never imported by ``mslesions3d_b200``).

The reference is pure Python (nothing to compile): the four modules its ``ssd3d`` imports at load time
(``ssd3d.py``, ``mobilenet.py``, ``utils.py``, ``base_network.py``) are copied byte for byte from the read-only
mount ``/root/reference/lesions3d`` into ``oracle/_ref/lesions3d/`` together with a SHA-256 manifest.
``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored, so the staged files travel to the GPU box, where ``/root/reference`` does not exist:
``bench.py --impl reference`` imports them through ``oracle/ref_shim.py`` and times the reference's own
``LSSD3D.forward`` + ``detect_objects`` (``kind: "reference"``).

    python -m oracle.make_ref          # also run by __graft_entry__.build() when /root/reference is mounted
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

SOURCE_DIR = "/root/reference/lesions3d"
FILES = ("ssd3d.py", "mobilenet.py", "utils.py", "base_network.py")
REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "lesions3d")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(source_dir: str = SOURCE_DIR, force: bool = False) -> str:
    """Copy the reference modules (unchanged) into oracle/_ref/lesions3d; returns that directory, or "" when
    the reference mount is absent (the GPU box: the files staged in the build container are used as they are)."""
    if not os.path.isfile(os.path.join(source_dir, "ssd3d.py")):
        return REF_DIR if os.path.isfile(os.path.join(REF_DIR, "ssd3d.py")) else ""
    os.makedirs(REF_DIR, exist_ok=True)
    manifest = {}
    for name in FILES:
        src, dst = os.path.join(source_dir, name), os.path.join(REF_DIR, name)
        digest = _sha(src)
        if force or not os.path.isfile(dst) or _sha(dst) != digest:
            shutil.copyfile(src, dst)
        manifest[name] = digest
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump({"source": source_dir, "sha256": manifest}, f, indent=1)
    return REF_DIR


def verify() -> bool:
    """True when every staged file still matches the manifest written at staging time."""
    try:
        man = json.load(open(os.path.join(REF_DIR, "MANIFEST.json")))["sha256"]
        return all(_sha(os.path.join(REF_DIR, k)) == v for k, v in man.items())
    except Exception:
        return False


if __name__ == "__main__":
    print(stage(force=True) or "reference mount absent and nothing staged")

"""ORACLE tooling (test infrastructure, NOT product code) -- place the UNMODIFIED reference implementation of the hot
path under oracle/_ref/ so that it can travel to the GPU box and be timed as the reference arm.

    python oracle/make_ref.py            # in the build container, where /root/reference exists

The reference is a script tree (no setup.py / pyproject.toml), so `pip install --target baseline/_ref` is not
applicable; its hot path, however, is 330 lines of pure Python + torch:

    model/__init__.py  model/posendf.py  model/network/{__init__,net_modules,net_utils}.py
    configs/{__init__,config}.py  configs/amass.yaml

They are copied byte for byte from /root/reference into oracle/_ref/ (git-ignored, so no reference source enters
the history; NOT gpurun-ignored, so the directory ships with the snapshot like a built .so).  model/posendf.py
imports `ipdb` at module top, which is not installed: a no-op stub oracle/_ref/ipdb.py is generated.
A MANIFEST.json records the sha256 of every file so that `load_reference()` can refuse a tampered copy.

bench.py (`--impl reference`, `cpu_baseline`) and tests/ are the only users (see oracle/__init__ docstring).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REF = os.environ.get("POSENDF_REFERENCE", "/root/reference")
FILES = [
    "model/__init__.py", "model/posendf.py", "model/network/__init__.py", "model/network/net_modules.py",
    "model/network/net_utils.py", "configs/__init__.py", "configs/config.py", "configs/amass.yaml",
]
IPDB_STUB = "def set_trace(*a, **k):\n    pass\n"


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make(force: bool = False) -> bool:
    """returns True if oracle/_ref is in place afterwards"""
    if not os.path.isdir(REF):
        return os.path.exists(os.path.join(DEST, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or _sha(src) != _sha(dst):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "ipdb.py"), "w") as f:
        f.write(IPDB_STUB)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "garvita-tiwari/PoseNDF (read-only copy of /root/reference)", "sha256": manifest}, f, indent=1)
    return True


def available() -> bool:
    return os.path.exists(os.path.join(DEST, "MANIFEST.json"))


def load_reference():
    """import the reference's own PoseNDF / gradient / load_config from oracle/_ref (verifying the manifest)."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    with open(os.path.join(DEST, "MANIFEST.json")) as f:
        manifest = json.load(f)["sha256"]
    for rel, h in manifest.items():
        if _sha(os.path.join(DEST, rel)) != h:
            raise RuntimeError(f"oracle/_ref/{rel} does not match its manifest")
    # our own compat/ shim exports the same top-level package names; the reference copy must win here
    for name in [m for m in sys.modules if m == "model" or m.startswith("model.") or m == "configs" or m.startswith("configs.")]:
        del sys.modules[name]
    sys.path.insert(0, DEST)
    try:
        from model.posendf import PoseNDF, gradient      # noqa
        from configs.config import load_config           # noqa
    finally:
        sys.path.remove(DEST)
    return PoseNDF, gradient, load_config


def amass_opt(device="cpu"):
    """configs/amass.yaml of the reference with the device overridden"""
    _, _, load_config = load_reference()
    opt = load_config(os.path.join(DEST, "configs", "amass.yaml"))
    opt["train"]["device"] = device
    return opt


if __name__ == "__main__":
    ok = make(force="--force" in sys.argv)
    print("oracle/_ref", "ready" if ok else "NOT available (no /root/reference here)")

"""Recipe for ``oracle/_ref/``: a verbatim, git-ignored copy of the reference's Python files for this path
(TEST / BASELINE INFRASTRUCTURE -- the product never reads it).

The reference is pure Python and cannot be "compiled"; what the GPU box needs in order to time the reference's OWN
implementation beside ours (``bench.py --impl reference``, ``cpu_baseline.kind == "reference"``) and to launch the
reference's own drivers on the swapped path (``examples/run_reference_driver.py``) is the files themselves.
``/root/reference`` exists only in the build container, so ``__graft_entry__.build()`` runs this recipe there and the
copies travel with the working tree (``oracle/_ref/`` is listed in ``.gitignore``, not in ``.gpurunignore``).
Nothing is edited: each copy's sha256 equals the source's, and ``MANIFEST.json`` records them.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = ["code/MMD.py", "code/fusion_net.py", "code/fusion_train.py", "code/fusion_test.py", "LICENSE"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(reference_root: str | None = None) -> str | None:
    """Copy FILES from the reference checkout to oracle/_ref/ (flat).  Returns the directory, or None when no
    reference checkout is present and no earlier copy exists."""
    root = reference_root or os.environ.get("EDRL_REFERENCE_ROOT", "/root/reference")
    if not os.path.isfile(os.path.join(root, "code", "MMD.py")):
        return OUT if os.path.isfile(os.path.join(OUT, "MMD.py")) else None
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for rel in FILES:
        src = os.path.join(root, rel)
        dst = os.path.join(OUT, os.path.basename(rel))
        if not os.path.isfile(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[os.path.basename(rel)] = {"source": rel, "sha256": _sha(dst)}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return OUT


def available() -> bool:
    return os.path.isfile(os.path.join(OUT, "MMD.py"))


def load_mmd():
    """The reference's MMD module from the copy, unmodified (torch only)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_edrl_ref_copy_MMD", os.path.join(OUT, "MMD.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build())

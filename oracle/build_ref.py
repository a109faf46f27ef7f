"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference modules of the hot path, staged so that they travel to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE.  ``/root/reference`` exists in the build container only; ``bench.py --impl reference``
and ``bench.py``'s ``cpu_baseline`` leg run on the GPU box's host cores.  ``__graft_entry__.build()`` calls ``stage()``
wherever the reference checkout is present: it copies the three files the timed path needs --

    unfolded_DLASSO.py   (DLASSO_unfolded, seq_hyperparam: the solver that is timed)
    gnn_dlasso_utils.py  (set_A, compute_loss: problem generator and loss of the reference drivers)
    gnn_data.py          (set_Data recipe)

-- byte for byte into ``oracle/_ref/`` together with a manifest of their SHA-256 digests.  ``oracle/_ref/`` is listed in
``.gitignore`` (reference sources never enter this repository's history) and NOT in ``.gpurunignore`` (it ships with the
snapshot, like the built ``.so``).  ``oracle/ref_harness.py`` imports from it when ``/root/reference`` is absent.  The
product package never imports anything from here.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("unfolded_DLASSO.py", "gnn_dlasso_utils.py", "gnn_data.py")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(reference_root: str = "/root/reference") -> bool:
    """Copy the reference files into ``oracle/_ref/``; returns False (and leaves ``_ref`` as it is) without a checkout."""
    if not os.path.isfile(os.path.join(reference_root, FILES[0])):
        return False
    os.makedirs(REF_DIR, exist_ok=True)
    manifest = {"source": reference_root, "files": {}}
    for name in FILES:
        src, dst = os.path.join(reference_root, name), os.path.join(REF_DIR, name)
        if not (os.path.isfile(dst) and _sha(dst) == _sha(src)):
            shutil.copyfile(src, dst)
        manifest["files"][name] = _sha(dst)
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return True


def staged() -> bool:
    """True when ``oracle/_ref/`` holds every file of the manifest with its recorded digest (i.e. unmodified)."""
    try:
        manifest = json.load(open(os.path.join(REF_DIR, "MANIFEST.json")))
        return all(_sha(os.path.join(REF_DIR, n)) == h for n, h in manifest["files"].items()) and set(FILES) <= set(manifest["files"])
    except Exception:
        return False


if __name__ == "__main__":
    print("staged" if stage() else "reference checkout not found; nothing staged")

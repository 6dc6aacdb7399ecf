"""Recipe: make the UNMODIFIED reference reachable on the GPU box (BASELINE INFRASTRUCTURE ONLY).

    python oracle/make_ref.py

The reference is pure Python (nothing compiles into ``oracle/_ref/``) and cannot be pip-installed
(no setup.py / pyproject; its requirements pin packages that are absent from the offline wheelhouse).
``/root/reference`` does not exist on the GPU box, so this script copies the six hot-path files
(``oracle/ref_shims.REF_FILES``) byte for byte into ``baseline/_ref/src/`` -- git-ignored,
NOT gpurun-ignored, so the copy travels with the snapshot but never enters the history.  ``bench.py
--impl reference`` and ``cpu_baseline`` then time those files through the loader in ``oracle/ref_shims.py``
(``kind: "reference"``); when no copy is reachable they fall back to the numpy port and say so
(``kind: "port"``).  ``__graft_entry__.build()`` runs this recipe whenever ``/root/reference`` is present.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shims import REF_FILES  # noqa: E402

SRC = "/root/reference/src"
DST = os.path.join(ROOT, "baseline", "_ref", "src")


def make_ref(verbose=True):
    """Copy the hot-path files; returns the destination root or None when the reference is not here."""
    if not all(os.path.exists(os.path.join(SRC, f)) for f in REF_FILES):
        return None
    lines = []
    for f in REF_FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), dst)
        with open(dst, "rb") as fh:
            lines.append("%s  %s" % (hashlib.sha256(fh.read()).hexdigest(), f))
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if verbose:
        print("reference hot-path files copied to", DST)
    return DST


if __name__ == "__main__":
    if make_ref() is None:
        print("no /root/reference here: nothing copied")

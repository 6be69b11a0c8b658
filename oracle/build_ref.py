"""Stage the reference's OWN hot-path modules for the CPU arm of bench.py (`--impl reference`, `cpu_baseline`).

TEST / MEASUREMENT INFRASTRUCTURE - never imported by the product.  The reference is pure Python, so "building" it is a
verbatim copy of the six modules of the path from /root/reference into oracle/_ref/ (git-ignored: nothing of the
reference enters the history; NOT gpurun-ignored: the directory travels to the GPU box, where /root/reference does not
exist).  bench.py imports them from there with oracle/ref_shim on the path (a cosmetic `colorama` stand-in; `skimage`
routed to oracle/skimage_compat because scikit-image is not installed in this image).  Run by __graft_entry__.build()
whenever /root/reference is present; a no-op elsewhere (the GPU box uses the staged copy)."""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["src/preprocessing/__init__.py", "src/preprocessing/fingerprint_preprocess.py", "src/preprocessing/orientation.py",
         "src/features/__init__.py", "src/features/extract_features.py", "src/features/post_processing.py"]


def stage() -> bool:
    if not os.path.isdir(os.path.join(REF, "src")):
        return os.path.isfile(os.path.join(DST, FILES[1]))
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isfile(src):
            shutil.copyfile(src, dst)
    return True


if __name__ == "__main__":
    print("oracle/_ref staged" if stage() else "no reference available")

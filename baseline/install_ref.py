#!/usr/bin/env python
"""Places an UNMODIFIED copy of the reference checkout under baseline/_ref/ (git-ignored, NOT gpurun-ignored, so it
travels to the GPU box with the snapshot).  The reference is pure Python with no setup.py, so `pip install --target
baseline/_ref /root/reference` has nothing to build: a plain copy of the tree is the install.  Nothing in it is edited;
the run-time shims the scripts need in this image (stub matplotlib / yacs / imgaug, see baseline/ref_env.py) live outside it.

    python baseline/install_ref.py [/path/to/reference]      # default: $TCVOM_REFERENCE or /root/reference
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def install(src=None) -> bool:
    src = src or os.environ.get("TCVOM_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(src, "models")):
        return os.path.isdir(os.path.join(DST, "models"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(src, DST, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc", "figures", "*.png", "*.jpg",
                                                            "*.gif", "*.mp4"))
    return True


if __name__ == "__main__":
    ok = install(sys.argv[1] if len(sys.argv) > 1 else None)
    print("baseline/_ref", "ready" if ok else "NOT available (no reference checkout found)")

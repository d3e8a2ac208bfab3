"""Stage the UNMODIFIED reference package into the git-ignored ``baseline/_ref/`` so that the GPU box (which has no
/root/reference) can run the reference's own EVQE loop on the B200 primitives (tests/test_gpu_reference_loop.py).

The sanctioned route is ``pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target
baseline/_ref /root/reference``; in this image it fails (the reference's build backend, poetry-core, is not in the wheelhouse:
"No module named 'poetry'"), so the package directory is placed there as pip would have placed it: a plain, unmodified copy
of ``queasars/``.  ``baseline/_ref`` is listed in .gitignore -- nothing of it enters the repository history -- and is not
gpurun-ignored, so it travels to the GPU box with the snapshot.  Run by ``__graft_entry__.build()`` when /root/reference exists."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REFERENCE = "/root/reference"
TARGET = os.path.join(ROOT, "baseline", "_ref")


def stage(force: bool = False) -> str:
    """-> how the package got there ("present", "pip", "copy") or "" when there is no reference checkout."""
    marker = os.path.join(TARGET, "queasars", "__init__.py")
    if os.path.exists(marker) and not force:
        return "present"
    if not os.path.isdir(os.path.join(REFERENCE, "queasars")):
        return ""
    os.makedirs(TARGET, exist_ok=True)
    pip = subprocess.run(
        [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse", "--target", TARGET, REFERENCE],
        capture_output=True, text=True,
    )
    if pip.returncode == 0 and os.path.exists(marker):
        return "pip"
    dst = os.path.join(TARGET, "queasars")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(os.path.join(REFERENCE, "queasars"), dst, ignore=shutil.ignore_patterns("__pycache__"))
    with open(os.path.join(TARGET, "STAGED.txt"), "w") as fh:
        fh.write("unmodified copy of /root/reference/queasars (pip install failed: " + (pip.stderr.strip().splitlines() or ["?"])[-1] + ")\n")
    return "copy"


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv) or "no reference checkout")

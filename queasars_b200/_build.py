"""Build of the native library: one nvcc invocation, sm_100a only, output in-tree next to the sources
(``queasars_b200/csrc/libqueasars_b200.so``) so it travels with the repository snapshot to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libqueasars_b200.so")
SOURCES = [os.path.join(CSRC, "qb_api.cu")]
HEADERS = [os.path.join(CSRC, "qb_kernels.cuh"), os.path.join(ROOT, "include", "queasars_b200.h")]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build_native(force: bool = False, verbose: bool = False, defines: tuple = (), out_path: str = LIB_PATH) -> str:
    """``defines`` / ``out_path``: A/B builds of kernel variants (tools/build_variants.py); the product library has none."""
    if not force and not needs_build() and out_path == LIB_PATH:
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", os.path.join(ROOT, "include"), "-o", out_path, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return out_path


PYHELPER_SRC = os.path.join(CSRC, "qb_pyhelper.c")
PYHELPER_PATH = os.path.join(CSRC, "libqb_pyhelper.so")


def build_pyhelper(force: bool = False) -> str:
    """The host-side list -> float64 marshalling helper (CPython C API; optional: the engine falls back to NumPy without it)."""
    import sysconfig

    if not force and os.path.exists(PYHELPER_PATH) and os.path.getmtime(PYHELPER_PATH) >= os.path.getmtime(PYHELPER_SRC):
        return PYHELPER_PATH
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-I", sysconfig.get_paths()["include"], "-o", PYHELPER_PATH, PYHELPER_SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return PYHELPER_PATH


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))

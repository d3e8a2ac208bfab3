"""Build A/B variants of the native library (kernel experiment switches of csrc/qb_kernels.cuh) next to the product library:
    python tools/build_variants.py            -> queasars_b200/csrc/variants/lib_c{CTAS}_g{GROUP}.so  (+ ptxas register report)
Select one at run time with QB_NATIVE_LIB=<path>.  The .so files are git-ignored; they travel to the GPU box with gpurun."""
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from queasars_b200 import _build  # noqa: E402

OUT = os.path.join(_build.CSRC, "variants")
os.makedirs(OUT, exist_ok=True)
variants = [(4, 1, 0), (4, 1, 1)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for ctas, group, prefetch in variants:
    path = os.path.join(OUT, f"lib_c{ctas}_g{group}_p{prefetch}.so")
    nvcc = "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *_build.NVCC_FLAGS, "-Xptxas=-v", f"-DQB_SWEEP_CTAS={ctas}", f"-DQB_DENSE_GROUP={group}", "-I", os.path.join(_build.ROOT, "include"), "-o", path,
           *_build.SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode:
        print(proc.stderr[-2000:])
        raise SystemExit(1)
    text = proc.stderr
    m = re.search(r"sweep_kernelIdLi4ELi11EjE.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", text, re.S)
    print(f"ctas={ctas} group={group} prefetch={prefetch}: regs={m.group(4)} spill_st={m.group(2)} spill_ld={m.group(3)}  -> {path}")

"""Opcode mix of the dominant kernel from the built library's SASS (cuobjdump -sass): how much of sweep_kernel<double,4,11,unsigned>
is FP64 arithmetic, shared-memory traffic, branches and integer work.  Writes profiles/<round>_sweep_sass_mix.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
lib = os.path.join(ROOT, "queasars_b200", "csrc", "libqueasars_b200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sweep_sass_mix.txt")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
target = next(b for b in blocks if b.startswith("_ZN2qb12sweep_kernelIdLi4ELi11EjEE"))
ops = collections.Counter()
lines = []
for line in target.splitlines():
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        ops[m.group(1).split(".")[0]] += 1
        lines.append(line.rstrip())
total = sum(ops.values())
groups = {
    "FP64 (DFMA DMUL DADD)": ("DFMA", "DMUL", "DADD"),
    "shared memory (LDS STS LDSM)": ("LDS", "STS", "LDSM"),
    "global memory (LDG STG LD ST)": ("LDG", "STG", "LD", "ST"),
    "branches / sync (BRA BRX BSSY BSYNC BAR WARPSYNC EXIT)": ("BRA", "BRX", "BSSY", "BSYNC", "BAR", "WARPSYNC", "EXIT", "CALL", "RET"),
    "integer / logic / moves": ("IMAD", "IADD3", "LOP3", "SHF", "LEA", "ISETP", "MOV", "SEL", "PRMT", "IADD", "POPC", "FLO", "BREV", "R2UR", "S2R", "S2UR", "UMOV", "ULOP3", "UIADD3", "USHF", "ULEA", "UISETP", "UIMAD", "P2R", "R2P", "PLOP3"),
}
with open(out_path, "w") as fh:
    fh.write("SASS opcode mix of qb::sweep_kernel<double, 4, 11, unsigned int> (cuobjdump -sass of the built libqueasars_b200.so, sm_100a)\n")
    fh.write(f"static instructions: {total}\n\n")
    for name, members in groups.items():
        n = sum(ops[o] for o in members)
        fh.write(f"{name:58s} {n:6d}  {100.0 * n / total:5.1f} %\n")
    fh.write("\nby opcode:\n")
    for op, n in ops.most_common(40):
        fh.write(f"  {op:10s} {n:6d}\n")
    # one uncontrolled REAL10 dense body as an excerpt: the longest run of consecutive DFMA/DMUL lines
    best, cur, start = (0, 0), 0, 0
    for i, l in enumerate(lines):
        if re.search(r"\b(DFMA|DMUL)\b", l):
            if cur == 0:
                start = i
            cur += 1
            if cur > best[0]:
                best = (cur, start)
        else:
            cur = 0
    fh.write(f"\nlongest straight run of FP64 instructions ({best[0]}), first 28 lines:\n")
    for l in lines[best[1] : best[1] + 28]:
        fh.write(l + "\n")
print(open(out_path).read()[:1800])

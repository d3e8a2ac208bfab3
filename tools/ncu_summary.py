"""Summarise an `ncu --page source --csv` dump: stall reasons and opcode shares of the first kernel in the file."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = Counter()
nsamp = instr = 0
by_op_s, by_op_i = Counter(), Counter()
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] in ("Kernel Name", "Address"):
        break
    s, n = int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]])
    nsamp += s
    instr += n
    for h in stall_cols:
        tot[h] += int(r[idx[h]])
    src = r[idx["Source"]].strip()
    op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
    by_op_s[op] += s
    by_op_i[op] += n
print(rows[0][1])
print("samples", nsamp, "warp instructions", instr)
print("stalls:", ", ".join(f"{h[6:]} {100 * v / nsamp:.1f}%" for h, v in tot.most_common(9)))
print("opcodes (instr% / sample%):", ", ".join(f"{op} {100 * by_op_i[op] / instr:.1f}/{100 * v / nsamp:.1f}" for op, v in by_op_s.most_common(12)))

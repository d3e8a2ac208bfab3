"""Read-only / write-only / copy bandwidth of this B200 with plain torch kernels (fill_, sum, copy_) on 1 / 4 / 16 GiB buffers:
the ceilings next to which the write-only first sweep and the store-free last sweep of a circuit have to be read
(MEASURED_PEAKS.json holds the copy figure only, read + write bytes counted)."""
import json
import sys

import torch

dev = torch.device("cuda:0")
out = {}
for gib in (0.25, 1, 4, 16):
    n = int(gib * (1 << 30)) // 8
    a = torch.empty(n, dtype=torch.float64, device=dev)
    b = torch.empty(n, dtype=torch.float64, device=dev)
    a.fill_(1.0), b.fill_(2.0)
    res = {}
    for name, fn, nbytes in (("write_only_fill", lambda: a.fill_(3.0), 8 * n), ("write_only_zero", lambda: a.zero_(), 8 * n),
                             ("read_only_sum", lambda: a.sum(), 8 * n), ("copy", lambda: b.copy_(a), 16 * n)):
        for _ in range(3):
            fn()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[name] = {"ms": round(best, 4), "GBps": round(nbytes / (best * 1e-3) / 1e9, 1)}
    out[f"{gib}GiB"] = res
    del a, b
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)

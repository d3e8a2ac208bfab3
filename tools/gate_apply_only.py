"""bench.py's gate_apply probe alone (24 / 26 / 28 / 30 qubits, whole-circuit GB/s): A/B of kernel variants via QB_NATIVE_LIB."""
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from queasars_b200.engine import Engine  # noqa: E402

peak, _ = bench.measured_peak_gbs()
sizes = {24: 6, 26: 6, 28: 4, 30: 4}
only = [int(v) for v in os.environ.get("QB_PROBE_QUBITS", "").split(",") if v]
out = bench.gate_apply_probe(Engine(0), peak, tuple((n, l) for n, l in sizes.items() if not only or n in only))
for n, v in out.items():
    for k in ("fused_evqe", "hbm_regime", "hbm_sparse"):
        e = v[k]
        print(n, k, e["gates_in_sweeps"], e["ms_per_sweep"], "whole %.0f GB/s = %.2f" % (e["GBps_whole_circuit"], e["frac_of_measured_hbm"]),
              "rw %.2f" % e.get("rw_sweeps", {}).get("frac_of_measured_hbm", 0))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"))

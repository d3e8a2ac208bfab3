"""Turn an .ncu-rep (ncu --set full) into the small, committed summaries under profiles/:
   python tools/ncu_profile_summary.py gpurun_out/r1_sweep20.ncu-rep profiles/r1_sweep20 [--traffic-json]"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]
want = [w for w in want if w in idx]
with open(out + "_metrics.csv", "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(want)
    w.writerow([units[idx[c]] for c in want])
    for r in rows[2:]:
        w.writerow([r[idx[c]] for c in want])


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(value) * scale


if "--traffic-json" in sys.argv:
    launches = []
    for r in rows[2:]:
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        launches.append({"grid": r[idx["Grid Size"]], "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr})
    payload = {
        "source": "ncu --set full --clock-control none, tools/profile_case.py --n 20 --layers 6 --batch 32 (the bench.py workload); one entry per sweep launch of one step",
        "per_launch": launches,
        "traffic_bytes_per_step": sum(l["traffic_bytes"] for l in launches),
        "mean_traffic_bytes_per_launch": sum(l["traffic_bytes"] for l in launches) / max(1, len(launches)),
    }
    json.dump(payload, open(out + "_traffic.json", "w"), indent=1)

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
open("/tmp/_src.csv", "w").write(src)
summary = subprocess.run([sys.executable, "tools/ncu_summary.py", "/tmp/_src.csv"], capture_output=True, text=True).stdout
open(out + "_stalls.txt", "w").write("# first captured launch; tools/ncu_summary.py over `ncu --page source --csv`\n" + summary)
print(open(out + "_metrics.csv").read())
print(summary)

#!/bin/bash
# GPU session G (one GPU): select fusion -- parity tests, then A/B on the bench workload and on the gate-apply probe.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2g_gputests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2g_gputests.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "value=%.0f e2e=%.0f hbm_frac=%.3f fp64_frac=%.3f ms=%.4f gates/sweep=%.1f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fp64"]["frac"], d["ms_per_step"], d["roofline"]["gates_per_sweep"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for f in 0 1 0 1; do
  QB_FUSE_SELECT=$f timeout 300 python bench.py --steps 200 --warmup 3 --skip-extras > gpurun_out/r2g_bench_fuse$f.json 2> gpurun_out/r2g_bench_fuse$f.err; summ gpurun_out/r2g_bench_fuse$f.json
done
for f in 0 1; do echo "gate apply, QB_FUSE_SELECT=$f"; QB_FUSE_SELECT=$f timeout 600 python tools/gate_apply_only.py 2>/dev/null | grep fused_evqe; done

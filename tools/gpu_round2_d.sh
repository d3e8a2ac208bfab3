#!/bin/bash
# GPU session D (one GPU): stream-group shapes (L2-resident groups taking turns on the streams) on the bench workload.
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "value=%.0f e2e=%.0f hbm_frac=%.3f fp64_frac=%.3f ms=%.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fp64"]["frac"], d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for cfg in "1 0" "2 0" "2 2" "2 3" "2 4" "2 8" "3 2" "3 3" "4 2" "4 4"; do
  set -- $cfg
  QB_SWEEP_STREAMS=$1 QB_SWEEP_GROUP=$2 timeout 300 python bench.py --steps 200 --warmup 3 --skip-extras > gpurun_out/r2d_bench_s$1_g$2.json 2> gpurun_out/r2d_bench_s$1_g$2.err; summ gpurun_out/r2d_bench_s$1_g$2.json
done
QB_L2_PREFETCH=1 timeout 300 python bench.py --steps 200 --warmup 3 --skip-extras > gpurun_out/r2d_bench_pf1.json 2> gpurun_out/r2d_bench_pf1.err; summ gpurun_out/r2d_bench_pf1.json
timeout 600 python tools/gate_apply_only.py gpurun_out/r2d_gate_apply.json 2> gpurun_out/r2d_gate_apply.err

#!/bin/bash
# GPU session C (one GPU): sweep-stream groups A/B on the bench workload, then the -m gpu tests with the chosen default.
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
for st in 1 2 3 4 8; do
  QB_SWEEP_STREAMS=$st timeout 300 python bench.py --steps 200 --warmup 3 --skip-extras > gpurun_out/r2c_bench_streams$st.json 2> gpurun_out/r2c_bench_streams$st.err; summ gpurun_out/r2c_bench_streams$st.json
done
QB_SWEEP_STREAMS=4 timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2c_gputests_streams4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_gputests_streams4.log

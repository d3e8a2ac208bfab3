#!/bin/bash
# GPU session A: all -m gpu tests, kernel variant A/B, full bench.  Every step under its own timeout.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2_gputests.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2_gputests.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "value=%.0f e2e=%.0f hbm_frac=%.3f fp64_frac=%.3f sweeps=%.2f ms=%.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fp64"]["frac"], d["roofline"]["sweeps_per_evaluation"], d["ms_per_step"]), d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for d in 0 1; do
  QB_DEFER_PHASES=$d timeout 300 python bench.py --steps 100 --warmup 3 --skip-extras > gpurun_out/r2_bench_defer$d.json 2> gpurun_out/r2_bench_defer$d.err; summ gpurun_out/r2_bench_defer$d.json
done
for v in queasars_b200/csrc/variants/*.so; do
  name=$(basename $v .so)
  QB_NATIVE_LIB=$PWD/$v timeout 300 python bench.py --steps 100 --warmup 3 --skip-extras > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err; summ gpurun_out/r2_bench_$name.json
done
timeout 1200 python bench.py --steps 100 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo "full bench rc=$?"; tail -3 gpurun_out/r2_bench_full.err; cat gpurun_out/r2_bench_full.json

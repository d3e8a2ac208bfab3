#!/bin/bash
# GPU session K (one GPU): tiles-per-CTA default x stream groups, gate-apply check, tests.
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "value=%.0f e2e=%.0f fp64_frac=%.3f ms=%.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["fp64"]["frac"], d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for st in 2 3 4; do
  QB_SWEEP_STREAMS=$st timeout 300 python bench.py --steps 200 --warmup 3 --skip-extras > gpurun_out/r2k_bench_s$st.json 2> gpurun_out/r2k_bench_s$st.err; summ gpurun_out/r2k_bench_s$st.json
done
timeout 600 python tools/gate_apply_only.py 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2k_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_gputests.log
timeout 300 python tools/optimizer_pattern.py 2>&1 | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    if isinstance(v,dict): print(' ',k, 'us/call %.1f evals/s %.0f' % (v['us_per_call'], v['evals_per_s']))
"

#!/bin/bash
# GPU session E (one GPU): -m gpu tests (CUDA-graph path, device restore), smoke, pipelined-submission and L2-prefetch A/B with the
# marshalling helper in place, the optimizer calling pattern (batch-1 latency), full bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2e_gputests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2e_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "value=%.0f e2e=%.0f hbm_frac=%.3f fp64_frac=%.3f ms=%.4f e2e_ms=%.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fp64"]["frac"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for pl in 1 0; do
  QB_PIPELINE=$pl timeout 300 python bench.py --steps 200 --warmup 3 --skip-extras > gpurun_out/r2e_bench_pipeline$pl.json 2> gpurun_out/r2e_bench_pipeline$pl.err; summ gpurun_out/r2e_bench_pipeline$pl.json
done
QB_L2_PREFETCH=1 timeout 300 python bench.py --steps 200 --warmup 3 --skip-extras > gpurun_out/r2e_bench_pf1.json 2> gpurun_out/r2e_bench_pf1.err; summ gpurun_out/r2e_bench_pf1.json
for g in 1 0; do echo "optimizer pattern, QB_GRAPHS=$g"; QB_GRAPHS=$g timeout 300 python tools/optimizer_pattern.py 2>&1 | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    if isinstance(v,dict): print(' ',k, 'us/call %.1f evals/s %.0f sweeps %d' % (v['us_per_call'], v['evals_per_s'], v['sweeps']))
"; done
for pf in 0 1; do echo "gate apply, QB_L2_PREFETCH=$pf"; QB_L2_PREFETCH=$pf timeout 600 python tools/gate_apply_only.py gpurun_out/r2e_gate_apply_pf$pf.json 2> gpurun_out/r2e_gate_apply_pf$pf.err | grep hbm_regime; done
timeout 1200 python bench.py > gpurun_out/r2e_bench_full.json 2> gpurun_out/r2e_bench_full.err; echo "full bench rc=$?"; tail -3 gpurun_out/r2e_bench_full.err; summ gpurun_out/r2e_bench_full.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err

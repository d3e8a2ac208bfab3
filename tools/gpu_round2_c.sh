#!/bin/bash
# GPU session C (one GPU): sweep-stream groups and next-tile L2 prefetch, A/B on the bench workload and on the gate-apply probe.
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
V=$PWD/queasars_b200/csrc/variants
for p in 0 1; do
for st in 1 2 4 8; do
  QB_NATIVE_LIB=$V/lib_c4_g1_p$p.so QB_SWEEP_STREAMS=$st timeout 300 python bench.py --steps 200 --warmup 3 --skip-extras > gpurun_out/r2c_bench_p${p}_streams$st.json 2> gpurun_out/r2c_bench_p${p}_streams$st.err; summ gpurun_out/r2c_bench_p${p}_streams$st.json
done
done
for p in 0 1; do
  echo "gate apply, prefetch=$p"
  QB_NATIVE_LIB=$V/lib_c4_g1_p$p.so timeout 600 python tools/gate_apply_only.py gpurun_out/r2c_gate_apply_p$p.json 2> gpurun_out/r2c_gate_apply_p$p.err
done
QB_NATIVE_LIB=$V/lib_c4_g1_p1.so QB_SWEEP_STREAMS=4 timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2c_gputests_p1_streams4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_gputests_p1_streams4.log

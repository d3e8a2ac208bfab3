#!/bin/bash
# Low-latency single-circuit path: mapped pinned parameters / value (QB_ZERO_COPY) and programmatic dependent launches (QB_PDL).
mkdir -p gpurun_out
L=gpurun_out/n_latency.log; : > $L
for cfg in "" "QB_PDL=0" "QB_ZERO_COPY=0" "QB_PDL=0 QB_ZERO_COPY=0"; do
  echo "== latency_breakdown [$cfg]" >> $L
  env $cfg timeout 150 python tools/latency_breakdown.py 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k, v in d.items(): print(k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items()})" >> $L 2>&1
done
echo "== optimizer_pattern [default]" >> $L
timeout 150 python tools/optimizer_pattern.py >> $L 2>&1
B=gpurun_out/n_bench.log; : > $B
for cfg in "" "QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_ld_ca.so" "QB_PDL=2" "QB_PDL=2 QB_SWEEP_STREAMS=1" "QB_PDL=0"; do
  echo "== bench --skip-extras [$cfg]" >> $B
  env $cfg timeout 200 python bench.py --skip-extras 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])" >> $B 2>&1
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/n_tests_default.log 2>&1
QB_PDL=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/n_tests_pdl2.log 2>&1
cat $L $B; tail -3 gpurun_out/n_tests_default.log gpurun_out/n_tests_pdl2.log

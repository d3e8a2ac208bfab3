#!/bin/bash
# A/B of the plain-op fast dispatch (QB_PLAIN_FAST) on the bench workload, parity on the new default.
mkdir -p gpurun_out
B=gpurun_out/t_bench.log; : > $B
for cfg in "" "QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_plain0.so" "" "QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_plain0.so"; do
  echo "== bench --skip-extras [$cfg]" >> $B
  env $cfg timeout 200 python bench.py --skip-extras 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['fp64']['frac'])" >> $B 2>&1
done
QB_PROBE_QUBITS=26,28 timeout 200 python tools/gate_apply_only.py >> $B 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/t_tests.log 2>&1
cat $B; tail -n 3 gpurun_out/t_tests.log

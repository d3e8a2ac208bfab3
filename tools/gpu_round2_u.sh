#!/bin/bash
# A/B of a kernel variant library (last use: control mask fetched with the dispatch word, QB_CMASK)
mkdir -p gpurun_out
B=gpurun_out/v_bench.log; : > $B
for cfg in "" "QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_cmask1.so" "" "QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_cmask1.so"; do
  echo "== bench --skip-extras [$cfg]" >> $B
  env $cfg timeout 200 python bench.py --skip-extras 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['fp64']['frac'])" >> $B 2>&1
done
QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_cmask1.so QB_PROBE_QUBITS=26,28 timeout 200 python tools/gate_apply_only.py >> $B 2>&1
QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_cmask1.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/v_tests.log 2>&1
cat $B; tail -n 3 gpurun_out/v_tests.log

#!/bin/bash
# A/B: control mask fetched with the dispatch word (QB_CMASK) x preferred shared-memory carve-out (QB_SMEM_CARVEOUT)
mkdir -p gpurun_out
B=gpurun_out/w_bench.log; : > $B
for cfg in "QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_cmask0.so" "QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_cmask0.so QB_SMEM_CARVEOUT=100" "QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_cmask1.so QB_SMEM_CARVEOUT=100" "QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_cmask1.so QB_SMEM_CARVEOUT=100" "QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_cmask0.so QB_SMEM_CARVEOUT=100"; do
  echo "== bench --skip-extras [$cfg]" >> $B
  env $cfg timeout 200 python bench.py --skip-extras 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['fp64']['frac'])" >> $B 2>&1
done
QB_NATIVE_LIB=queasars_b200/csrc/variants/lib_cmask1.so QB_SMEM_CARVEOUT=100 QB_PROBE_QUBITS=26,28 timeout 200 python tools/gate_apply_only.py >> $B 2>&1
cat $B

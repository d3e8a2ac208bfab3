#!/bin/bash
# compute-sanitizer over a small pass of every kernel family + the -m gpu suite under the non-default launch switches.
mkdir -p gpurun_out
timeout 120 python tools/sanitizer_case.py > gpurun_out/s_plain.log 2>&1; echo "plain rc=$?"; tail -n 2 gpurun_out/s_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_case.py > gpurun_out/s_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|done" gpurun_out/s_$tool.log | tail -n 6
done
for cfg in "QB_PDL=2" "QB_ZERO_COPY=0 QB_PDL=0" "QB_GRAPHS=0"; do
  echo "== pytest -m gpu [$cfg]" >> gpurun_out/s_switches.log
  env $cfg timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2 >> gpurun_out/s_switches.log
done
cat gpurun_out/s_switches.log

#!/bin/bash
# Sweep images (host-encoded dispatch words) + PDL + mapped I/O: parity, latency, headline.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/o_tests.log 2>&1
L=gpurun_out/o_latency.log; : > $L
for cfg in "" "QB_PDL=0"; do
  echo "== latency_breakdown [$cfg]" >> $L
  env $cfg timeout 150 python tools/latency_breakdown.py 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k, v in d.items(): print(k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items()})" >> $L 2>&1
done
B=gpurun_out/o_bench.log; : > $B
for cfg in "" "QB_TILES_LOG2=2"; do
  echo "== bench --skip-extras [$cfg]" >> $B
  env $cfg timeout 200 python bench.py --skip-extras 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])" >> $B 2>&1
done
QB_PROBE_QUBITS=24,28 timeout 200 python tools/gate_apply_only.py >> $B 2>&1
cat $L $B; tail -n 3 gpurun_out/o_tests.log

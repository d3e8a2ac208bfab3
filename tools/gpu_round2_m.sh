#!/bin/bash
# Directional HBM ceilings + tiles-per-CTA A/B of the gate-apply probe at 24 / 26 qubits.
mkdir -p gpurun_out
timeout 120 python tools/hbm_directional_probe.py gpurun_out/m_hbm_directional.json > gpurun_out/m_hbm_directional.log 2>&1
for t in default 0 1 2 3; do
  if [ "$t" = default ]; then unset QB_TILES_LOG2; else export QB_TILES_LOG2=$t; fi
  echo "== QB_TILES_LOG2=$t" >> gpurun_out/m_tiles.log
  QB_PROBE_QUBITS=24,26 timeout 200 python tools/gate_apply_only.py >> gpurun_out/m_tiles.log 2>&1
done
unset QB_TILES_LOG2
tail -30 gpurun_out/m_hbm_directional.log; cat gpurun_out/m_tiles.log

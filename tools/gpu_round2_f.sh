#!/bin/bash
# GPU session F: gate-apply probe with 2^3 amplitudes per thread / 2^12 tiles (few gates per sweep: is a leaner thread better there?)
for cfg in "4 11" "3 11" "4 12"; do set -- $cfg; echo "== reg_bits=$1 tile_bits=$2"; QB_REG_BITS=$1 QB_TILE_BITS=$2 timeout 600 python tools/gate_apply_only.py 2>&1 | grep -v "^$"; done

#!/bin/bash
# A/B of sweep-kernel settings on the GPU box: 20 q x 32 bench workload and single 26 q states.
# usage: tools/ab_sweep.sh "QB_SWEEP_V1=1" "QB_TILES_LOG2=0" ...   (each argument = one environment setting)
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "=== $cfg"
  env $cfg python bench.py --steps 100 --skip-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('20q x32: evals/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'hbm frac', round(d['roofline']['frac'],3), 'fp64 frac', round(d['roofline']['fp64']['frac'],3))"
  env $cfg python tools/profile_case.py --n 26 --layers 6 2>&1 | head -1
  env $cfg python tools/profile_case.py --n 26 --layers 3 --simple 7 2>&1 | head -1
done

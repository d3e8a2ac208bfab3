#!/bin/bash
# GPU session B (one GPU): defer A/B on the product library, bench + extras, ncu launch list, ncu --set full of the C2 sweeps and of
# the HBM-regime sweeps at 26 and 30 qubits.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "value=%.0f e2e=%.0f hbm_frac=%.3f fp64_frac=%.3f sweeps=%.2f ms=%.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fp64"]["frac"], d["roofline"]["sweeps_per_evaluation"], d["ms_per_step"]), d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for d in 0 1; do
  QB_DEFER_PHASES=$d timeout 300 python bench.py --steps 100 --warmup 3 --skip-extras > gpurun_out/r2b_bench_defer$d.json 2> gpurun_out/r2b_bench_defer$d.err; summ gpurun_out/r2b_bench_defer$d.json
done
timeout 1200 python bench.py --steps 100 > gpurun_out/r2b_bench_full.json 2> gpurun_out/r2b_bench_full.err; echo "full bench rc=$?"; tail -3 gpurun_out/r2b_bench_full.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; cat gpurun_out/r2b_bench_ref.json
timeout 300 python bench.py --steps 2 --warmup 3 --skip-extras > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/profile_case.py --n 20 --layers 6 --batch 32 --runs 1 > gpurun_out/prof_plain20.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel --launch-count 4 -o gpurun_out/r2_sweep20 -f python tools/profile_case.py --n 20 --layers 6 --batch 32 --runs 1 > gpurun_out/ncu_sweep20.log 2>&1
for n in 26 30; do
timeout 300 python tools/profile_case.py --n $n --hbm-regime 1 --runs 1 > gpurun_out/prof_hbm$n.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel --launch-count 3 -o gpurun_out/r2_hbm$n -f python tools/profile_case.py --n $n --hbm-regime 1 --runs 0 > gpurun_out/ncu_hbm$n.log 2>&1
done
cat gpurun_out/prof_plain20.log gpurun_out/prof_hbm26.log gpurun_out/prof_hbm30.log
ls -la gpurun_out/*.ncu-rep

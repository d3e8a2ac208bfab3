#!/bin/bash
# Final one-GPU evidence of the round: -m gpu suite, bench with all extras, reference arm, ncu launch list and --set full of the
# four sweep launches of one bench step (each ncu pass only after the same command exited 0 without it).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/f_tests.log
timeout 1200 python bench.py --steps 100 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/f_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; cat gpurun_out/f_bench_ref.json
timeout 300 python bench.py --steps 2 --warmup 3 --skip-extras > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_bench_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/f_ncu_bench.log 2>&1
timeout 300 python tools/profile_case.py --n 20 --layers 6 --batch 32 --runs 1 > gpurun_out/f_prof_plain20.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel --launch-count 4 -o gpurun_out/f_sweep20 -f python tools/profile_case.py --n 20 --layers 6 --batch 32 --runs 1 > gpurun_out/f_ncu_sweep20.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/f_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"])
for k in ("optimizer_calls", "e2e_threaded", "c3_24q_tfim", "c4_26q_sampler", "c2_complex64", "cpu_baseline"):
    print(k, json.dumps(d.get(k))[:600])
for n, v in d["gate_apply"].items():
    print(n, {k: round(e["frac_of_measured_hbm"], 3) for k, e in v.items()})
PY
ls -la gpurun_out/f_sweep20.ncu-rep

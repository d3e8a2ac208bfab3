#!/bin/bash
# One-GPU evidence run for profiles/: tests, bench line, ncu launch list of the bench command, ncu --set full captures.
# (numbers printed by runs under ncu are never bench values)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 2 --warmup 3 --skip-extras > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/ncu_bench.log 2>&1
python tools/profile_case.py --n 20 --layers 6 --batch 32 --runs 1 > gpurun_out/prof_plain20.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel --launch-count 4 -o gpurun_out/sweep20 -f python tools/profile_case.py --n 20 --layers 6 --batch 32 --runs 1 > gpurun_out/ncu_sweep20.log 2>&1
python tools/profile_case.py --n 26 --layers 6 --runs 1 > gpurun_out/prof_plain26.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel --launch-skip 1 --launch-count 2 -o gpurun_out/sweep26 -f python tools/profile_case.py --n 26 --layers 6 --runs 1 > gpurun_out/ncu_sweep26.log 2>&1
cat gpurun_out/prof_plain20.log gpurun_out/prof_plain26.log

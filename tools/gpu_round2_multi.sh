#!/bin/bash
# Multi-GPU session (gpurun --gpus N): bench at N with the strong / all-devices / sharded extras, the reference arm at N, then the
# -m gpu tests (device sets, sharded states over real peers).   usage: bash tools/gpu_round2_multi.sh N [skip-tests]
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n$N.err; cat gpurun_out/r2_bench_n$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err; cat gpurun_out/r2_bench_ref_n$N.json
if [ -z "$2" ]; then
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/r2_gputests_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gputests_${N}gpu.log
fi

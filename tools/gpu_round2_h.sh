#!/bin/bash
# GPU session H (one GPU): final-state validation -- all -m gpu tests, smoke, the default bench line with every extra, reference arm,
# ncu launch list and the C2 sweep capture of the shipped kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2h_gputests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2h_bench_full.json 2> gpurun_out/r2h_bench_full.err; echo "full bench rc=$?"; tail -3 gpurun_out/r2h_bench_full.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench_full.json"))
print("value=%.0f e2e=%.0f hbm_frac=%.3f fp64_frac=%.3f ms=%.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fp64"]["frac"], d["ms_per_step"]))
for k in ("c3_24q_tfim","c4_26q_sampler","e2e_threaded","c1_jssp_reference_loop","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:700])
for n,v in d["gate_apply"].items():
    e=v["hbm_regime"]; f=v["fused_evqe"]; print(n, "hbm whole %.2f rw %.2f | evqe whole %.2f" % (e["frac_of_measured_hbm"], e["rw_sweeps"]["frac_of_measured_hbm"], f["frac_of_measured_hbm"]))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err; cut -c1-200 gpurun_out/r2h_bench_ref.json
timeout 300 python bench.py --steps 2 --warmup 3 --skip-extras > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/profile_case.py --n 20 --layers 6 --batch 32 --runs 1 > gpurun_out/prof_plain20.log 2>&1 && \
QB_SWEEP_STREAMS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel --launch-count 4 -o gpurun_out/r2_sweep20 -f python tools/profile_case.py --n 20 --layers 6 --batch 32 --runs 1 > gpurun_out/ncu_sweep20.log 2>&1
for n in 26 30; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel --launch-count 3 -o gpurun_out/r2_hbm$n -f python tools/profile_case.py --n $n --hbm-regime 1 --runs 0 > gpurun_out/ncu_hbm$n.log 2>&1
done
cat gpurun_out/prof_plain20.log

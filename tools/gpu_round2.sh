#!/bin/bash
# Evidence round for the rows beside the env step (policy forward f1, rollout loop, obs normaliser): all GPU tests, smoke,
# the contract bench line, then ncu: launch list of the packaged rollout and one full capture of the statistics kernel
# (each only after the plain command exited 0).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --steps ${STEPS:-50} --warmup 12 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log | cut -c1-600
CMD="env GRAPH=0 REPS=1 POLICY=kernel python tools/rollout_bench.py"
$CMD > gpurun_out/rollout_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rollout_launches.csv $CMD > gpurun_out/rollout_ncu_list.log 2>&1
echo "ncu rollout list rc=$?"
python tools/gpu_normalizer.py > gpurun_out/norm_n1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:obs_stats_partial -s 10 -c 1 -f -o gpurun_out/prof_norm python tools/gpu_normalizer.py > gpurun_out/norm_ncu_full.log 2>&1
echo "ncu norm rc=$?"; tail -1 gpurun_out/norm_n1.log

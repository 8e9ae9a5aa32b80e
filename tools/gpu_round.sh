#!/bin/bash
# One GPU-box round: parity tests, smoke, bench (both arms), then the ncu launch list and one full capture of the
# fused kernel (each ncu run only after the same command exited 0 without ncu).  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.csv 2>&1
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --steps ${STEPS:-50} --warmup 12 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.log
if [ -z "$NO_NCU" ]; then
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vnl_env_kernel -s 5 -c 1 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log | cut -c1-200
fi

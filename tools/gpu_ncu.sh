#!/bin/bash
# one full ncu capture of the fused step kernel (after the same command ran clean)
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vnl_env_kernel -s 5 -c 1 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log; cat gpurun_out/plain2.log | cut -c1-400

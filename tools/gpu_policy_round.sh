#!/bin/bash
# Row f1 evidence on one B200: GPU tests, smoke, rollout with the policy kernel and with torch/cuBLAS, policy timing,
# then the ncu launch list and one full capture of the policy kernel (each only after the plain command exited 0).
set -x
timeout -s KILL 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout -s KILL 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
POLICY=kernel timeout -s KILL 300 python tools/rollout_bench.py > gpurun_out/rollout_kernel_n1.log 2>&1; echo "rollout kernel rc=$?"
POLICY=torch timeout -s KILL 300 python tools/rollout_bench.py > gpurun_out/rollout_torch_n1.log 2>&1; echo "rollout torch rc=$?"
timeout -s KILL 150 python tools/gpu_policy.py 300 > gpurun_out/pol0.log 2>&1 && \
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/policy_launches.csv python tools/gpu_policy.py 300 > gpurun_out/pol_ncu_list.log 2>&1 && \
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:vnl_policy_kernel -s 8 -c 1 -f -o gpurun_out/prof_policy python tools/gpu_policy.py 300 > gpurun_out/pol_ncu_full.log 2>&1
echo "ncu rc=$?"
tail -2 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/rollout_kernel_n1.log; tail -1 gpurun_out/rollout_torch_n1.log; tail -1 gpurun_out/pol0.log

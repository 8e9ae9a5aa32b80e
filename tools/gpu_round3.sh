python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 150 python tools/gpu_policy.py 300 > gpurun_out/pol0.log 2>&1 && \
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:vnl_policy_kernel -s 8 -c 1 -f -o gpurun_out/prof_policy python tools/gpu_policy.py 300 > gpurun_out/pol_ncu_full.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/pol0.log | cut -c1-120
POLICY=kernel timeout -s KILL 300 python tools/rollout_bench.py > gpurun_out/rollout_kernel_n1.log 2>&1; tail -1 gpurun_out/rollout_kernel_n1.log
GRAPH=1 timeout -s KILL 300 python tools/rollout_bench.py > gpurun_out/rollout_graph_n1.log 2>&1; tail -1 gpurun_out/rollout_graph_n1.log
python bench.py --steps 50 --warmup 12 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench.log

timeout -s KILL 300 python -m pytest tests/test_rollout.py -x -q -m gpu > gpurun_out/rollout_test.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/rollout_test.log
GRAPH=1 timeout -s KILL 300 python tools/rollout_bench.py > gpurun_out/rollout_graph_n1.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/rollout_graph_n1.log

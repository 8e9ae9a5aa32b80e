TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547"
GRAPH=1 timeout -s KILL 400 $TR tools/rollout_bench.py > gpurun_out/rollout_graph_n8.log 2>&1; echo rc=$?; tail -1 gpurun_out/rollout_graph_n8.log
timeout -s KILL 300 $TR tools/gpu_normalizer.py > gpurun_out/norm_n8.log 2>&1; echo rc=$?; tail -1 gpurun_out/norm_n8.log

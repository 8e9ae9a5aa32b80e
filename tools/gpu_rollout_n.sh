#!/bin/bash
# rollout (config 4) at N GPUs with the tcgen05 policy kernel: eager two-launch loop and the packaged CUDA-graph loop
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
POLICY=kernel timeout -s KILL 400 $TR tools/rollout_bench.py > gpurun_out/rollout_kernel_n$N.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/rollout_kernel_n$N.log
GRAPH=1 timeout -s KILL 400 $TR tools/rollout_bench.py > gpurun_out/rollout_graph_n$N.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/rollout_graph_n$N.log

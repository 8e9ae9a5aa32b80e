#!/bin/bash
# env-group width experiment: parity + throughput at 1 and 2 warps per env
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for ew in ${EWS:-1 2}; do
  echo "=== VNL_ENV_WARPS=$ew"
  if [ -z "$NOTEST" ]; then VNL_ENV_WARPS=$ew timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -${PT:-8}; fi
  for w in ${WARPS:-0}; do echo "== VNL_WARPS=$w"; VNL_ENV_WARPS=$ew VNL_WARPS=$w B=${B:-4096} timeout 300 python tools/gpu_prof.py 2>&1 | tail -${TL:-30}; done
done

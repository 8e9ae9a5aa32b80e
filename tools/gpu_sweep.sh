#!/bin/bash
# throughput sweep over env-group width x lockstep groups (x envs per CTA)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for ew in ${EWS:-1 2}; do for g in ${GROUPS_:-1 2 3 7}; do for w in ${WARPS:-0}; do
  echo -n "EW=$ew LSGROUPS=$g WARPS=$w LOCKSTEP=${VNL_LOCKSTEP:-1}: "
  VNL_ENV_WARPS=$ew VNL_LSGROUPS=$g VNL_WARPS=$w B=${B:-4096} timeout 300 python tools/gpu_prof.py 2>&1 | tail -1
done; done; done

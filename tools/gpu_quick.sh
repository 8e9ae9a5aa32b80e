#!/bin/bash
# quick GPU check: parity tests + phase profile / throughput
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv
if [ -z "$NOTEST" ]; then python -m pytest tests -x -q -m gpu 2>&1 | tail -${PT:-15}; fi
for w in ${WARPS:-7}; do echo "== VNL_WARPS=$w"; VNL_WARPS=$w B=${B:-4096} python tools/gpu_prof.py 2>&1 | tail -${TL:-22}; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv

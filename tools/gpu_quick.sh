#!/bin/bash
# quick GPU check: parity tests + phase profile / throughput
python -m pytest tests -x -q -m gpu 2>&1 | tail -${PT:-15}
for w in ${WARPS:-7}; do echo "== VNL_WARPS=$w"; VNL_WARPS=$w B=${B:-4096} python tools/gpu_prof.py 2>&1 | tail -22; done

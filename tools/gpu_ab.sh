#!/bin/bash
# A/B of library builds on the same box, interleaved: VNL_B200_LIB picks the .so (libvnl_b200_<variant>.so)
for rep in 1 2 3; do for v in ${VARIANTS:-old new}; do
  echo -n "$v: "; VNL_B200_LIB=$PWD/vnl-brax-imitation_b200/libvnl_b200_$v.so B=${B:-4096} timeout 300 python tools/gpu_prof.py 2>&1 | tail -1
done; done

B=4096 python tools/gpu_prof.py 2>&1 | tail -${TL:-24}

#!/bin/bash
# memcheck of the fused step on a tiny batch (one tool per gpurun call)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "
import __graft_entry__ as g
g.smoke()
" 2>&1 | tail -25

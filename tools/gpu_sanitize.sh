#!/bin/bash
# compute-sanitizer passes over the fused step on tiny batches: memcheck, then racecheck (shared-memory hazards between
# the tenants of the phase-recycled layout), then initcheck is skipped (the workspace is write-before-read by design).
# TOOL=memcheck|racecheck|synccheck selects one.
for tool in ${TOOL:-memcheck racecheck}; do
  echo "=== $tool"
  timeout ${TMO:-600} compute-sanitizer --tool $tool --error-exitcode 3 python - <<'PY' 2>&1 | tail -${TL:-25}
import importlib, numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
envs = importlib.import_module("vnl-brax-imitation_b200.envs")
rod = importlib.import_module("vnl-brax-imitation_b200.envs.rodent")
model, clip = rod.packaged_rodent()
env = envs.RodentTracking(reference_clip=clip, model=model, device="cuda:0", **rod.RODENT_ENV_ARGS)
B = 20  # two CTAs' worth of slots are not filled: partial CTAs, inactive warps and the lockstep barriers are exercised
s = env.reset(np.random.default_rng(0), batch_size=B)
for i in range(2):
    a = torch.rand(B, 30, device="cuda") * 2 - 1
    s = env.step(s, a)
torch.cuda.synchronize()
print("rodent ok", float(s.reward.sum()))
antm = importlib.import_module("vnl-brax-imitation_b200.envs.ant")
m, c = antm.packaged_ant()
ant = envs.AntTracking(model=m, reference_clip=c, device="cuda:0")
s = ant.reset(batch_size=5)
s = ant.step(s, torch.rand(5, 8, device="cuda") * 2 - 1)
torch.cuda.synchronize()
print("ant ok", float(s.reward.sum()))
PY
done

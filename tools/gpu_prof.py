"""Developer tool: per-phase clock64 profile of one CTA of the fused step + throughput sweep."""
import ctypes
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
envs = importlib.import_module("vnl-brax-imitation_b200.envs")
rod = importlib.import_module("vnl-brax-imitation_b200.envs.rodent")
PHASES = ["fk", "com/cinert/cdof", "vel/acc down + crb/rne up", "smooth/act", "M", "factor+K", "solve(smooth)", "constraints",
          "warm select", "solver init", "linesearch", "update+beta", "(fwd tail)", "euler factor", "euler rest", "load + task outputs", "  factor: eliminate", "  factor: normalise", "  factor: invert", "  ls: mul_m", "  ls: jmul", "  ls: sums+points", "  upd: jtmul_force", "  upd: reductions", "  upd: solve_m", "  fk: local transforms", "  fk: compose levels", "  vel/acc down", "  local rne"]


def main():
    BB = int(os.environ.get("B", 4096))
    model, clip = rod.packaged_rodent()
    env = envs.RodentTracking(reference_clip=clip, model=model, **rod.RODENT_ENV_ARGS)
    eng = env.engine
    rng = np.random.default_rng(0)
    rt = env._ref_traj
    start = rng.integers(0, 235, size=BB).astype(np.int32)
    qp = np.hstack([rt.position[start], rt.quaternion[start], rt.joints[start]]).astype(np.float32)
    qp += (1e-3 * rng.standard_normal(qp.shape)).astype(np.float32)
    qv = np.hstack([rt.velocity[start], rt.angular_velocity[start], rt.joints_velocity[start]]).astype(np.float32)
    s0 = env.reset_from(qp, qv, start)
    a = torch.rand(BB, 30, device="cuda") * 2 - 1
    st_a = dict(s0.pipeline_state); st_a["cur_frame"] = s0.info["cur_frame"]; st_a["sub_clip_frame"] = s0.info["sub_clip_frame"]
    st_b, out = eng.alloc_state(BB), eng.alloc_outputs(BB)
    for _ in range(5):
        eng.step(st_a, a, st_b, out)
        st_a, st_b = st_b, st_a
    torch.cuda.synchronize()
    prof = torch.zeros(32, dtype=torch.int64, device="cuda")
    A, Bs, O = eng._state(st_a), eng._state(st_b), eng._outputs(out)
    rc = eng.lib.vnl_step_profiled(eng._cref, eng.model_dev.data_ptr(), eng.task_dev.data_ptr(), BB, ctypes.byref(A), a.data_ptr(),
                                   ctypes.byref(Bs), ctypes.byref(O), eng._stream(), prof.data_ptr(), BB // 2)
    torch.cuda.synchronize()
    assert rc == 0
    pr = prof.cpu().numpy()
    tot = pr.sum()
    print("phase profile (env %d of %d, %d envs per CTA), total %.0f kcycles" % (BB // 2, BB, eng.envs_per_cta, tot / 1e3))
    for i, n in enumerate(PHASES):
        print("  %-16s %9.1f kcyc %5.1f%%" % (n, pr[i] / 1e3, 100.0 * pr[i] / max(tot, 1)))
    print("  stats", out["stats"][BB // 2].tolist())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        eng.step(st_a, a, st_b, out)
        st_a, st_b = st_b, st_a
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("B=%d  %.3f ms/step  %.3f M env-steps/s  mean stats %s" % (BB, ms, BB / ms / 1e3, out["stats"].float().mean(0).tolist()))


if __name__ == "__main__":
    main()

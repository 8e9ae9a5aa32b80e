"""Developer check on a GPU box: stage-by-stage kernel-vs-oracle comparison + quick timing."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402

envs = importlib.import_module("vnl-brax-imitation_b200.envs")
rod = importlib.import_module("vnl-brax-imitation_b200.envs.rodent")


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    m = np.isfinite(a) & np.isfinite(b)
    if not m.any():
        return float("nan"), 0
    return float(np.abs(a - b)[m].max() / (np.abs(b)[m].max() + 1e-30)), int(m.sum())


def main():
    B = int(os.environ.get("B", 16))
    model, clip = rod.packaged_rodent()
    env = envs.RodentTracking(reference_clip=clip, model=model, **rod.RODENT_ENV_ARGS)
    eng = env.engine
    dims = eng.dims
    print("dims", dims, "smem", eng.smem_bytes, flush=True)
    rng = np.random.default_rng(0)
    start = rng.integers(0, 235, size=B).astype(np.int32)
    rt = env._ref_traj
    qpos = np.hstack([rt.position[start], rt.quaternion[start], rt.joints[start]]).astype(np.float32)
    qpos = qpos + (1e-3 * rng.standard_normal(qpos.shape)).astype(np.float32)
    qvel = np.hstack([rt.velocity[start], rt.angular_velocity[start], rt.joints_velocity[start]]).astype(np.float32)
    # ---- stage dump ----
    st = dict(qpos=torch.tensor(qpos, device="cuda"), qvel=torch.tensor(qvel, device="cuda"))
    dump = eng.forward_dump(st).cpu().numpy().astype(np.float64)
    torch.cuda.synchronize()
    g = oracle.split_dump(dims, dump)
    ost = dict(qpos=qpos.astype(np.float64), qvel=qvel.astype(np.float64))
    o32 = oracle.forward_dump(env.model_blob, ost, None, precision=32, dims=dims)
    o64 = oracle.forward_dump(env.model_blob, ost, None, precision=64, dims=dims)
    print("%-16s %12s %12s %8s" % ("stage", "gpu-vs-o32", "o32-vs-o64", "n"))
    for name, _ in oracle.dump_layout(dims):
        if name == "subtree_com":
            ga, a32, a64 = g[name][:, 1], o32[name][:, 1], o64[name][:, 1]
        else:
            ga, a32, a64 = g[name], o32[name], o64[name]
        e1, n = relerr(ga, a32)
        e2, _ = relerr(np.where(np.isfinite(ga), a32, np.nan), a64)
        print("%-16s %12.3e %12.3e %8d" % (name, e1, e2, n))
    print("counters gpu", g["counters"][:4].astype(int).tolist(), "\n         o32", o32["counters"][:4].astype(int).tolist())
    # ---- reset + steps, teacher-forced against the oracle ----
    s = env.reset_from(qpos, qvel, start)
    torch.cuda.synchronize()
    so, oo = oracle.reset(env.model_blob, env.task_blob, qpos.astype(np.float64), qvel.astype(np.float64), start, precision=32,
                          dims=dims, obs_size=eng.obs_size, traj_size=eng.traj_size)
    for k in ("qpos", "qacc_warmstart", "xpos", "subtree_com", "qfrc_actuator"):
        print("reset", k, relerr(s.pipeline_state[k].cpu().numpy(), so[k])[0])
    print("reset obs", relerr(s.obs.cpu().numpy(), oo["obs"])[0], "traj", relerr(s.info["traj"].cpu().numpy(), oo["traj"])[0],
          "term", relerr(s.info["termination_error"].cpu().numpy(), oo["metrics"][:, 6])[0])
    act_rng = np.random.default_rng(1)
    for it in range(12):
        action = act_rng.uniform(-1, 1, size=(B, 30)).astype(np.float32)
        # teacher forcing: oracle steps from the GPU's current state
        st_np = {k: v.cpu().numpy().astype(np.float64) for k, v in s.pipeline_state.items()}
        st_np["cur_frame"] = s.info["cur_frame"].cpu().numpy()
        st_np["sub_clip_frame"] = s.info["sub_clip_frame"].cpu().numpy()
        so, oo = oracle.step(env.model_blob, env.task_blob, st_np, action.astype(np.float64), precision=32, dims=dims,
                             obs_size=eng.obs_size, traj_size=eng.traj_size)
        s64, o64_ = oracle.step(env.model_blob, env.task_blob, st_np, action.astype(np.float64), precision=64, dims=dims,
                                obs_size=eng.obs_size, traj_size=eng.traj_size)
        s = env.step(s, torch.tensor(action, device="cuda"))
        torch.cuda.synchronize()
        g_ = lambda k: s.pipeline_state[k].cpu().numpy()
        print("step %2d qpos %.2e (o32/o64 %.2e) qvel %.2e (%.2e) rew %.2e done_eq %s frames %s stats gpu %s o32 %s" % (
            it, relerr(g_("qpos"), so["qpos"])[0], relerr(so["qpos"], s64["qpos"])[0], relerr(g_("qvel"), so["qvel"])[0],
            relerr(so["qvel"], s64["qvel"])[0], np.abs(s.reward.cpu().numpy() - oo["reward"]).max(),
            bool((s.done.cpu().numpy() == oo["done"]).all()), bool((s.info["cur_frame"].cpu().numpy() == so["cur_frame"]).all()),
            s.info["solver_stats"].cpu().numpy().mean(0).round(1).tolist(), oo["stats"].mean(0).round(1).tolist()), flush=True)
    print("obs", relerr(s.obs.cpu().numpy(), oo["obs"])[0], "traj", relerr(s.info["traj"].cpu().numpy(), oo["traj"])[0],
          "metrics", np.abs(np.stack([s.metrics[k].cpu().numpy() for k in rod.METRIC_KEYS], 1) - oo["metrics"]).max(0))
    # ---- timing ----
    for BB in (1024, 4096, 16384):
        start = rng.integers(0, 235, size=BB).astype(np.int32)
        qp = np.hstack([rt.position[start], rt.quaternion[start], rt.joints[start]]).astype(np.float32)
        qv = np.hstack([rt.velocity[start], rt.angular_velocity[start], rt.joints_velocity[start]]).astype(np.float32)
        s0 = env.reset_from(qp, qv, start)
        a = torch.rand(BB, 30, device="cuda") * 2 - 1
        st_a = dict(s0.pipeline_state); st_a["cur_frame"] = s0.info["cur_frame"]; st_a["sub_clip_frame"] = s0.info["sub_clip_frame"]
        st_b, out = eng.alloc_state(BB), eng.alloc_outputs(BB)
        for _ in range(3):
            eng.step(st_a, a, st_b, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            eng.step(st_a, a, st_b, out)
            st_a, st_b = st_b, st_a
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("B=%d  %.3f ms/step  %.3f M env-steps/s  stats %s" % (BB, ms, BB / ms / 1e3, out["stats"].float().mean(0).tolist()), flush=True)


if __name__ == "__main__":
    main()

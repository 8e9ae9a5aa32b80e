"""Error-vs-step curves of the fused env step against the CPU oracle (SURVEY section 7 "report both"):

  * regime: SETTLED CONTACT with BOUNDED random actions (the fp64 oracle settles clip poses for 600 substeps; actions
    U(-amp, amp) held for one env step), 20 env steps = 100 physics substeps, B envs;
  * free-run: kernel, fp32 oracle and fp64 oracle each evolve their own state from the common start;
  * teacher-forced: every env step starts from the KERNEL's state; one step of each is compared;
  * columns: kernel-vs-o32 beside o32-vs-o64 (the oracle's own rounding spread): relative qpos / qvel error (max |d| over the
    env's vector / max |ref| over the batch), absolute reward error; median and max over envs.

Writes profiles/r02_parity_curve.json.  Usage: python tools/parity_curve.py [amp=0.3] [B=64]"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402  (developer tool: the checker)

envs = importlib.import_module("vnl-brax-imitation_b200.envs")
rod = importlib.import_module("vnl-brax-imitation_b200.envs.rodent")
KEYS = ("qpos", "qvel", "act", "qacc_warmstart", "xpos", "xquat", "subtree_com", "qfrc_actuator")


def to_np(state):
    d = {k: v.cpu().numpy().astype(np.float64) for k, v in state.pipeline_state.items()}
    d["cur_frame"] = state.info["cur_frame"].cpu().numpy()
    d["sub_clip_frame"] = state.info["sub_clip_frame"].cpu().numpy()
    return d


def err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    e = np.abs(a - b).reshape(a.shape[0], -1).max(1) / (np.abs(b).max() + 1e-30)
    return float(np.median(e)), float(e.max())


def main():
    amp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    T = 20
    model, clip = rod.packaged_rodent()
    env = envs.RodentTracking(reference_clip=clip, model=model, **rod.RODENT_ENV_ARGS)
    eng = env.engine
    c = env._ref_traj
    fr = (np.arange(B) * 3) % 200
    qpos = np.hstack([c.position[fr], c.quaternion[fr], c.joints[fr]]).astype(np.float64)
    settled, _ = oracle.pipeline_step(env.model_blob, dict(qpos=qpos, qvel=np.zeros((B, 73))), None, 600, precision=64, dims=eng.dims)
    qp, qv = settled["qpos"].astype(np.float32), settled["qvel"].astype(np.float32)
    kw = dict(dims=eng.dims, obs_size=eng.obs_size, traj_size=eng.traj_size)
    rng = np.random.default_rng(0)
    acts = rng.uniform(-amp, amp, size=(T, B, 30)).astype(np.float32)
    s = env.reset_from(qp, qv, fr.astype(np.int32))
    s32, _ = oracle.reset(env.model_blob, env.task_blob, qp.astype(np.float64), qv.astype(np.float64), fr.astype(np.int32), precision=32, **kw)
    s64, _ = oracle.reset(env.model_blob, env.task_blob, qp.astype(np.float64), qv.astype(np.float64), fr.astype(np.int32), precision=64, **kw)
    rows = []
    for t in range(T):
        a64 = acts[t].astype(np.float64)
        k_in = to_np(s)
        # teacher-forced: one step of each from the kernel's state
        tf32, tfo32 = oracle.step(env.model_blob, env.task_blob, k_in, a64, precision=32, **kw)
        tf64, tfo64 = oracle.step(env.model_blob, env.task_blob, k_in, a64, precision=64, **kw)
        s = env.step(s, torch.tensor(acts[t], device="cuda"))
        torch.cuda.synchronize()
        g = to_np(s)
        # free-run: each from its own state
        s32, o32 = oracle.step(env.model_blob, env.task_blob, s32, a64, precision=32, **kw)
        s64, o64 = oracle.step(env.model_blob, env.task_blob, s64, a64, precision=64, **kw)
        rew = s.reward.cpu().numpy().astype(np.float64)
        row = {"env_step": t + 1, "substeps": 5 * (t + 1)}
        for k in ("qpos", "qvel"):
            row["tf_kernel_vs_o32_" + k] = err(g[k], tf32[k]); row["tf_o32_vs_o64_" + k] = err(tf32[k], tf64[k])
            row["free_kernel_vs_o32_" + k] = err(g[k], s32[k]); row["free_o32_vs_o64_" + k] = err(s32[k], s64[k])
        ar = lambda x, y: (float(np.median(np.abs(x - y))), float(np.abs(x - y).max()))
        row["tf_kernel_vs_o32_reward"] = ar(rew, tfo32["reward"]); row["tf_o32_vs_o64_reward"] = ar(tfo32["reward"], tfo64["reward"])
        row["free_kernel_vs_o32_reward"] = ar(rew, o32["reward"]); row["free_o32_vs_o64_reward"] = ar(o32["reward"], o64["reward"])
        row["done_equal_tf"] = bool(np.array_equal(s.done.cpu().numpy(), tfo32["done"]))
        st = s.info["solver_stats"].cpu().numpy()
        row["active_contacts_per_substep"] = float(st[:, 2].mean() / 5)
        row["contact_counts_equal_tf"] = float((st[:, 2] == tfo32["stats"][:, 2]).mean())
        rows.append(row)
        print(json.dumps(row))
    out = {"what": "rodent, settled contact, U(-%.2f, %.2f) actions, %d envs, [median, max] over envs; qpos / qvel relative, reward absolute" % (amp, amp, B),
           "rows": rows}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_parity_curve_amp%.2f.json" % amp), "w"), indent=1)


if __name__ == "__main__":
    main()

"""Dump golden vectors from the REAL reference (needs jax, mujoco, mujoco-mjx, brax, dm_control and the reference
checkout -- none of which exist in the build image, so this script is never run there).

    python tools/dump_reference.py /path/to/VNL-Brax-Imitation tests/golden/rodent_reference_dump.npz

Writes, for B fixed-seed start states and T random-action steps: the inputs (qpos, qvel, act, qacc_warmstart, action,
cur_frame, sub_clip_frame) and the reference outputs per step (mjx.Data fields of oracle.dump_layout that MJX
materialises, obs, traj, reward, done, metrics).  tests/test_oracle.py::test_reference_dump consumes the file."""
import os
import sys

import numpy as np


def main(ref, out, B=8, T=20, seed=0):
    sys.path.insert(0, ref)
    os.chdir(ref)
    import jax
    import jax.numpy as jp
    from preprocessing import mjx_preprocess as mp  # noqa: F401
    import pickle
    import yaml
    from envs.rodent import RodentTracking

    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, here)
    import importlib
    clipm = importlib.import_module("vnl-brax-imitation_b200.clip")
    old = clipm.load_pickle(os.path.join(ref, "clips", "transform_snips_groom.p"))
    qpos = np.hstack([old.position, old.quaternion, old.joints]).astype(np.float32)
    tmp = "/tmp/_vnl_raw_clip.p"
    pickle.dump({"qpos": qpos}, open(tmp, "wb"))
    clip = mp.process_clip(tmp, start_step=0, clip_length=250)
    cfg = yaml.safe_load(open(os.path.join(ref, "configs", "env_config.yaml")))["rodent"]["env_args"]
    env = RodentTracking(reference_clip=clip, **{k: v for k, v in cfg.items() if k != "stac_path"})
    step = jax.jit(env.step)
    reset = jax.jit(env.reset)
    rng = np.random.default_rng(seed)
    rec = {k: [] for k in ("qpos_in", "qvel_in", "act_in", "warm_in", "action", "cur_frame_in", "sub_clip_frame_in", "qpos", "qvel",
                           "act", "qacc_warmstart", "xpos", "xquat", "subtree_com", "qfrc_actuator", "qM", "qfrc_bias",
                           "qfrc_passive", "qacc", "qfrc_constraint", "obs", "traj", "reward", "done")}
    for b in range(B):
        st = reset(jax.random.PRNGKey(seed + b))
        for t in range(T):
            a = jp.asarray(rng.uniform(-1, 1, size=env.action_size).astype(np.float32))
            d0 = st.pipeline_state
            rec["qpos_in"].append(np.asarray(d0.qpos)); rec["qvel_in"].append(np.asarray(d0.qvel))
            rec["act_in"].append(np.asarray(d0.act)); rec["warm_in"].append(np.asarray(d0.qacc_warmstart))
            rec["action"].append(np.asarray(a)); rec["cur_frame_in"].append(int(st.info["cur_frame"]))
            rec["sub_clip_frame_in"].append(int(st.info["sub_clip_frame"]))
            st = step(st, a)
            d = st.pipeline_state
            for k in ("qpos", "qvel", "act", "qacc_warmstart", "xpos", "xquat", "subtree_com", "qfrc_actuator", "qM", "qfrc_bias",
                      "qfrc_passive", "qacc", "qfrc_constraint"):
                rec[k].append(np.asarray(getattr(d, k)))
            rec["obs"].append(np.asarray(st.obs)); rec["traj"].append(np.asarray(st.info["traj"]))
            rec["reward"].append(float(st.reward)); rec["done"].append(float(st.done))
    np.savez_compressed(out, B=B, T=T, **{k: np.asarray(v) for k, v in rec.items()})
    print("wrote", out)


if __name__ == "__main__":
    main(*sys.argv[1:3])

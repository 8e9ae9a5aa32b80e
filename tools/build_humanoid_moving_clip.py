"""A MOVING reference clip for HumanoidTracking (BASELINE configs[2]): the reference's `clips/humanoid_traj_stand.p`
is git-ignored and absent, and the stand-in used so far is one pose tiled 256 times.  This script makes a second packaged clip out of
states of a physics rollout of the CPU oracle: the humanoid starts standing (qpos0) with a joint-space PD controller (through its torque
actuators) following small phase-shifted sinusoidal joint targets.  Without a balance controller it sinks and topples within ~0.5 s, and
everything after that lies on the floor, outside the env's healthy range -- so the clip keeps the frames while the torso is above 1.06 m
(a 20 cm dip with swinging arms, feet loaded) and plays them forwards and backwards (ping-pong) to fill 256 frames: every frame is a
state the physics visited, the reference keeps moving, it stays inside `healthy_z_range`, and tracking it needs the feet on the floor.
The frames go through the same pipeline as the rodent's clips (`clip.process_clip_qpos`: kinematics of every frame + finite-difference
velocities) plus the `center_of_mass` field humanoid.py:279 reads.

    python tools/build_humanoid_moving_clip.py     ->  vnl-brax-imitation_b200/data/humanoid_moving_clip.npz

Test infrastructure writes it (the oracle), the product only reads the resulting table: `envs.humanoid.packaged_humanoid(moving=True)`."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = lambda n: importlib.import_module("vnl-brax-imitation_b200." + n)


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    hum, mb, clipm, mjcf = P("envs.humanoid"), P("model_blob"), P("clip"), P("mjcf")
    model, _ = hum.packaged_humanoid()
    blob = mb.build_model_blob(model)
    dims = mb.read_dims(blob)
    A = model.arrays
    T, n_frames, kp, kd, amp = 256, 5, 6.0, 0.4, 0.3
    dt = model.timestep * n_frames
    jq = 7 + A["actuator_dofadr"] - 6  # qpos address of each actuated hinge
    legs = np.array([0.3 if any(s in n for s in ("hip", "knee", "ankle")) else 1.0 for n in model.act_names])
    k = np.arange(model.nu)
    st = dict(qpos=A["qpos0"][None].astype(np.float64), qvel=np.zeros((1, model.nv)), act=np.zeros((1, model.na)),
              qacc_warmstart=np.zeros((1, model.nv)))
    seg = [st["qpos"][0].copy()]
    for t in range(200):
        tgt = A["qpos0"][jq] + amp * legs * np.sin(2 * np.pi * (0.5 + 0.05 * k) * t * dt + 0.9 * k)
        u = np.clip(kp * (tgt - st["qpos"][0][jq]) - kd * st["qvel"][0][A["actuator_dofadr"]], -1.0, 1.0)
        st, _ = oracle.pipeline_step(blob, st, u[None], n_frames, precision=64, dims=dims)
        if st["qpos"][0][2] < 1.06:
            break
        seg.append(st["qpos"][0].copy())
    seg = np.array(seg)
    order = list(range(len(seg))) + list(range(len(seg) - 2, 0, -1))  # one ping-pong period
    qpos = seg[[order[i % len(order)] for i in range(T)]]
    clip = clipm.process_clip_qpos(model, qpos, max_qvel=20.0, dt=dt)
    com = []
    for t in range(T):
        kin = mjcf.kinematics(model, qpos[t])
        com.append(mjcf.subtree_com(model, kin["xipos"])[1])
    clip.center_of_mass = np.asarray(com, dtype=np.float32)
    out = os.path.join(ROOT, "vnl-brax-imitation_b200", "data", "humanoid_moving_clip.npz")
    np.savez_compressed(out, **clipm.clip_to_npz_dict(clip))
    print("wrote", out, "%d rollout frames, period %d, root z %.3f .. %.3f, largest joint excursion %.2f rad"
          % (len(seg), len(order), qpos[:, 2].min(), qpos[:, 2].max(), float(np.ptp(qpos[:, 7:], axis=0).max())))


if __name__ == "__main__":
    main()

"""Reference-clip tables: loading the reference's pickles and restating `process_clip`.

Reference: `preprocessing/mjx_preprocess.py:21-193` (new 8-field `ReferenceClip`, built by
`process_clip` from raw mocap qpos) and `mocap_preprocess.py:326-340` (old 13-field clip, the
format of the only clip shipped in the reference repo, `clips/transform_snips_groom.p`).

jax / flax / dm_control are absent here, so the pickle is read with a stub unpickler and
`process_clip` is restated in numpy float64 on top of `mjcf.kinematics`.  The fields the old
clip stores redundantly (body positions / quaternions, velocities) are the known-answer
check for this restatement (tests/test_mjcf_clip.py).
"""
from __future__ import annotations

import pickle
from dataclasses import dataclass, fields, replace
from typing import Optional

import numpy as np

from . import mjcf


class _OldClip:
    def __setstate__(self, st):
        self.__dict__.update(st)


def _reconstruct(fun, args, state, aval=None):
    a = fun(*args)
    a.__setstate__(state)
    return a


class _StubUnpickler(pickle.Unpickler):
    """Reads clips pickled with jax arrays / `mocap_preprocess.ReferenceClip` without jax."""

    def find_class(self, module, name):
        if name == "ReferenceClip":
            return _OldClip
        if module == "jax._src.array" and name == "_reconstruct_array":
            return _reconstruct
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core")
        return super().find_class(module, name)


def load_pickle(path: str):
    with open(path, "rb") as f:
        return _StubUnpickler(f).load()


@dataclass
class ReferenceClip:
    """Field-for-field mirror of `mjx_preprocess.ReferenceClip` (numpy instead of jax arrays)."""
    position: Optional[np.ndarray] = None
    quaternion: Optional[np.ndarray] = None
    joints: Optional[np.ndarray] = None
    body_positions: Optional[np.ndarray] = None
    velocity: Optional[np.ndarray] = None
    joints_velocity: Optional[np.ndarray] = None
    angular_velocity: Optional[np.ndarray] = None
    body_quaternions: Optional[np.ndarray] = None
    center_of_mass: Optional[np.ndarray] = None  # old 13-field format only (mocap_preprocess.py:326-340); humanoid.py:279 reads it

    def replace(self, **kw):
        return replace(self, **kw)


# ---- preprocessing/transformations.py restated in numpy ------------------------------------
def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]], dtype=q.dtype)


def quat_diff(source, target):
    """`transformations.py:102-114`."""
    return mjcf.quat_mul(quat_conj(source), target).astype(source.dtype)


def quat_to_axisangle(quat):
    """`transformations.py:117-139` (angle wrapped to [-pi, pi), zero below 1e-10)."""
    dt = quat.dtype.type
    angle = dt(2) * np.arccos(np.clip(quat[0], dt(-1), dt(1)))
    if angle < 1e-10:
        return np.zeros(3, dtype=quat.dtype)
    qn = np.sin(angle / dt(2))
    angle = (angle + dt(np.pi)) % dt(2 * np.pi) - dt(np.pi)
    return quat[1:4] / qn * angle


def compute_velocity_from_kinematics(qpos_trajectory: np.ndarray, dt: float) -> np.ndarray:
    """`mjx_preprocess.py:170-193`, evaluated in the dtype of `qpos_trajectory`."""
    q = qpos_trajectory
    t = q.dtype.type
    lin = (q[1:, :3] - q[:-1, :3]) / t(dt)
    gyro = []
    for i in range(q.shape[0] - 1):
        d = quat_diff(q[i, 3:7], q[i + 1, 3:7])
        d = d / np.linalg.norm(d).astype(q.dtype)
        gyro.append(quat_to_axisangle(d) / t(dt))
    jnt = (q[1:, 7:] - q[:-1, 7:]) / t(dt)
    return np.concatenate([lin, np.stack(gyro), jnt], axis=1)


def process_clip_qpos(model: mjcf.Model, mocap_qpos: np.ndarray, max_qvel: float = 20.0,
                      dt: float = 0.02) -> ReferenceClip:
    """`process_clip` from an in-memory `[T, nq]` qpos array (`mjx_preprocess.py:88-107`).

    FK runs in float64 on the fp32 inputs and results are stored as fp32 (the reference runs
    the scan in fp32 on device; differences are at fp32 rounding)."""
    mocap_qpos = np.asarray(mocap_qpos, dtype=np.float32)
    T = mocap_qpos.shape[0]
    pos, quat, jnt, xpos, xquat = [], [], [], [], []
    for t in range(T):
        k = mjcf.kinematics(model, mocap_qpos[t].astype(np.float64))
        qn = k["qpos"]
        pos.append(qn[:3]); quat.append(qn[3:7]); jnt.append(qn[7:])
        xpos.append(k["xpos"]); xquat.append(k["xquat"])
    padded = np.concatenate([mocap_qpos, mocap_qpos[-1:]], axis=0)
    qvel = compute_velocity_from_kinematics(padded, dt)
    qvel[:, 6:] = np.clip(qvel[:, 6:], -max_qvel, max_qvel)
    f32 = lambda a: np.asarray(a, dtype=np.float32)
    return ReferenceClip(position=f32(pos), quaternion=f32(quat), joints=f32(jnt), body_positions=f32(xpos),
                         velocity=f32(qvel[:, :3]), joints_velocity=f32(qvel[:, 6:]),
                         angular_velocity=f32(qvel[:, 3:6]), body_quaternions=f32(xquat))


def process_clip(stac_path: str, mjcf_path: str = "./assets/rodent.xml", scale_factor: float = 0.9,
                 start_step: int = 0, clip_length: int = 250, max_qvel: float = 20.0,
                 dt: float = 0.02) -> ReferenceClip:
    """Signature-compatible `process_clip` (`mjx_preprocess.py:43-50`); also accepts the old
    13-field clip pickle, from which qpos is rebuilt as hstack(position, quaternion, joints)."""
    d = load_pickle(stac_path)
    if isinstance(d, dict):
        qpos = np.asarray(d["qpos"])
    else:
        qpos = np.hstack([np.asarray(d.position), np.asarray(d.quaternion), np.asarray(d.joints)])
    qpos = qpos[start_step:start_step + clip_length]
    model = mjcf.load_rodent(mjcf_path, scale_factor, torque=False)
    return process_clip_qpos(model, qpos, max_qvel, dt)


def clip_to_npz_dict(clip: ReferenceClip) -> dict:
    return {f.name: getattr(clip, f.name) for f in fields(clip) if getattr(clip, f.name) is not None}


def clip_from_npz(path: str) -> ReferenceClip:
    z = np.load(path)
    return ReferenceClip(**{f.name: z[f.name] for f in fields(ReferenceClip) if f.name in z.files})


def tiled_clip(model: mjcf.Model, qpos: np.ndarray, length: int = 256) -> ReferenceClip:
    """Synthetic clip = one pose repeated `length` times with zero velocities, the way the reference's notebooks built
    `ant_traj_still.p` (tile of the reset state, `notebooks/environments_explore.ipynb`) and the stand-in used here for the
    absent `clips/humanoid_traj_stand.p`.  body_positions holds ALL bodies (humanoid.py:257 compares against data.xpos)."""
    k = mjcf.kinematics(model, np.asarray(qpos, dtype=np.float64))
    com = mjcf.subtree_com(model, k["xipos"])[1]
    f32 = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float32), (length,) + np.shape(a)))
    q = k["qpos"]
    nv = model.nv
    return ReferenceClip(position=f32(q[:3]), quaternion=f32(q[3:7]), joints=f32(q[7:]), body_positions=f32(k["xpos"]),
                         velocity=f32(np.zeros(3)), joints_velocity=f32(np.zeros(nv - 6)), angular_velocity=f32(np.zeros(3)),
                         body_quaternions=f32(k["xquat"]), center_of_mass=f32(com))

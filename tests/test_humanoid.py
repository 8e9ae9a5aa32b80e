"""HumanoidTracking (envs/humanoid.py) -- same kernel, humanoid task switches.

CPU: the oracle's humanoid task logic against a line-by-line numpy restatement of envs/humanoid.py:185-311,346-433.
GPU: kernel vs oracle on the humanoid model (per-stage arrays, reset, task logic on the kernel's own state, teacher-forced
steps bounded by the oracle's fp32-vs-fp64 spread)."""
import numpy as np
import pytest

from conftest import pkg

mjcf = pkg("mjcf")


@pytest.fixture(scope="module")
def humanoid():
    hum, mb = pkg("envs.humanoid"), pkg("model_blob")
    model, clip = hum.packaged_humanoid()
    task_blob, obs_size, traj_size = hum.humanoid_task_tables(model, clip)
    model_blob = mb.build_model_blob(model)
    return dict(model=model, clip=clip, model_blob=model_blob, task_blob=task_blob, dims=mb.read_dims(model_blob),
                obs_size=obs_size, traj_size=traj_size)


def test_humanoid_dims_and_sizes(humanoid):
    d, m = humanoid["dims"], humanoid["model"]
    assert (m.nbody, m.nq, m.nv, m.nu, m.na) == (17, 28, 27, 21, 0)
    assert (d["ncon"], d["nlimit"], d["nefc"], d["eulerdamp"], d["solver"], d["iterations"]) == (10, 21, 61, 0, 1, 6)
    assert (humanoid["obs_size"], humanoid["traj_size"]) == (55, 630)
    assert m.body_names[2] == "head" or "head" in m.body_names  # notebooks: head body id 2


def _states(humanoid, B, seed):
    m = humanoid["model"]
    rng = np.random.default_rng(seed)
    qpos = np.tile(m.arrays["qpos0"], (B, 1)).astype(np.float32)
    lo, hi = m.arrays["jnt_range"][1:, 0], m.arrays["jnt_range"][1:, 1]
    qpos[:, 7:] += (0.1 * rng.standard_normal((B, m.nq - 7)) * np.minimum(1.0, (hi - lo))).astype(np.float32)
    qpos[:, 2] += rng.uniform(-0.02, 0.02, size=B).astype(np.float32)
    qvel = (0.2 * rng.standard_normal((B, m.nv))).astype(np.float32)
    return qpos, qvel, rng.integers(0, 95, size=B).astype(np.int32)


def _numpy_humanoid_task(h, old, new, xmat_torso, cur_frame_old):
    """envs/humanoid.py: reward + done from the OLD state, obs / traj from the new one."""
    c = h["clip"]
    f64 = lambda a: np.asarray(a, np.float64)
    T = c.position.shape[0]
    f = min(max(int(cur_frame_old), 0), T - 1)
    cur = cur_frame_old + 1
    rcom = np.exp(-100 * np.linalg.norm(old["subtree_com"] - f64(c.center_of_mass[f])))
    qvel_ref = np.hstack([f64(c.velocity[f]), f64(c.angular_velocity[f]), f64(c.joints_velocity[f])])
    rvel = np.exp(-0.1 * np.linalg.norm(old["qvel"] - qvel_ref))
    ej = np.mean(np.abs(f64(c.joints[f]) - old["qpos"][7:]))
    eb = np.mean(np.abs(f64(c.body_positions[f]) - old["xpos"]))
    rtrunk = 1 - (0.5 * 1.0 * eb + 0.5 * ej) / float(np.float32(0.9))  # the task blob holds the threshold as fp32
    qs = old["qpos"][3:7] / np.linalg.norm(old["qpos"][3:7])
    qt = f64(c.quaternion[f]) / np.linalg.norm(f64(c.quaternion[f]))
    rquat = np.exp(-2 * np.abs(0.5 * np.arccos(min(1.0, 2 * float(qs @ qt) ** 2 - 1))))
    ract = -0.015 * np.mean(np.square(old["qfrc_actuator"]))
    healthy = 0.0 if old["qpos"][2] < 1.0 else 1.0
    healthy = 0.0 if old["qpos"][2] > 2.0 else healthy
    done = 1.0 if rtrunk < 0.5 else 0.0
    rcom, rvel, rtrunk, rquat, ract = rcom * 0.01, rvel * 0.01, rtrunk * 0.01, rquat * 0.01, ract * 1e-4
    reward = rcom + rvel + rtrunk + rquat + ract
    done = max(1.0 - healthy, done)
    obs = np.hstack([new["qpos"], new["qvel"]])
    s = min(max(cur + 1, 0), T - 5)
    w = slice(s, s + 5)
    diff = f64(c.body_positions[w]) - new["xpos"][None]
    traj = np.hstack([(diff @ xmat_torso).flatten(), diff.flatten(), ((f64(c.position[w]) - new["qpos"][:3]) @ xmat_torso).flatten(),
                      (f64(c.joints[w]) - new["qpos"][7:]).flatten()])
    return dict(reward=reward, done=done, obs=obs, traj=traj, metrics=[rcom, rvel, rtrunk, rquat, ract, 0.0, rtrunk], cur_frame=cur)


@pytest.mark.parametrize("cur0", [0, 57, 253, 400])
def test_oracle_humanoid_step_matches_numpy_restatement(oracle_mod, humanoid, cur0):
    h = humanoid
    B = 3
    qpos, qvel, start = _states(h, B, 1)
    kw = dict(precision=64, dims=h["dims"], obs_size=h["obs_size"], traj_size=h["traj_size"])
    s0, _ = oracle_mod.reset(h["model_blob"], h["task_blob"], qpos, qvel, start, **kw)
    s0["cur_frame"][:] = cur0
    if cur0 == 57:
        s0["qpos"][1, 2] = 0.9  # unhealthy on the OLD state -> done
    action = np.random.default_rng(2).uniform(-1.2, 1.2, size=(B, 21))
    s1, o1 = oracle_mod.step(h["model_blob"], h["task_blob"], s0, action, **kw)
    for e in range(B):
        old = {k: s0[k][e] for k in ("qpos", "qvel", "xpos", "subtree_com", "qfrc_actuator")}
        new = {k: s1[k][e] for k in ("qpos", "qvel", "xpos")}
        want = _numpy_humanoid_task(h, old, new, mjcf.quat_to_mat(s1["xquat"][e, 1]), cur0)
        assert s1["cur_frame"][e] == want["cur_frame"]
        assert o1["done"][e] == want["done"]
        # rquat = exp(-arccos(.)) with the root orientation ON the reference: arccos near 1 amplifies the last ulp to ~1e-10
        assert abs(o1["reward"][e] - want["reward"]) < 1e-9
        dm = np.abs(o1["metrics"][e] - np.array(want["metrics"]))
        assert np.delete(dm, 3).max() < 1e-12 and dm[3] < 1e-9
        assert np.abs(o1["obs"][e] - want["obs"]).max() < 1e-12
        assert np.abs(o1["traj"][e] - want["traj"]).max() < 1e-12
    if cur0 == 57:
        assert o1["done"][1] == 1.0


def test_oracle_humanoid_reset_and_stand(oracle_mod, humanoid):
    h = humanoid
    m = h["model"]
    qpos = np.tile(m.arrays["qpos0"], (2, 1)).astype(np.float64)
    kw = dict(precision=64, dims=h["dims"], obs_size=h["obs_size"], traj_size=h["traj_size"])
    s0, o0 = oracle_mod.reset(h["model_blob"], h["task_blob"], qpos, np.zeros((2, 27)), np.array([0, 9], np.int32), **kw)
    assert np.abs(o0["metrics"][:, 6] - 1.0).max() < 1e-6  # the tiled clip IS the reset pose: zero tracking error
    assert (o0["done"] == 0).all() and np.abs(o0["obs"][:, :28] - s0["qpos"]).max() == 0
    s1, o1 = oracle_mod.step(h["model_blob"], h["task_blob"], s0, np.zeros((2, 21)), **kw)
    assert (o1["done"] == 0).all() and (o1["reward"] > 0.03).all()  # rcom, rvel, rtrunk, rquat all ~0.01 on the reference pose
    assert (o1["stats"][:, 2] > 0).all()  # feet in contact during the step


def test_moving_humanoid_clip(oracle_mod):
    """The packaged moving clip (tools/build_humanoid_moving_clip.py: rollout states, ping-pong): it moves, stays inside the env's
    healthy range, its derived fields are consistent with the kinematics, and the oracle env steps on it with the feet in contact --
    tracking error small when the env starts ON the reference, larger a few frames later under zero action (the reference moves on)."""
    hum, mb = pkg("envs.humanoid"), pkg("model_blob")
    model, clip = hum.packaged_humanoid(moving=True)
    T = clip.position.shape[0]
    assert T == 256 and clip.body_positions.shape == (256, model.nbody, 3) and clip.center_of_mass.shape == (256, 3)
    z = clip.position[:, 2]
    assert z.min() > 1.0 and z.max() < 2.0 and np.ptp(z) > 0.15            # healthy throughout, and it dips by > 15 cm
    assert float(np.ptp(clip.joints, axis=0).max()) > 1.0                   # arms swing by more than a radian
    assert np.abs(clip.velocity).max() > 0.3 and np.abs(clip.joints_velocity).max() > 2.0
    for t in (0, 17, 40, 255):                                              # fields == kinematics of the frame's qpos
        q = np.concatenate([clip.position[t], clip.quaternion[t], clip.joints[t]]).astype(np.float64)
        k = mjcf.kinematics(model, q)
        assert np.abs(k["xpos"] - clip.body_positions[t]).max() < 1e-6
        assert np.abs(mjcf.subtree_com(model, k["xipos"])[1] - clip.center_of_mass[t]).max() < 1e-6
    task_blob, obs_size, traj_size = hum.humanoid_task_tables(model, clip)
    model_blob = mb.build_model_blob(model)
    kw = dict(precision=64, dims=mb.read_dims(model_blob), obs_size=obs_size, traj_size=traj_size)
    start = np.array([0, 10, 30, 60], np.int32)
    qpos = np.hstack([clip.position[start], clip.quaternion[start], clip.joints[start]]).astype(np.float64)
    qvel = np.hstack([clip.velocity[start], clip.angular_velocity[start], clip.joints_velocity[start]]).astype(np.float64)
    s, o = oracle_mod.reset(model_blob, task_blob, qpos, qvel, start, **kw)
    assert (o["done"] == 0).all() and np.abs(o["metrics"][:, 6] - 1.0).max() < 1e-5  # on the reference: zero termination error
    err = []
    for _ in range(6):
        s, o = oracle_mod.step(model_blob, task_blob, s, np.zeros((4, model.nu)), **kw)
        err.append(1.0 - o["metrics"][:, 6])
        assert np.isfinite(o["reward"]).all() and np.isfinite(o["obs"]).all()
    assert (o["stats"][:, 2] > 0).all()                                    # feet on the floor
    assert (err[-1] > err[0]).all() and (err[-1] > 1e-3).all()             # a limp body does not follow a moving reference


# ---- GPU ------------------------------------------------------------------------------------------------------------
def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.fixture(scope="module")
def gpu_humanoid(humanoid):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    hum = pkg("envs.humanoid")
    return hum.HumanoidTracking(model=humanoid["model"], reference_clip=humanoid["clip"], device="cuda:0")


@pytest.mark.gpu
def test_gpu_humanoid_forward_stages(gpu_humanoid, humanoid, oracle_mod):
    import torch
    h = humanoid
    B = 12
    qpos, qvel, _ = _states(h, B, 3)
    qpos[:, 2] -= 0.05  # push the feet into the floor: active contacts
    rng = np.random.default_rng(4)
    ctrl = rng.uniform(-1.5, 1.5, size=(B, 21)).astype(np.float32)
    warm = rng.standard_normal((B, 27)).astype(np.float32)
    eng = gpu_humanoid.engine
    st = {k: torch.tensor(v, device="cuda") for k, v in dict(qpos=qpos, qvel=qvel, qacc_warmstart=warm).items()}
    g = oracle_mod.split_dump(h["dims"], eng.forward_dump(st, torch.tensor(ctrl, device="cuda")).cpu().numpy().astype(np.float64))
    ost = {k: v.astype(np.float64) for k, v in dict(qpos=qpos, qvel=qvel, qacc_warmstart=warm).items()}
    o32 = oracle_mod.forward_dump(h["model_blob"], ost, ctrl.astype(np.float64), precision=32, dims=h["dims"])
    assert (o32["counters"][:, 2] > 0).any()
    for name, tol in dict(xpos=2e-6, xquat=2e-6, xipos=2e-6, xanchor=2e-6, xaxis=3e-6, cinert=3e-6, cdof=3e-6, crb=3e-6, qM=3e-6,
                          cvel=3e-6, qfrc_bias=2e-5, qfrc_actuator=1e-6, qfrc_smooth=2e-5, qacc_smooth=1e-4, con_dist=5e-6,
                          con_frame=3e-6, qacc=5e-3, qfrc_constraint=5e-3).items():
        assert np.isfinite(g[name]).all(), name
        assert _rel(g[name], o32[name]) < tol, (name, _rel(g[name], o32[name]))
    assert np.array_equal(g["counters"][:, 2:4], o32["counters"][:, 2:4])
    assert np.array_equal(np.isfinite(g["efc_pos"]), o32["efc_pos"] < 0)


@pytest.mark.gpu
def test_gpu_humanoid_task_logic_and_steps(gpu_humanoid, humanoid, oracle_mod):
    import torch
    h = humanoid
    B = 10
    qpos, qvel, start = _states(h, B, 5)
    s0 = gpu_humanoid.reset_from(qpos, qvel, start)
    kw = dict(precision=32, dims=h["dims"], obs_size=h["obs_size"], traj_size=h["traj_size"])
    so, oo = oracle_mod.reset(h["model_blob"], h["task_blob"], qpos, qvel, start, **kw)
    assert _rel(s0.obs.cpu().numpy(), oo["obs"]) < 1e-6 and _rel(s0.info["traj"].cpu().numpy(), oo["traj"]) < 3e-6
    assert np.abs(s0.info["termination_error"].cpu().numpy() - oo["metrics"][:, 6]).max() < 1e-5
    rng = np.random.default_rng(6)
    to_np = lambda st: dict({k: v.cpu().numpy().astype(np.float64) for k, v in st.pipeline_state.items()},
                            cur_frame=st.info["cur_frame"].cpu().numpy(), sub_clip_frame=st.info["sub_clip_frame"].cpu().numpy())
    for it in range(6):
        a = rng.uniform(-1, 1, size=(B, 21)).astype(np.float32)
        old = to_np(s0)
        o32s, o32o = oracle_mod.step(h["model_blob"], h["task_blob"], old, a.astype(np.float64), **kw)
        o64s, _ = oracle_mod.step(h["model_blob"], h["task_blob"], old, a.astype(np.float64), **dict(kw, precision=64))
        s1 = gpu_humanoid.step(s0, torch.tensor(a, device="cuda"))
        new = to_np(s1)
        assert np.array_equal(new["cur_frame"], o32s["cur_frame"])
        # reward and done depend only on the OLD state for the humanoid: they must match the oracle tightly
        assert np.array_equal(s1.done.cpu().numpy(), o32o["done"])
        assert np.abs(s1.reward.cpu().numpy() - o32o["reward"]).max() < 1e-5
        for e in range(B):
            want = _numpy_humanoid_task(h, {k: old[k][e] for k in ("qpos", "qvel", "xpos", "subtree_com", "qfrc_actuator")},
                                        {k: new[k][e] for k in ("qpos", "qvel", "xpos")}, mjcf.quat_to_mat(new["xquat"][e, 1]),
                                        int(old["cur_frame"][e]))
            assert np.abs(s1.obs[e].cpu().numpy() - want["obs"]).max() < 1e-6 * max(1.0, np.abs(want["obs"]).max())
            assert np.abs(s1.info["traj"][e].cpu().numpy() - want["traj"]).max() < 3e-6 * max(1.0, np.abs(want["traj"]).max())
        for k in ("qpos", "qvel"):
            eg = np.abs(new[k] - o32s[k]).max(1) / (np.abs(o32s[k]).max() + 1e-30)
            eo = np.abs(o32s[k] - o64s[k]).max(1) / (np.abs(o32s[k]).max() + 1e-30)
            assert np.median(eg) < 10 * np.median(eo) + 1e-4, (it, k, np.median(eg), np.median(eo))
        s0 = s1

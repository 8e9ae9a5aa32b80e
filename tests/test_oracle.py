"""CPU tests of the oracle (oracle/vnl_oracle.cpp).  The reference pins no dynamics (SURVEY 8c:
"parity unpinned"), so besides the golden FK/COM vectors the restatement is held to physics
invariants, to an independent numpy formulation of the mass matrix / gravity torque, to its fp64
twin, and -- for the env logic -- to a line-by-line numpy restatement of envs/rodent.py."""
import numpy as np
import pytest

from conftest import pkg, start_states

mjcf = pkg("mjcf")


def _fwd(oracle_mod, rodent, qpos, qvel, precision=64, ctrl=None, act=None, warm=None):
    st = dict(qpos=np.asarray(qpos, np.float64), qvel=np.asarray(qvel, np.float64))
    if act is not None:
        st["act"] = np.asarray(act, np.float64)
    if warm is not None:
        st["qacc_warmstart"] = np.asarray(warm, np.float64)
    return oracle_mod.forward_dump(rodent["model_blob"], st, ctrl, precision=precision, dims=rodent["dims"])


def test_oracle_fk_com_match_clip_golden(oracle_mod, rodent, golden):
    frames = np.arange(0, 250, 5)
    qpos = np.hstack([golden["position"][frames], golden["quaternion"][frames], golden["joints"][frames]])
    for prec, tol in ((64, 1e-7), (32, 2e-6)):
        d = _fwd(oracle_mod, rodent, qpos, np.zeros((len(frames), 73)), precision=prec)
        ids = rodent["idx"]["body_idxs"]
        assert np.abs(d["xpos"][:, ids] - golden["body_positions"][frames]).max() < tol
        assert np.abs(d["subtree_com"][:, 1] - golden["center_of_mass"][frames]).max() < tol
        q, g = d["xquat"][:, ids], golden["body_quaternions"][frames].astype(np.float64)
        s = np.sign((q * g).sum(-1, keepdims=True))
        assert np.abs(q * s - g).max() < tol


def test_mass_matrix_two_formulations(oracle_mod, rodent):
    """Composite-rigid-body qM (oracle) == sum_b J_b^T I_b J_b (mjcf.mass_matrix); symmetric, PD."""
    qpos, qvel, _ = start_states(rodent, 3, seed=3)
    d = _fwd(oracle_mod, rodent, qpos, qvel)
    for e in range(3):
        M = d["qM"][e]
        M2 = mjcf.mass_matrix(rodent["model"], qpos[e].astype(np.float64))
        assert np.abs(M - M.T).max() == 0
        # the blob holds the tables as fp32 (as mjx.put_model does); mjcf.mass_matrix uses the float64 compile
        assert np.abs(M - M2).max() < 2e-7 * np.abs(M2).max()
        assert np.linalg.eigvalsh(M).min() > 0
        # fwd_acceleration: M qacc_smooth = qfrc_smooth
        assert np.abs(M @ d["qacc_smooth"][e] - d["qfrc_smooth"][e]).max() < 1e-9 * np.abs(d["qfrc_smooth"][e]).max()
        assert np.allclose(d["qfrc_smooth"][e], d["qfrc_passive"][e] - d["qfrc_bias"][e] + d["qfrc_actuator"][e], atol=1e-15)


def test_gravity_torque_at_rest(oracle_mod, rodent):
    """qvel = 0  =>  qfrc_bias = -sum_b m_b J_b^T g (RNE against the Jacobian formulation)."""
    m = rodent["model"]
    qpos, _, _ = start_states(rodent, 2, seed=4)
    d = _fwd(oracle_mod, rodent, qpos, np.zeros((2, 73)))
    for e in range(2):
        k = mjcf.kinematics(m, qpos[e].astype(np.float64))
        want = np.zeros(m.nv)
        for b in range(1, m.nbody):
            jp_, _ = mjcf.body_jacobian(m, k, b, k["xipos"][b])
            want -= m.arrays["body_mass"][b] * jp_.T @ np.asarray(m.gravity, np.float64)
        assert np.abs(d["qfrc_bias"][e] - want).max() < 2e-7 * np.abs(want).max()  # fp32 tables vs float64 compile


def test_fp32_vs_fp64_stages(oracle_mod, rodent):
    qpos, qvel, _ = start_states(rodent, 8, seed=5)
    a, b = _fwd(oracle_mod, rodent, qpos, qvel, 32), _fwd(oracle_mod, rodent, qpos, qvel, 64)
    for name, tol in (("xpos", 2e-6), ("cinert", 2e-6), ("cdof", 3e-6), ("qM", 3e-6), ("cvel", 3e-6), ("qfrc_bias", 3e-5),
                      ("qfrc_smooth", 3e-5), ("con_dist", 3e-6), ("con_frame", 3e-6), ("efc_D", 1e-4), ("efc_aref", 1e-4)):
        err = np.abs(a[name] - b[name]).max() / (np.abs(b[name]).max() + 1e-30)
        assert err < tol, (name, err)


def test_free_fall_com_and_unconstrained_solution(oracle_mod, rodent):
    """Lifted clear of the floor with limits inactive => no active rows: qacc = qacc_smooth, and the
    whole-body COM follows semi-implicit Euler free fall exactly (internal forces cancel)."""
    m = rodent["model"]
    qpos = np.zeros((1, 74)); qpos[0, :] = m.arrays["qpos0"]; qpos[0, 2] = 0.45
    lo, hi = m.arrays["jnt_range"][1:, 0], m.arrays["jnt_range"][1:, 1]
    qpos[0, 7:] = 0.5 * (lo + hi)
    d = _fwd(oracle_mod, rodent, qpos, np.zeros((1, 73)))
    assert d["counters"][0, 2] == 0 and d["counters"][0, 3] == 0
    assert np.abs(d["qacc"][0] - d["qacc_smooth"][0]).max() == 0 and np.abs(d["qfrc_constraint"]).max() == 0
    st = dict(qpos=qpos, qvel=np.zeros((1, 73)))
    n, dt, g = 20, 0.002, 9.81
    com0 = _fwd(oracle_mod, rodent, qpos, np.zeros((1, 73)))["subtree_com"][0, 1]
    out, stats = oracle_mod.pipeline_step(rodent["model_blob"], st, None, n, precision=64, dims=rodent["dims"])
    assert stats[0, 2] == 0
    # subtree_com lags qpos by one substep: it reflects n-1 integrations
    want_z = com0[2] - g * dt * dt * (n - 1) * n / 2
    assert abs(out["subtree_com"][0, 2] - want_z) < 1e-9
    assert np.abs(out["subtree_com"][0, :2] - com0[:2]).max() < 1e-7  # O(dt^2) joint-space integration error


def test_contact_pushes_out_and_settles(oracle_mod, rodent):
    """Clip start states sit with the tail inside the floor.  Constraint forces must push
    up, the state must stay finite and bounded, and the deepest penetration must decay."""
    c = rodent["fclip"]
    fr = np.array([0, 40, 90, 150])  # frames whose straight tail dips 3-5 cm below the floor
    qpos = np.hstack([c.position[fr], c.quaternion[fr], c.joints[fr]]).astype(np.float64)
    qvel = np.zeros((4, 73))
    d0 = _fwd(oracle_mod, rodent, qpos, 0 * qvel)
    assert (d0["counters"][:, 2] >= 1).all()
    pen0 = d0["con_dist"].min(1)
    assert (pen0 < -1e-3).all()
    assert (d0["qfrc_constraint"][:, 2] > 0).all()  # net upward force on the root
    st = dict(qpos=qpos.astype(np.float64), qvel=0.0 * qvel.astype(np.float64))
    out, _ = oracle_mod.pipeline_step(rodent["model_blob"], st, None, 400, precision=64, dims=rodent["dims"])
    assert np.isfinite(out["qpos"]).all() and np.abs(out["qvel"]).max() < 50.0
    d1 = _fwd(oracle_mod, rodent, out["qpos"], out["qvel"], act=out["act"], warm=out["qacc_warmstart"])
    assert (d1["con_dist"].min(1) > -0.002).all()  # soft-contact equilibrium penetration is below a millimetre
    assert np.abs(out["qvel"]).max() < 5.0 and (out["qpos"][:, 2] > 0.0).all()


def test_limit_rows(oracle_mod, rodent):
    m = rodent["model"]
    qpos = np.zeros((1, 74)); qpos[0] = m.arrays["qpos0"]; qpos[0, 2] = 0.45
    lo, hi = m.arrays["jnt_range"][1:, 0], m.arrays["jnt_range"][1:, 1]
    qpos[0, 7:] = 0.5 * (lo + hi)
    j = 20
    qpos[0, 7 + j - 1] = hi[j - 1] + 0.05  # joint j (dof 5 + j) beyond its upper limit
    d = _fwd(oracle_mod, rodent, qpos, np.zeros((1, 73)))
    assert d["counters"][0, 3] == 1
    r = int(np.flatnonzero(d["efc_pos"][0, :67] < 0)[0])
    assert abs(d["efc_pos"][0, r] + 0.05) < 1e-7  # range table is fp32
    assert d["efc_J"][0, r, 5 + j] == -1.0 and np.abs(d["efc_J"][0, r]).sum() == 1.0
    assert d["qfrc_constraint"][0, 5 + j] < 0  # pushes back inside the range


# ---- envs/rodent.py restated in numpy (independent of the C++ task code) ----------------------------------
def _clampi(i, n):
    return min(max(int(i), 0), n - 1)


def _numpy_task(rodent, old, new, xmat_torso, cur_frame_old, sub_clip_frame_old):
    from types import SimpleNamespace
    fc, idx = rodent["fclip"], rodent["idx"]
    c = SimpleNamespace(**{k: np.asarray(getattr(fc, k), np.float64) for k in (
        "position", "quaternion", "joints", "body_positions", "velocity", "angular_velocity", "joints_velocity")})
    T = c.position.shape[0]
    take = lambda a, ids, axis: np.take(a, np.clip(ids, 0, a.shape[axis] - 1), axis=axis)  # JAX gather clamps
    f = _clampi(cur_frame_old, T)
    cur, sub = cur_frame_old + 1, sub_clip_frame_old + 1
    # _calculate_reward (rodent.py:266-316)
    com_ref = take(c.body_positions, [idx["com_idx"]], 1)[f, 0]
    rcom = np.exp(-100 * np.linalg.norm(new["subtree_com"] - com_ref))
    qvel_ref = np.hstack([c.velocity[f], c.angular_velocity[f], c.joints_velocity[f]])
    rvel = np.exp(-0.1 * np.linalg.norm(new["qvel"] - qvel_ref))
    ej = np.linalg.norm(c.joints[f] - old["qpos"][7:], ord=1)
    eb = np.linalg.norm(c.body_positions[f] - old["xpos"][idx["body_idxs"]], ord=1)  # matrix 1-norm (Q9)
    rtrunk = 1 - (0.5 * 1.0 * eb + 0.5 * ej) / 5.0
    qs, qt = new["qpos"][3:7] / np.linalg.norm(new["qpos"][3:7]), c.quaternion[f] / np.linalg.norm(c.quaternion[f])
    rquat = np.exp(-2 * np.abs(0.5 * np.arccos(min(1.0, 2 * float(qs @ qt) ** 2 - 1))))
    ract = -0.015 * np.mean(np.square(new["qfrc_actuator"]))
    app_ref = take(c.body_positions, idx["app_idx"], 1)[f].flatten()
    rapp = np.exp(-400 * np.linalg.norm(new["xpos"][idx["app_idx"]].flatten() - app_ref))
    healthy = 0.0 if new["qpos"][2] < 0.05 else 1.0
    healthy = 0.0 if new["qpos"][2] > 0.5 else healthy
    rcom, rvel, rapp, rtrunk, rquat, ract = rcom * 0.01, rvel * 0.01, rapp * 0.01, rtrunk * 0.01, rquat * 0.01, ract * 1e-4
    reward = rcom + rvel + rtrunk + rquat + ract + rapp
    done = max(1.0 - healthy, 1.0 if rtrunk < 0 else 0.0, 0.0 if sub < 10 else 1.0)
    # _get_obs / _get_traj (rodent.py:318-448)
    obs = np.hstack([new["qpos"], new["qvel"], new["qfrc_actuator"], new["xpos"][idx["end_eff_idx"]].flatten()])
    s = min(max(cur + 1, 0), T - 5)  # dynamic_slice_in_dim clamps the start
    w = slice(s, s + 5)
    bp = c.body_positions[w]
    diff = bp - new["xpos"][idx["body_idxs"]][None]
    traj = np.hstack([take(bp, idx["app_idx"], 1).flatten(), (diff @ xmat_torso).flatten(), diff.flatten(),
                      ((c.position[w] - new["qpos"][:3]) @ xmat_torso).flatten(),
                      take(c.joints[w] - new["qpos"][7:], idx["joint_idxs"], 1).flatten()])
    return dict(reward=reward, done=done, obs=obs, traj=traj, metrics=[rcom, rvel, rtrunk, rquat, ract, rapp, rtrunk],
                cur_frame=cur, sub_clip_frame=sub)


@pytest.mark.parametrize("cur0,sub0", [(3, 0), (120, 8), (243, 9), (248, 3), (400, 50)])
def test_env_step_matches_numpy_restatement_of_rodent_py(oracle_mod, rodent, cur0, sub0):
    B = 3
    qpos, qvel, start = start_states(rodent, B, seed=7)
    kw = dict(precision=64, dims=rodent["dims"], obs_size=232, traj_size=795)
    s0, _ = oracle_mod.reset(rodent["model_blob"], rodent["task_blob"], qpos, qvel, start, **kw)
    s0["cur_frame"][:] = cur0
    s0["sub_clip_frame"][:] = sub0
    rng = np.random.default_rng(8)
    action = rng.uniform(-1.2, 1.2, size=(B, 30))
    s1, o1 = oracle_mod.step(rodent["model_blob"], rodent["task_blob"], s0, action, **kw)
    # physics of the same call, no task logic (PipelineEnv.pipeline_step, rodent.py:181); ctrl clamped to +-1 inside
    p1, _ = oracle_mod.pipeline_step(rodent["model_blob"], s0, action, 5, precision=64, dims=rodent["dims"])
    for k in ("qpos", "qvel", "act", "xpos", "qfrc_actuator"):
        assert np.array_equal(p1[k], s1[k]), k
    for e in range(B):
        old = {k: s0[k][e] for k in ("qpos", "xpos")}
        new = {k: s1[k][e] for k in ("qpos", "qvel", "xpos", "subtree_com", "qfrc_actuator")}
        R = mjcf.quat_to_mat(s1["xquat"][e, 1])
        want = _numpy_task(rodent, old, new, R, cur0, sub0)
        assert s1["cur_frame"][e] == want["cur_frame"] and s1["sub_clip_frame"][e] == want["sub_clip_frame"]
        assert o1["done"][e] == want["done"]
        assert abs(o1["reward"][e] - want["reward"]) < 1e-12
        assert np.abs(o1["metrics"][e] - np.array(want["metrics"])).max() < 1e-12
        assert np.abs(o1["obs"][e] - want["obs"]).max() < 1e-12
        assert np.abs(o1["traj"][e] - want["traj"]).max() < 1e-12


def test_reset_matches_forward_and_reference_semantics(oracle_mod, rodent):
    B = 4
    qpos, qvel, start = start_states(rodent, B, seed=9)
    kw = dict(precision=64, dims=rodent["dims"], obs_size=232, traj_size=795)
    s0, o0 = oracle_mod.reset(rodent["model_blob"], rodent["task_blob"], qpos, qvel, start, **kw)
    d = _fwd(oracle_mod, rodent, qpos, qvel)
    assert np.array_equal(s0["xpos"], d["xpos"]) and np.array_equal(s0["qacc_warmstart"], d["qacc"])
    assert (s0["cur_frame"] == start).all() and (s0["sub_clip_frame"] == 0).all()
    assert (o0["reward"] == 0).all() and (o0["done"] == 0).all() and (o0["metrics"][:, :6] == 0).all()
    assert np.abs(s0["act"]).max() == 0  # mjx.make_data zeros
    # quaternion is normalised and written back into qpos by kinematics
    assert np.abs(np.linalg.norm(s0["qpos"][:, 3:7], axis=1) - 1).max() < 1e-12
    c = rodent["fclip"]
    for e in range(B):
        f = start[e]
        err = 0.5 * np.linalg.norm(c.body_positions[f] - s0["xpos"][e][rodent["idx"]["body_idxs"]], ord=1) \
            + 0.5 * np.linalg.norm(c.joints[f] - s0["qpos"][e, 7:], ord=1)
        assert abs(o0["metrics"][e, 6] - (1 - err / 5.0)) < 1e-12


def test_nan_guard(oracle_mod, rodent):
    qpos, qvel, start = start_states(rodent, 2, seed=10)
    kw = dict(precision=32, dims=rodent["dims"], obs_size=232, traj_size=795)
    s0, _ = oracle_mod.reset(rodent["model_blob"], rodent["task_blob"], qpos, qvel, start, **kw)
    s0["qvel"][1, 10] = np.nan
    s1, o1 = oracle_mod.step(rodent["model_blob"], rodent["task_blob"], s0, np.zeros((2, 30)), **kw)
    assert o1["done"][1] == 1.0 and np.isfinite(o1["obs"][1]).all() and np.isfinite(o1["reward"][1])
    assert o1["done"][0] == 0.0


def test_autoreset_quirk_q7_done_forever(oracle_mod, rodent):
    """sub_clip_frame only ever increments (rodent.py:185), so after sub_clip_length steps done stays 1."""
    qpos, qvel, start = start_states(rodent, 2, seed=11)
    kw = dict(precision=32, dims=rodent["dims"], obs_size=232, traj_size=795)
    s, _ = oracle_mod.reset(rodent["model_blob"], rodent["task_blob"], qpos, qvel, start, **kw)
    dones = []
    for _ in range(12):
        s, o = oracle_mod.step(rodent["model_blob"], rodent["task_blob"], s, np.zeros((2, 30)), **kw)
        dones.append(o["done"].copy())
    dones = np.array(dones)
    assert (dones[9:] == 1).all()
    assert (s["sub_clip_frame"] == 12).all()


def test_reference_dump(oracle_mod, rodent):
    """Real-reference golden vectors (tools/dump_reference.py, produced where jax + mujoco-mjx + brax exist).  The build
    image cannot run the reference, so the file is absent there and dynamics parity stays UNPINNED."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "rodent_reference_dump.npz")
    if not os.path.exists(path):
        pytest.skip("no real-reference dump available (jax / mujoco-mjx / brax are not installable here)")
    z = np.load(path)
    n = len(z["reward"])
    st = dict(qpos=z["qpos_in"], qvel=z["qvel_in"], act=z["act_in"], qacc_warmstart=z["warm_in"],
              xpos=np.zeros((n, 66, 3)), xquat=np.zeros((n, 66, 4)), subtree_com=np.zeros((n, 3)), qfrc_actuator=np.zeros((n, 73)),
              cur_frame=z["cur_frame_in"].astype(np.int32), sub_clip_frame=z["sub_clip_frame_in"].astype(np.int32))
    kw = dict(precision=32, dims=rodent["dims"], obs_size=232, traj_size=795)
    s1, o1 = oracle_mod.step(rodent["model_blob"], rodent["task_blob"], st, z["action"], **kw)
    for k in ("qpos", "qvel"):
        err = np.abs(s1[k] - z[k]).max() / np.abs(z[k]).max()
        assert err < 1e-4, (k, err)

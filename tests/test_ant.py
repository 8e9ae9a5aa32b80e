"""SURVEY section 8 row a10 / BASELINE configs[0]: AntTracking (envs/ant.py) -- ant.xml the brax way (jointless bodies
fused), Newton solver (1 iteration, 4 line-search iterations), eulerdamp off, 256 envs.

CPU: compiled model facts, the oracle's Newton branch against its CG branch (two solvers, one optimum), the oracle's ant
task logic against a line-by-line numpy restatement of envs/ant.py:167-309.
GPU: kernel vs oracle on the ant model (per-stage arrays incl. the Newton qacc, reset, env steps) and the configs[0]
workload: 256 envs, U(-1, 1) actions, 100 env steps, bounded by the oracle's own fp32-vs-fp64 spread."""
import numpy as np
import pytest

from conftest import pkg

mjcf = pkg("mjcf")


@pytest.fixture(scope="module")
def ant():
    antm, mb = pkg("envs.ant"), pkg("model_blob")
    model, clip = antm.packaged_ant()
    task_blob, state_size, traj_size = antm.ant_task_tables(model, clip)
    model_blob = mb.build_model_blob(model)
    return dict(model=model, clip=clip, model_blob=model_blob, task_blob=task_blob, dims=mb.read_dims(model_blob),
                obs_size=state_size, traj_size=traj_size)


def test_ant_model_is_the_fused_brax_load(ant):
    d, m = ant["dims"], ant["model"]
    # notebooks/environments_explore.ipynb: the brax-loaded ant has 10 bodies (the four jointless *_leg bodies are fused)
    assert (m.nbody, m.nq, m.nv, m.nu, m.na, m.njnt) == (10, 15, 14, 8, 0, 9)
    assert m.body_names[:3] == ["world", "torso", "aux_1"] and not any(n.endswith("_leg") for n in m.body_names)
    assert (d["ncon"], d["nlimit"], d["nefc"], d["eulerdamp"], d["solver"], d["iterations"], d["ls_iterations"]) == (4, 8, 24, 0, 2, 1, 4)
    assert abs(m.timestep - 0.01) < 1e-12 and (ant["obs_size"], ant["traj_size"]) == (29, 355)  # 355 + 29 = 384 = obs
    # torso = sphere + the four aux capsules of the fused bodies (density 5): mass of a sphere r = 0.25 plus 4 capsules
    cap = 5.0 * (np.pi * 0.08 ** 2 * np.hypot(0.2, 0.2) + 4.0 / 3.0 * np.pi * 0.08 ** 3)
    assert abs(m.arrays["body_mass"][1] - (5.0 * 4.0 / 3.0 * np.pi * 0.25 ** 3 + 4 * cap)) < 1e-9
    assert np.allclose(m.arrays["init_qpos"], [0, 0, 0.55, 1, 0, 0, 0, 0, 1, 0, -1, 0, -1, 0, 1])


def _states(ant, B, seed, press=0.0):
    m = ant["model"]
    rng = np.random.default_rng(seed)
    qpos = np.tile(m.arrays["init_qpos"], (B, 1)).astype(np.float32)
    qpos[:, 7:] += (0.15 * rng.standard_normal((B, 8))).astype(np.float32)
    qpos[:, 2] += rng.uniform(-0.03, 0.03, size=B).astype(np.float32) - press
    qvel = (0.3 * rng.standard_normal((B, m.nv))).astype(np.float32)
    return qpos, qvel, rng.integers(0, 95, size=B).astype(np.int32)


def test_oracle_newton_and_cg_reach_the_same_optimum(oracle_mod, ant):
    """The constraint problem is convex: Newton (dense Hessian, few iterations) and CG (many) must agree on qacc."""
    mb = pkg("model_blob")
    m = ant["model"]
    qpos, qvel, _ = _states(ant, 6, 7, press=0.06)  # feet pressed into the floor, some joints beyond their limits
    qpos[:, 8] += 0.4
    ctrl = np.random.default_rng(8).uniform(-1, 1, size=(6, 8))
    st = dict(qpos=qpos.astype(np.float64), qvel=qvel.astype(np.float64))
    out = {}
    for name, solver, iters in (("newton", mjcf.SOLVER_NEWTON, 30), ("cg", mjcf.SOLVER_CG, 300)):
        m.solver, m.iterations, m.ls_iterations = solver, iters, 50
        blob = mb.build_model_blob(m)
        out[name] = oracle_mod.forward_dump(blob, st, ctrl, precision=64, dims=mb.read_dims(blob))
    m.solver, m.iterations, m.ls_iterations = mjcf.SOLVER_NEWTON, 1, 4
    assert (out["newton"]["counters"][:, 2] > 0).all() and (out["newton"]["counters"][:, 3] > 0).any()
    scale = np.abs(out["cg"]["qacc"]).max()
    assert np.abs(out["newton"]["qacc"] - out["cg"]["qacc"]).max() < 1e-6 * scale
    assert np.abs(out["newton"]["qacc"] - out["newton"]["qacc_smooth"]).max() > 1e-2 * scale  # the constraints do act


def _numpy_ant_task(a, old, new, action, cur_frame_old):
    """envs/ant.py:167-309: reward + done from the OLD state and OLD frame, obs from the new state with the window at
    OLD cur_frame + 1 and the identity rotation (data.xmat[0])."""
    c = a["clip"]
    f64 = lambda x: np.asarray(x, np.float64)
    T = c.position.shape[0]
    f = min(max(int(cur_frame_old), 0), T - 1)
    rcom = np.exp(-100 * np.linalg.norm(old["subtree_com"] - f64(c.center_of_mass[f])))
    qvel_ref = np.hstack([f64(c.velocity[f]), f64(c.angular_velocity[f]), f64(c.joints_velocity[f])])
    rvel = np.exp(-0.1 * np.linalg.norm(old["qvel"] - qvel_ref))
    ej = np.mean(np.abs(f64(c.joints[f]) - old["qpos"][7:]))
    eb = np.mean(np.abs(f64(c.body_positions[f]) - old["xpos"]))
    rtrunk = 1 - (0.5 * 1.0 * eb + 0.5 * ej) / float(np.float32(0.9))  # the task blob holds the threshold as fp32
    qs = old["qpos"][3:7] / np.linalg.norm(old["qpos"][3:7])
    qt = f64(c.quaternion[f]) / np.linalg.norm(f64(c.quaternion[f]))
    rquat = np.exp(-2 * np.abs(0.5 * np.arccos(min(1.0, 2 * float(qs @ qt) ** 2 - 1))))
    ract = 0.01 * -0.015 * np.sum(np.square(action)) / len(action)
    healthy = 0.0 if old["qpos"][2] < float(np.float32(0.2)) else 1.0
    healthy = 0.0 if old["qpos"][2] > 1.0 else healthy
    w = (0.05, 0.01, 0.20, 0.01, 0.001)
    reward = w[0] * rcom + w[1] * rvel + w[2] * rtrunk + w[3] * rquat + w[4] * ract
    done = max(1.0 - healthy, 1.0 if rtrunk < 0 else 0.0)
    s = min(max(cur_frame_old + 1, 0), T - 5)
    win = slice(s, s + 5)
    diff = f64(c.body_positions[win]) - new["xpos"][None]
    obs = np.hstack([diff.flatten(), diff.flatten(), (f64(c.position[win]) - new["qpos"][:3]).flatten(),
                     (f64(c.joints[win]) - new["qpos"][7:]).flatten(), new["qpos"], new["qvel"]])
    return dict(reward=reward, done=done, obs=obs, metrics=[rcom, rvel, rtrunk, rquat, ract, 0.0, rtrunk], cur_frame=cur_frame_old + 1)


@pytest.mark.parametrize("cur0", [0, 57, 253, 400])
def test_oracle_ant_step_matches_numpy_restatement(oracle_mod, ant, cur0):
    a = ant
    B = 3
    qpos, qvel, start = _states(a, B, 1)
    kw = dict(precision=64, dims=a["dims"], obs_size=a["obs_size"], traj_size=a["traj_size"])
    s0, _ = oracle_mod.reset(a["model_blob"], a["task_blob"], qpos, qvel, start, **kw)
    s0["cur_frame"][:] = cur0
    if cur0 == 57:
        s0["qpos"][1, 2] = 0.15  # unhealthy on the OLD state -> done
    action = np.random.default_rng(2).uniform(-1.2, 1.2, size=(B, 8))  # beyond ctrlrange: ract reads the RAW action
    s1, o1 = oracle_mod.step(a["model_blob"], a["task_blob"], s0, action, **kw)
    for e in range(B):
        old = {k: s0[k][e] for k in ("qpos", "qvel", "xpos", "subtree_com")}
        new = {k: s1[k][e] for k in ("qpos", "qvel", "xpos")}
        want = _numpy_ant_task(a, old, new, action[e], cur0)
        assert s1["cur_frame"][e] == want["cur_frame"] and o1["done"][e] == want["done"]
        assert abs(o1["reward"][e] - want["reward"]) < 1e-9
        dm = np.abs(o1["metrics"][e] - np.array(want["metrics"]))
        assert np.delete(dm, 3).max() < 1e-12 and dm[3] < 1e-9
        got = np.hstack([o1["traj"][e], o1["obs"][e]])  # ant.py:300-309: one vector, reference block first
        assert got.shape == (384,) and np.abs(got - want["obs"]).max() < 1e-12
    if cur0 == 57:
        assert o1["done"][1] == 1.0


def test_oracle_ant_still_clip_rollout(oracle_mod, ant):
    a = ant
    m = a["model"]
    qpos = np.tile(m.arrays["init_qpos"], (2, 1)).astype(np.float64)
    kw = dict(precision=64, dims=a["dims"], obs_size=a["obs_size"], traj_size=a["traj_size"])
    s, o = oracle_mod.reset(a["model_blob"], a["task_blob"], qpos, np.zeros((2, 14)), np.zeros(2, np.int32), **kw)
    assert np.abs(o["metrics"][:, 6] - 1.0).max() < 1e-6 and (o["done"] == 0).all()  # the clip IS the reset pose
    for _ in range(40):  # the ant drops from z = 0.55 onto its feet and settles: contacts become active, nothing blows up
        s, o = oracle_mod.step(a["model_blob"], a["task_blob"], s, np.zeros((2, 8)), **kw)
    assert np.isfinite(s["qpos"]).all() and (o["stats"][:, 2] > 0).all() and (o["stats"][:, 0] == 5).all()  # 1 Newton iteration x 5
    assert 0.25 < s["qpos"][0, 2] < 0.6 and np.abs(s["qvel"]).max() < 0.5
    assert (o["reward"] > 0.1).all() and (o["done"] == 0).all()  # rtrunk ~ 1 carries weight 0.20


# ---- GPU ------------------------------------------------------------------------------------------------------------
def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.fixture(scope="module")
def gpu_ant(ant):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg("envs.ant").AntTracking(model=ant["model"], reference_clip=ant["clip"], device="cuda:0")


@pytest.mark.gpu
def test_gpu_ant_forward_stages_newton(gpu_ant, ant, oracle_mod):
    import torch
    a = ant
    B = 16
    qpos, qvel, _ = _states(a, B, 3, press=0.07)  # feet into the floor: active contacts; some ankles beyond their limits
    qpos[::2, 8] += 0.4
    rng = np.random.default_rng(4)
    ctrl = rng.uniform(-1.5, 1.5, size=(B, 8)).astype(np.float32)
    warm = rng.standard_normal((B, 14)).astype(np.float32)
    eng = gpu_ant.engine
    st = {k: torch.tensor(v, device="cuda") for k, v in dict(qpos=qpos, qvel=qvel, qacc_warmstart=warm).items()}
    g = oracle_mod.split_dump(a["dims"], eng.forward_dump(st, torch.tensor(ctrl, device="cuda")).cpu().numpy().astype(np.float64))
    ost = {k: v.astype(np.float64) for k, v in dict(qpos=qpos, qvel=qvel, qacc_warmstart=warm).items()}
    o32 = oracle_mod.forward_dump(a["model_blob"], ost, ctrl.astype(np.float64), precision=32, dims=a["dims"])
    o64 = oracle_mod.forward_dump(a["model_blob"], ost, ctrl.astype(np.float64), precision=64, dims=a["dims"])
    assert (o32["counters"][:, 2] > 0).all() and (o32["counters"][:, 3] > 0).any()
    for name, tol in dict(xpos=2e-6, xquat=2e-6, xipos=2e-6, xanchor=2e-6, xaxis=3e-6, cinert=3e-6, cdof=3e-6, crb=3e-6, qM=3e-6,
                          cvel=3e-6, qfrc_bias=2e-5, qfrc_actuator=1e-6, qfrc_smooth=2e-5, qacc_smooth=1e-4, con_dist=5e-6,
                          con_frame=3e-6).items():
        assert np.isfinite(g[name]).all(), name
        assert _rel(g[name], o32[name]) < tol, (name, _rel(g[name], o32[name]))
    # Newton iterations, active contacts and limits; the line-search count is not an invariant (a Newton step lands on
    # the minimiser of the current active set, where the bracketing tests compare rounding noise)
    assert np.array_equal(g["counters"][:, [0, 2, 3]], o32["counters"][:, [0, 2, 3]])
    for name in ("qacc", "qfrc_constraint"):  # one Newton step: bounded by the oracle's own fp32-vs-fp64 spread
        spread = _rel(o32[name], o64[name])
        assert _rel(g[name], o32[name]) < 10 * spread + 1e-4, (name, _rel(g[name], o32[name]), spread)


@pytest.mark.gpu
def test_gpu_ant_reset_and_steps(gpu_ant, ant, oracle_mod):
    import torch
    a = ant
    B = 10
    qpos, qvel, start = _states(a, B, 5)
    s0 = gpu_ant.reset_from(qpos, qvel, start)
    kw = dict(precision=32, dims=a["dims"], obs_size=a["obs_size"], traj_size=a["traj_size"])
    so, oo = oracle_mod.reset(a["model_blob"], a["task_blob"], qpos, qvel, start, **kw)
    assert s0.obs.shape == (B, 384) and gpu_ant.observation_size == 384
    assert _rel(s0.obs.cpu().numpy(), np.hstack([oo["traj"], oo["obs"]])) < 3e-6
    assert np.abs(s0.info["termination_error"].cpu().numpy() - oo["metrics"][:, 6]).max() < 1e-5
    rng = np.random.default_rng(6)
    to_np = lambda st: dict({k: v.cpu().numpy().astype(np.float64) for k, v in st.pipeline_state.items()},
                            cur_frame=st.info["cur_frame"].cpu().numpy(), sub_clip_frame=st.info["sub_clip_frame"].cpu().numpy())
    for it in range(8):
        act = rng.uniform(-1.2, 1.2, size=(B, 8)).astype(np.float32)
        old = to_np(s0)
        o32s, o32o = oracle_mod.step(a["model_blob"], a["task_blob"], old, act.astype(np.float64), **kw)
        o64s, _ = oracle_mod.step(a["model_blob"], a["task_blob"], old, act.astype(np.float64), **dict(kw, precision=64))
        s1 = gpu_ant.step(s0, torch.tensor(act, device="cuda"))
        new = to_np(s1)
        assert np.array_equal(new["cur_frame"], o32s["cur_frame"])
        assert np.array_equal(s1.done.cpu().numpy(), o32o["done"])  # reward / done read the OLD state only
        assert np.abs(s1.reward.cpu().numpy() - o32o["reward"]).max() < 1e-5
        assert np.abs(np.stack([s1.metrics[k].cpu().numpy() for k in ("rcom", "rvel", "rtrunk", "rquat", "ract")], 1)
                      - o32o["metrics"][:, :5]).max() < 1e-5
        for e in range(B):
            want = _numpy_ant_task(a, {k: old[k][e] for k in ("qpos", "qvel", "xpos", "subtree_com")},
                                   {k: new[k][e] for k in ("qpos", "qvel", "xpos")}, act[e].astype(np.float64), int(old["cur_frame"][e]))
            assert np.abs(s1.obs[e].cpu().numpy() - want["obs"]).max() < 3e-6 * max(1.0, np.abs(want["obs"]).max())
        for k in ("qpos", "qvel"):
            eg = np.abs(new[k] - o32s[k]).max(1) / (np.abs(o32s[k]).max() + 1e-30)
            eo = np.abs(o32s[k] - o64s[k]).max(1) / (np.abs(o32s[k]).max() + 1e-30)
            assert np.median(eg) < 10 * np.median(eo) + 1e-4, (it, k, np.median(eg), np.median(eo))
        s0 = s1


@pytest.mark.gpu
def test_gpu_ant_config0_workload(gpu_ant, ant, oracle_mod):
    """BASELINE configs[0]: 256 envs, U(-1, 1) actions, seed 0, 100 env steps.  The ant tumbles chaotically under random
    torques, so the trajectories are compared teacher-forced every 10 steps and end to end on the invariants."""
    import torch
    a = ant
    B, K = 256, 100
    s = gpu_ant.reset(batch_size=B)
    rng = np.random.default_rng(0)
    kw = dict(precision=32, dims=a["dims"], obs_size=a["obs_size"], traj_size=a["traj_size"])
    to_np = lambda st: dict({k: v.cpu().numpy().astype(np.float64) for k, v in st.pipeline_state.items()},
                            cur_frame=st.info["cur_frame"].cpu().numpy(), sub_clip_frame=st.info["sub_clip_frame"].cpu().numpy())
    worst = 0.0
    for it in range(K):
        act = rng.uniform(-1, 1, size=(B, 8)).astype(np.float32)
        if it % 10 == 0:
            old = to_np(s)
            o32s, o32o = oracle_mod.step(a["model_blob"], a["task_blob"], old, act.astype(np.float64), nthreads=8, **kw)
        s = gpu_ant.step(s, torch.tensor(act, device="cuda"))
        if it % 10 == 0:
            new = to_np(s)
            for k in ("qpos", "qvel"):
                err = np.abs(new[k] - o32s[k]).max(1) / (np.abs(o32s[k]).max() + 1e-30)
                worst = max(worst, float(np.median(err)))
                assert np.median(err) < 1e-4 and (err < 1e-2).mean() > 0.98, (it, k, np.median(err), err.max())
            assert np.array_equal(s.done.cpu().numpy(), o32o["done"])
            assert np.abs(s.reward.cpu().numpy() - o32o["reward"]).max() < 1e-5
    q = s.pipeline_state["qpos"]
    assert torch.isfinite(q).all() and torch.isfinite(s.obs).all()
    assert (s.info["cur_frame"].cpu().numpy() == K).all()
    assert (torch.linalg.norm(q[:, 3:7], dim=1) - 1).abs().max() < 1e-5
    assert torch.equal(s.obs[:, 355:370], torch.nan_to_num(q)) and torch.equal(s.obs[:, 370:], s.pipeline_state["qvel"])

"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances.  BASELINE.json asks for qpos/qvel within 1e-4 relative after 100 physics steps and rewards
within 1e-5, with frame indices / done flags bit-exact.  Integer outputs are tested bit-exact.  For the
floating-point state the bound is only meaningful where the dynamics are not chaotic at fp32 resolution:
  * contact-free flight and settled contact: 1e-4 relative after 100 substeps (tested as stated);
  * clip start states (tail 3-5 cm inside the floor, 6 CG iterations, stiff solref): the fp32 and fp64
    builds of the ORACLE ITSELF differ by O(1e-2) after one env step; there the kernel must stay within a
    small multiple of that fp32-vs-fp64 spread, and the per-stage arrays of one forward pass, which are
    not chaotic, must match to fp32 rounding.
The reward / obs / traj / termination logic is checked to 1e-6 against the oracle on the kernel's own
post-step state, which removes the physics sensitivity from the comparison."""
import numpy as np
import pytest

from conftest import pkg, start_states

pytestmark = pytest.mark.gpu

STATE_KEYS = ("qpos", "qvel", "act", "qacc_warmstart", "xpos", "xquat", "subtree_com", "qfrc_actuator")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def to_np(state):
    d = {k: v.cpu().numpy().astype(np.float64) for k, v in state.pipeline_state.items()}
    d["cur_frame"] = state.info["cur_frame"].cpu().numpy()
    d["sub_clip_frame"] = state.info["sub_clip_frame"].cpu().numpy()
    return d


def okw(rodent, precision=32):
    return dict(precision=precision, dims=rodent["dims"], obs_size=rodent["obs_size"], traj_size=rodent["traj_size"])


def test_forward_stages_match_oracle(gpu_env, rodent, oracle_mod):
    import torch
    B = 16
    qpos, qvel, _ = start_states(rodent, B, seed=0)
    rng = np.random.default_rng(1)
    act = rng.uniform(0, 1, size=(B, 30)).astype(np.float32)
    warm = rng.standard_normal((B, 73)).astype(np.float32)
    ctrl = rng.uniform(-1.5, 1.5, size=(B, 30)).astype(np.float32)
    eng = gpu_env.engine
    st = {k: torch.tensor(v, device="cuda") for k, v in dict(qpos=qpos, qvel=qvel, act=act, qacc_warmstart=warm).items()}
    dump = eng.forward_dump(st, torch.tensor(ctrl, device="cuda")).cpu().numpy().astype(np.float64)
    g = oracle_mod.split_dump(rodent["dims"], dump)
    ost = {k: v.astype(np.float64) for k, v in dict(qpos=qpos, qvel=qvel, act=act, qacc_warmstart=warm).items()}
    o32 = oracle_mod.forward_dump(rodent["model_blob"], ost, ctrl.astype(np.float64), precision=32, dims=rodent["dims"])
    tol = dict(xpos=2e-6, xquat=2e-6, xipos=2e-6, xanchor=2e-6, xaxis=3e-6, cinert=2e-6, cdof=3e-6, crb=3e-6, qM=3e-6, cvel=3e-6,
               qfrc_passive=2e-6, qfrc_bias=1e-5, qfrc_actuator=1e-6, act_dot=1e-6, qfrc_smooth=1e-5, qacc_smooth=5e-5,
               con_dist=3e-6, con_pos=3e-6, con_frame=3e-6, efc_pos=1e-5, efc_D=1e-4, efc_aref=1e-4, qacc=2e-3,
               qfrc_constraint=2e-3)
    checked = 0
    for name, t in tol.items():
        ga, oa = g[name], o32[name]
        if name in ("efc_pos", "efc_D", "efc_aref"):  # the kernel writes active rows only (inactive rows are exactly inert)
            m = np.isfinite(ga)
            assert m.any()
            assert np.array_equal(m, o32["efc_pos"] < 0), name  # exactly the oracle's active rows, for each of the three arrays
            err = float(np.abs(ga[m] - oa[m]).max() / (np.abs(oa[m]).max() + 1e-30))
        else:
            assert np.isfinite(ga).all(), name
            err = rel(ga, oa)
        assert err < t, (name, err)
        checked += 1
    sc = g["subtree_com"][:, 1]
    assert rel(sc, o32["subtree_com"][:, 1]) < 2e-6
    # active sets are bit-exact (contact pair sets, limit rows)
    assert np.array_equal(g["counters"][:, 2:4], o32["counters"][:, 2:4])
    act_gpu = np.isfinite(g["efc_pos"])
    act_o = o32["efc_pos"] < 0
    assert np.array_equal(act_gpu, act_o)
    assert checked == len(tol)


def test_rodent_newton_solver_matches_oracle(rodent, oracle_mod):
    """The Newton branch on the big model (dense 73 x 73 Hessian in shared memory, inertia streamed from the workspace):
    `solver: newton` is a legal setting of the reference's rodent config (envs/rodent.py:55-63).  One forward pass and a
    few pipeline steps against the oracle's Newton branch, bounded by the oracle's own fp32-vs-fp64 spread."""
    import copy
    import torch
    mb, libm, mj = pkg("model_blob"), pkg("_lib"), pkg("mjcf")
    model = copy.deepcopy(rodent["model"])
    model.solver, model.iterations, model.ls_iterations = mj.SOLVER_NEWTON, 2, 4
    blob = mb.build_model_blob(model)
    dims = mb.read_dims(blob)
    eng = libm.Engine(blob, None, device="cuda:0")
    assert 1 <= eng.envs_per_cta < 14  # the Hessian costs residency: supported, not the fast path
    B = 12
    qpos, qvel, _ = start_states(rodent, B, seed=51)
    qpos[:, 2] -= 0.01  # belly into the floor: contacts act
    rng = np.random.default_rng(52)
    ctrl = rng.uniform(-1, 1, size=(B, 30)).astype(np.float32)
    st = dict(qpos=torch.tensor(qpos, device="cuda"), qvel=torch.tensor(qvel, device="cuda"))
    g = oracle_mod.split_dump(dims, eng.forward_dump(st, torch.tensor(ctrl, device="cuda")).cpu().numpy().astype(np.float64))
    ost = dict(qpos=qpos.astype(np.float64), qvel=qvel.astype(np.float64))
    o32 = oracle_mod.forward_dump(blob, ost, ctrl.astype(np.float64), precision=32, dims=dims)
    o64 = oracle_mod.forward_dump(blob, ost, ctrl.astype(np.float64), precision=64, dims=dims)
    assert (o32["counters"][:, 2] > 0).all() and (o32["counters"][:, 0] == 2).all()
    assert np.array_equal(g["counters"][:, [0, 2, 3]], o32["counters"][:, [0, 2, 3]])
    for name in ("qacc", "qfrc_constraint"):
        spread = rel(o32[name], o64[name])
        assert np.isfinite(g[name]).all() and rel(g[name], o32[name]) < 10 * spread + 1e-3, (name, rel(g[name], o32[name]), spread)
    assert rel(g["qacc_smooth"], o32["qacc_smooth"]) < 1e-4 and rel(g["qM"], o32["qM"]) < 5e-6


def test_reset_matches_oracle(gpu_env, rodent, oracle_mod):
    B = 33
    qpos, qvel, start = start_states(rodent, B, seed=2)
    s = gpu_env.reset_from(qpos, qvel, start)
    so, oo = oracle_mod.reset(rodent["model_blob"], rodent["task_blob"], qpos, qvel, start, **okw(rodent))
    g = to_np(s)
    assert np.array_equal(g["cur_frame"], so["cur_frame"]) and np.array_equal(g["sub_clip_frame"], so["sub_clip_frame"])
    for k, t in dict(qpos=1e-6, qvel=1e-6, xpos=2e-6, xquat=2e-6, subtree_com=2e-6, qfrc_actuator=1e-6, act=1e-6, qacc_warmstart=2e-3).items():
        assert rel(g[k], so[k]) < t or np.abs(so[k]).max() == 0 and np.abs(g[k]).max() == 0, k
    assert rel(s.obs.cpu().numpy(), oo["obs"]) < 1e-6
    assert rel(s.info["traj"].cpu().numpy(), oo["traj"]) < 2e-6
    assert np.abs(s.info["termination_error"].cpu().numpy() - oo["metrics"][:, 6]).max() < 1e-5
    assert (s.reward.cpu().numpy() == 0).all() and (s.done.cpu().numpy() == 0).all()


def test_task_logic_on_kernel_state(gpu_env, rodent, oracle_mod):
    """reward / obs / traj / done of the kernel == the oracle's task logic evaluated on the kernel's OWN physics result
    (the oracle is stepped from the pre-step state, then its post-physics state is overwritten with the kernel's)."""
    import torch
    B = 24
    qpos, qvel, start = start_states(rodent, B, seed=3)
    s0 = gpu_env.reset_from(qpos, qvel, start)
    rng = np.random.default_rng(4)
    for it, (cur0, sub0) in enumerate([(None, None), (243, 8), (248, 3), (300, 20)]):
        if cur0 is not None:
            s0.info["cur_frame"].fill_(cur0)
            s0.info["sub_clip_frame"].fill_(sub0)
        a = rng.uniform(-1, 1, size=(B, 30)).astype(np.float32)
        old = to_np(s0)
        s1 = gpu_env.step(s0, torch.tensor(a, device="cuda"))
        new = to_np(s1)
        # numpy restatement of envs/rodent.py from tests/test_oracle.py on (old, new)
        from test_oracle import _numpy_task
        mj = pkg("mjcf")
        for e in range(B):
            want = _numpy_task(rodent, {k: old[k][e] for k in ("qpos", "xpos")},
                               {k: new[k][e] for k in ("qpos", "qvel", "xpos", "subtree_com", "qfrc_actuator")},
                               mj.quat_to_mat(new["xquat"][e, 1]), int(old["cur_frame"][e]), int(old["sub_clip_frame"][e]))
            assert new["cur_frame"][e] == want["cur_frame"] and new["sub_clip_frame"][e] == want["sub_clip_frame"]
            assert float(s1.done[e]) == want["done"]
            # north star: rewards within 1e-5.  (rquat = exp(-arccos(2 (q.q')^2 - 1)) is ill-conditioned in fp32 when the root
            # orientation tracks the reference: an ulp of the dot product moves arccos by ~5e-4, the reward by ~5e-6)
            assert abs(float(s1.reward[e]) - want["reward"]) < 1e-5
            m = np.array([float(s1.metrics[k][e]) for k in pkg("envs.rodent").METRIC_KEYS])
            dm = np.abs(m - np.array(want["metrics"]))
            assert np.delete(dm, 3).max() < 1e-6 and dm[3] < 1e-5
            assert np.abs(s1.obs[e].cpu().numpy() - want["obs"]).max() < 1e-6 * max(1.0, np.abs(want["obs"]).max())
            assert np.abs(s1.info["traj"][e].cpu().numpy() - want["traj"]).max() < 2e-6 * max(1.0, np.abs(want["traj"]).max())
        s0 = s1


def test_step_teacher_forced_within_fp32_spread(gpu_env, rodent, oracle_mod):
    """Violent regime (clip start states): kernel-vs-fp32-oracle error is bounded by a multiple of the oracle's own
    fp32-vs-fp64 spread; integer outputs are exact."""
    import torch
    B = 16
    qpos, qvel, start = start_states(rodent, B, seed=5)
    s = gpu_env.reset_from(qpos, qvel, start)
    rng = np.random.default_rng(6)
    ratios = []
    for it in range(10):
        a = rng.uniform(-1, 1, size=(B, 30)).astype(np.float32)
        st = to_np(s)
        so, oo = oracle_mod.step(rodent["model_blob"], rodent["task_blob"], st, a.astype(np.float64), **okw(rodent, 32))
        s64, _ = oracle_mod.step(rodent["model_blob"], rodent["task_blob"], st, a.astype(np.float64), **okw(rodent, 64))
        s = gpu_env.step(s, torch.tensor(a, device="cuda"))
        g = to_np(s)
        assert np.array_equal(g["cur_frame"], so["cur_frame"]) and np.array_equal(g["sub_clip_frame"], so["sub_clip_frame"])
        assert np.isfinite(g["qpos"]).all()
        # contact / limit activity summed over the 5 substeps: equal except where the trajectories already diverged
        stats = s.info["solver_stats"].cpu().numpy()
        assert (np.abs(stats[:, 2] - oo["stats"][:, 2]) <= 2).mean() > 0.8
        for k in ("qpos", "qvel"):
            eg = np.abs(g[k] - so[k]).max(1) / (np.abs(so[k]).max() + 1e-30)
            eo = np.abs(so[k] - s64[k]).max(1) / (np.abs(so[k]).max() + 1e-30)
            ratios.append(np.median(eg) / (np.median(eo) + 1e-9))
            assert np.median(eg) < 10 * np.median(eo) + 1e-4, (it, k, np.median(eg), np.median(eo))
    assert np.median(ratios) < 3.0, ratios


def _flight_state(rodent, B, seed):
    m = rodent["model"]
    rng = np.random.default_rng(seed)
    qpos, qvel, _ = start_states(rodent, B, seed=seed)
    qpos[:, 2] = 0.45  # lifted clear of the floor (no contact within 100 substeps of free fall: drop ~0.2 m)
    qvel = (0.05 * rng.standard_normal(qvel.shape)).astype(np.float32)
    return qpos, qvel


def test_100_substeps_contact_free_within_1e4(gpu_env, rodent, oracle_mod):
    """BASELINE.json tolerance as stated: qpos / qvel within 1e-4 relative after 100 physics steps, in the regime where
    that is meaningful: free flight with zero control (joint limits, springs, dampers, gravity and the actuator filter
    active; ~5 active limit rows per substep).  There the oracle's own fp32 and fp64 builds agree to ~3e-6.  With random
    full-range torques the rodent reaches hundreds of rad/s within 0.2 s and fp32-vs-fp64 of the oracle is O(1)."""
    import torch
    B = 12
    qpos, qvel = _flight_state(rodent, B, 7)
    ctrl = np.zeros((B, 30), dtype=np.float32)
    eng = gpu_env.engine
    st = dict(qpos=torch.tensor(qpos, device="cuda"), qvel=torch.tensor(qvel, device="cuda"))
    out = eng.alloc_state(B)
    stats = torch.zeros(B, 4, dtype=torch.int32, device="cuda")
    eng.pipeline_step(st, torch.tensor(ctrl, device="cuda"), out, 100, stats)
    ost = dict(qpos=qpos.astype(np.float64), qvel=qvel.astype(np.float64))
    o32, st32 = oracle_mod.pipeline_step(rodent["model_blob"], ost, ctrl.astype(np.float64), 100, precision=32, dims=rodent["dims"])
    o64, _ = oracle_mod.pipeline_step(rodent["model_blob"], ost, ctrl.astype(np.float64), 100, precision=64, dims=rodent["dims"])
    assert (stats.cpu().numpy()[:, 2] == 0).all() and (st32[:, 2] == 0).all()  # no contacts on this path
    assert np.array_equal(stats.cpu().numpy()[:, 3] > 0, st32[:, 3] > 0)
    for k in ("qpos", "qvel"):
        eg, eo = rel(out[k].cpu().numpy(), o32[k]), rel(o32[k], o64[k])
        assert eo < 2e-5, (k, eo)  # the regime is well conditioned
        assert eg < 1e-4, (k, eg, eo)
    assert np.abs(out["act"].cpu().numpy()).max() == 0


def test_settled_contact_step(gpu_env, rodent, oracle_mod):
    """Resting contact (state settled by the fp64 oracle for 600 substeps): one env step agrees far better than in the
    violent regime, and the active contact set is identical."""
    import torch
    B = 8
    c = rodent["fclip"]
    fr = np.arange(0, 80, 10)
    qpos = np.hstack([c.position[fr], c.quaternion[fr], c.joints[fr]]).astype(np.float64)
    st = dict(qpos=qpos, qvel=np.zeros((B, 73)))
    settled, _ = oracle_mod.pipeline_step(rodent["model_blob"], st, None, 600, precision=64, dims=rodent["dims"])
    s_in = {k: np.ascontiguousarray(settled[k].astype(np.float32)) for k in ("qpos", "qvel", "act", "qacc_warmstart")}
    eng = gpu_env.engine
    tin = {k: torch.tensor(v, device="cuda") for k, v in s_in.items()}
    out = eng.alloc_state(B)
    stats = torch.zeros(B, 4, dtype=torch.int32, device="cuda")
    eng.pipeline_step(tin, None, out, 5, stats)
    oin = {k: v.astype(np.float64) for k, v in s_in.items()}
    o32, st32 = oracle_mod.pipeline_step(rodent["model_blob"], oin, None, 5, precision=32, dims=rodent["dims"])
    o64, _ = oracle_mod.pipeline_step(rodent["model_blob"], oin, None, 5, precision=64, dims=rodent["dims"])
    g = stats.cpu().numpy()
    assert (g[:, 2] > 0).all()
    assert (np.abs(g[:, 2] - st32[:, 2]) <= 1).all() and (np.abs(g[:, 3] - st32[:, 3]) <= 1).all()
    eq, eo = rel(out["qpos"].cpu().numpy(), o32["qpos"]), rel(o32["qpos"], o64["qpos"])
    assert eq < 1e-4 + 5 * eo, (eq, eo)
    ev, eov = rel(out["qvel"].cpu().numpy(), o32["qvel"]), rel(o32["qvel"], o64["qvel"])
    assert ev < 5e-3 + 5 * eov, (ev, eov)


def test_100_substeps_settled_contact_bounded_actions(gpu_env, rodent, oracle_mod):
    """the 100-substep comparison in SETTLED CONTACT with BOUNDED random actions (U(-0.3, 0.3) held per
    env step, ~6 active contacts per substep), free-running and teacher-forced, the kernel-vs-fp32-oracle error beside the
    oracle's own fp32-vs-fp64 spread (curves: tools/parity_curve.py -> profiles/r02_parity_curve.json).

    Finding the test encodes: contact + actuation is chaotic at fp32 resolution -- the ORACLE's fp32 and fp64 builds are
    ~6e-2 apart in qpos after 100 substeps (1e-4 after the first env step) -- so "1e-4 after 100 steps" cannot hold for any
    fp32 implementation in this regime, MJX included.  What can and does hold: the kernel is no further from the fp32 oracle
    than the fp32 oracle is from its fp64 twin (ratio of medians ~1 over the run; bound 2.5; single steps 10), free-running and, free-running and
    teacher-forced; frame counters and done flags are exact; the teacher-forced reward differs by ~1e-5 (median)."""
    import torch
    B, T, amp = 16, 20, 0.3
    c = rodent["fclip"]
    fr = (np.arange(B) * 11) % 200
    qpos = np.hstack([c.position[fr], c.quaternion[fr], c.joints[fr]]).astype(np.float64)
    settled, _ = oracle_mod.pipeline_step(rodent["model_blob"], dict(qpos=qpos, qvel=np.zeros((B, 73))), None, 600, precision=64,
                                          dims=rodent["dims"])
    qp, qv = settled["qpos"].astype(np.float32), settled["qvel"].astype(np.float32)
    acts = np.random.default_rng(21).uniform(-amp, amp, size=(T, B, 30)).astype(np.float32)
    s = gpu_env.reset_from(qp, qv, fr.astype(np.int32))
    mb_, tb_ = rodent["model_blob"], rodent["task_blob"]
    s32, _ = oracle_mod.reset(mb_, tb_, qp.astype(np.float64), qv.astype(np.float64), fr.astype(np.int32), **okw(rodent, 32))
    s64, _ = oracle_mod.reset(mb_, tb_, qp.astype(np.float64), qv.astype(np.float64), fr.astype(np.int32), **okw(rodent, 64))
    med = lambda a, b: float(np.median(np.abs(a - b).reshape(B, -1).max(1) / (np.abs(b).max() + 1e-30)))
    ratios_tf, ratios_free, rew_tf = [], [], []
    for t in range(T):
        a64 = acts[t].astype(np.float64)
        k_in = to_np(s)
        tf32, tfo32 = oracle_mod.step(mb_, tb_, k_in, a64, **okw(rodent, 32))
        tf64, _ = oracle_mod.step(mb_, tb_, k_in, a64, **okw(rodent, 64))
        s = gpu_env.step(s, torch.tensor(acts[t], device="cuda"))
        g = to_np(s)
        s32, _ = oracle_mod.step(mb_, tb_, s32, a64, **okw(rodent, 32))
        s64, _ = oracle_mod.step(mb_, tb_, s64, a64, **okw(rodent, 64))
        assert np.array_equal(g["cur_frame"], tf32["cur_frame"]) and np.array_equal(s.done.cpu().numpy(), tfo32["done"]), t
        assert np.isfinite(g["qpos"]).all()
        e_tf, s_tf = med(g["qpos"], tf32["qpos"]), med(tf32["qpos"], tf64["qpos"])
        e_fr, s_fr = med(g["qpos"], s32["qpos"]), med(s32["qpos"], s64["qpos"])
        assert e_tf < 10 * s_tf + 1e-4, (t, e_tf, s_tf)  # single steps: medians over 16 envs are noisy (same bound as the clip-start test)
        assert e_fr < 10 * s_fr + 1e-4, (t, e_fr, s_fr)
        ratios_tf.append(e_tf / (s_tf + 1e-9)); ratios_free.append(e_fr / (s_fr + 1e-9))
        rew_tf.append(float(np.median(np.abs(s.reward.cpu().numpy() - tfo32["reward"]))))
    assert np.median(ratios_tf) < 2.5 and np.median(ratios_free) < 2.5, (ratios_tf, ratios_free)  # over the 20 steps: ~1
    assert np.median(rew_tf) < 1e-4, rew_tf
    assert float(s.info["solver_stats"][:, 2].float().mean()) / 5 > 2.0  # still in contact after 100 substeps


def test_ieee_div_sqrt_ab(gpu_env, rodent, oracle_mod):
    """VERDICT r1 item 1e.  The product kernels are built with -prec-div=false -prec-sqrt=false and use rsqrtf in the
    factorisation; XLA (the reference) divides in IEEE.  A/B against the twin library built with IEEE division / square root
    (libvnl_b200_ieee.so, same sources): (i) per-stage arrays of one forward pass: the two builds agree with each other far
    inside the tolerance either has against the oracle; (ii) one env step from clip start states: the approximate build is
    not further from the fp32 oracle than the IEEE build is (the difference between them is below the oracle's own
    fp32-vs-fp64 spread).  Cost of the IEEE build in time: tools/gpu_ab.sh (profiles/README.md)."""
    import torch
    libm = pkg("_lib")
    ieee = libm.Engine(rodent["model_blob"], rodent["task_blob"], lib_path=libm.IEEE_LIB_PATH)
    fast = gpu_env.engine
    B = 32
    qpos, qvel, start = start_states(rodent, B, seed=31)
    rng = np.random.default_rng(32)
    st = {k: torch.tensor(v, device="cuda") for k, v in dict(qpos=qpos, qvel=qvel, act=rng.uniform(0, 1, (B, 30)).astype(np.float32),
                                                             qacc_warmstart=rng.standard_normal((B, 73)).astype(np.float32)).items()}
    ctrl = torch.tensor(rng.uniform(-1, 1, (B, 30)).astype(np.float32), device="cuda")
    df = oracle_mod.split_dump(rodent["dims"], fast.forward_dump(st, ctrl).cpu().numpy().astype(np.float64))
    di = oracle_mod.split_dump(rodent["dims"], ieee.forward_dump(st, ctrl).cpu().numpy().astype(np.float64))
    ost = {k: v.cpu().numpy().astype(np.float64) for k, v in st.items()}
    o32 = oracle_mod.forward_dump(rodent["model_blob"], ost, ctrl.cpu().numpy().astype(np.float64), precision=32, dims=rodent["dims"])
    for name, tol in (("xpos", 1e-6), ("qM", 1e-6), ("qfrc_bias", 3e-6), ("qacc_smooth", 2e-5), ("qacc", 1e-3)):
        ab, fo, io = rel(df[name], di[name]), rel(df[name], o32[name]), rel(di[name], o32[name])
        assert ab < tol, (name, ab)
        assert fo < 3 * io + tol, (name, fo, io)  # the approximate build is as close to the oracle as the IEEE build
    # (ii) one env step
    def one_step(eng):
        sin = dict(qpos=torch.tensor(qpos, device="cuda"), qvel=torch.tensor(qvel, device="cuda"), cur_frame=torch.tensor(start, device="cuda"))
        s0, o0 = eng.alloc_state(B), eng.alloc_outputs(B)
        eng.reset(sin, s0, o0)
        s1, o1 = eng.alloc_state(B), eng.alloc_outputs(B)
        eng.step(s0, ctrl, s1, o1)
        torch.cuda.synchronize()
        return s0, s1, o1
    f0, f1, fo1 = one_step(fast)
    i0, i1, io1 = one_step(ieee)
    k_in = {k: f0[k].cpu().numpy().astype(np.float64) for k in STATE_KEYS}
    k_in["cur_frame"], k_in["sub_clip_frame"] = f0["cur_frame"].cpu().numpy(), f0["sub_clip_frame"].cpu().numpy()
    s32, _ = oracle_mod.step(rodent["model_blob"], rodent["task_blob"], k_in, ctrl.cpu().numpy().astype(np.float64), **okw(rodent, 32))
    s64, _ = oracle_mod.step(rodent["model_blob"], rodent["task_blob"], k_in, ctrl.cpu().numpy().astype(np.float64), **okw(rodent, 64))
    med = lambda a, b: float(np.median(np.abs(a - b).reshape(B, -1).max(1) / (np.abs(b).max() + 1e-30)))
    spread = med(s32["qpos"], s64["qpos"])
    ab = med(f1["qpos"].cpu().numpy().astype(np.float64), i1["qpos"].cpu().numpy().astype(np.float64))
    e_fast, e_ieee = med(f1["qpos"].cpu().numpy().astype(np.float64), s32["qpos"]), med(i1["qpos"].cpu().numpy().astype(np.float64), s32["qpos"])
    assert ab < 3 * spread + 1e-5, (ab, spread)
    assert e_fast < 3 * e_ieee + 1e-5, (e_fast, e_ieee, spread)
    assert torch.equal(fo1["done"], io1["done"]) and torch.equal(f1["cur_frame"], i1["cur_frame"])
    print("ieee A/B: |fast - ieee| %.2e, |fast - o32| %.2e, |ieee - o32| %.2e, oracle fp32-vs-fp64 %.2e" % (ab, e_fast, e_ieee, spread))


def _run_steps(env, qpos, qvel, start, actions):
    import torch
    s = env.reset_from(qpos, qvel, start)
    for a in actions:
        s = env.step(s, torch.tensor(a, device="cuda"))
    torch.cuda.synchronize()
    return s


def test_batch_independence_and_determinism(gpu_env, rodent):
    """Env i's result does not depend on the batch it is stepped in, nor on the run (fixed reduction trees, no
    atomics): B = 1, 7 and 300 give bit-identical rows.  This is also the multi-GPU sharding contract."""
    B = 300
    qpos, qvel, start = start_states(rodent, B, seed=9)
    acts = np.random.default_rng(10).uniform(-1, 1, size=(3, B, 30)).astype(np.float32)
    big = _run_steps(gpu_env, qpos, qvel, start, acts)
    again = _run_steps(gpu_env, qpos, qvel, start, acts)
    for k in STATE_KEYS:
        assert np.array_equal(big.pipeline_state[k].cpu().numpy(), again.pipeline_state[k].cpu().numpy()), k
    assert np.array_equal(big.obs.cpu().numpy(), again.obs.cpu().numpy())
    for lo, hi in ((0, 1), (5, 12), (293, 300)):
        sm = _run_steps(gpu_env, qpos[lo:hi], qvel[lo:hi], start[lo:hi], acts[:, lo:hi])
        for k in STATE_KEYS:
            assert np.array_equal(sm.pipeline_state[k].cpu().numpy(), big.pipeline_state[k].cpu().numpy()[lo:hi]), (k, lo)
        assert np.array_equal(sm.reward.cpu().numpy(), big.reward.cpu().numpy()[lo:hi])
        assert np.array_equal(sm.info["traj"].cpu().numpy(), big.info["traj"].cpu().numpy()[lo:hi])


def test_streams_graph_capture_and_aliasing(gpu_env, rodent):
    """The entry points only enqueue on the caller's stream: they run on a non-default stream, are CUDA-graph
    capturable (no allocation / sync inside), and accept in == out buffers."""
    import torch
    B = 64
    eng = gpu_env.engine
    qpos, qvel, start = start_states(rodent, B, seed=11)
    s0 = gpu_env.reset_from(qpos, qvel, start)
    a = torch.tensor(np.random.default_rng(12).uniform(-1, 1, size=(B, 30)).astype(np.float32), device="cuda")
    st_in = dict(s0.pipeline_state); st_in["cur_frame"] = s0.info["cur_frame"]; st_in["sub_clip_frame"] = s0.info["sub_clip_frame"]
    ref_out, ref_o = eng.alloc_state(B), eng.alloc_outputs(B)
    eng.step(st_in, a, ref_out, ref_o)
    torch.cuda.synchronize()
    # non-default stream
    side = torch.cuda.Stream()
    o2, oo2 = eng.alloc_state(B), eng.alloc_outputs(B)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eng.step(st_in, a, o2, oo2)
    side.synchronize()
    for k in STATE_KEYS:
        assert torch.equal(o2[k], ref_out[k]), k
    # graph capture + replay
    o3, oo3 = eng.alloc_state(B), eng.alloc_outputs(B)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        eng.step(st_in, a, o3, oo3)
    gr.replay()
    torch.cuda.synchronize()
    for k in STATE_KEYS:
        assert torch.equal(o3[k], ref_out[k]), k
    assert torch.equal(oo3["reward"], ref_o["reward"]) and torch.equal(oo3["traj"], ref_o["traj"])
    # aliasing: step in place
    st_alias = {k: v.clone() for k, v in st_in.items()}
    oo4 = eng.alloc_outputs(B)
    eng.step(st_alias, a, st_alias, oo4)
    torch.cuda.synchronize()
    for k in STATE_KEYS:
        assert torch.equal(st_alias[k], ref_out[k]), k
    assert torch.equal(oo4["obs"], ref_o["obs"])


def test_nan_guard_and_autoreset_quirk(gpu_env, rodent):
    import torch
    B = 8
    qpos, qvel, start = start_states(rodent, B, seed=13)
    s = gpu_env.reset_from(qpos, qvel, start)
    s.pipeline_state["qvel"][3, 20] = float("nan")
    a = torch.zeros(B, 30, device="cuda")
    s1 = gpu_env.step(s, a)
    assert float(s1.done[3]) == 1.0 and torch.isfinite(s1.obs[3]).all() and torch.isfinite(s1.reward[3])
    assert float(s1.done[0]) == 0.0
    # Q7: sub_clip_frame only increments -> done = 1 from step sub_clip_length on
    s = gpu_env.reset_from(qpos, qvel, start)
    dones = []
    for _ in range(12):
        s = gpu_env.step(s, a)
        dones.append(s.done.cpu().numpy().copy())
    assert (np.array(dones)[9:] == 1).all() and (s.info["sub_clip_frame"].cpu().numpy() == 12).all()


def eng_resident(env):
    return env.engine.resident_envs


def test_full_size_properties(gpu_env, rodent):
    """BASELINE config 2 size (4096 envs): size-independent properties + agreement of a slice with a small batch."""
    import torch
    B = 4096
    qpos, qvel, start = start_states(rodent, B, seed=14)
    acts = np.random.default_rng(15).uniform(-1, 1, size=(2, B, 30)).astype(np.float32)
    s = _run_steps(gpu_env, qpos, qvel, start, acts)
    q = s.pipeline_state["qpos"]
    assert torch.isfinite(q).all() and torch.isfinite(s.obs).all() and torch.isfinite(s.info["traj"]).all()
    assert (s.info["cur_frame"].cpu().numpy() == start + 2).all() and (s.info["sub_clip_frame"] == 2).all()
    assert (torch.linalg.norm(q[:, 3:7], dim=1) - 1).abs().max() < 1e-5
    assert torch.equal(s.obs[:, :74], torch.nan_to_num(q))  # obs = [qpos, qvel, qfrc_actuator, xpos[ee]] (rodent.py:337-344)
    assert torch.equal(s.obs[:, 74:147], s.pipeline_state["qvel"])
    assert torch.equal(s.obs[:, 147:220], s.pipeline_state["qfrc_actuator"])
    ee = rodent["idx"]["end_eff_idx"]
    assert torch.equal(s.obs[:, 220:], s.pipeline_state["xpos"][:, ee].reshape(B, 12))
    total = sum(s.metrics[k] for k in ("rcom", "rvel", "rtrunk", "rquat", "ract", "rapp"))
    assert (s.reward - total).abs().max() < 1e-7
    # a slice of the first wave of the persistent grid and one of the second (same CTA slots and workspace rows, reused)
    assert B > eng_resident(gpu_env)
    for lo, hi in ((2000, 2016), (B - 16, B)):
        sm = _run_steps(gpu_env, qpos[lo:hi], qvel[lo:hi], start[lo:hi], acts[:, lo:hi])
        assert torch.equal(sm.pipeline_state["qpos"], q[lo:hi]) and torch.equal(sm.reward, s.reward[lo:hi]), lo
        assert torch.equal(sm.info["traj"], s.info["traj"][lo:hi]) and torch.equal(sm.pipeline_state["xpos"], s.pipeline_state["xpos"][lo:hi])


def test_fused_autoreset_equals_wrapper_semantics(gpu_env, rodent):
    """vnl_step_autoreset == vnl_step followed by brax AutoResetWrapper.step's tree_map(where(done, first, cur)) on the
    pipeline state and obs; info (frames, traj), reward, done, metrics untouched (quirk Q7)."""
    import torch
    B = 96
    eng = gpu_env.engine
    qpos, qvel, start = start_states(rodent, B, seed=16)
    s0 = gpu_env.reset_from(qpos, qvel, start)
    first, first_obs = dict(s0.pipeline_state), s0.obs
    st = dict(first); st["cur_frame"] = s0.info["cur_frame"].clone(); st["sub_clip_frame"] = s0.info["sub_clip_frame"].clone()
    st["sub_clip_frame"][::3] = 9  # every third env terminates on this step (sub_clip_frame reaches sub_clip_length)
    st["qpos"] = st["qpos"].clone(); st["qpos"][1::3, 2] = 0.6  # and every third one is "unhealthy" (z above range)
    a = torch.tensor(np.random.default_rng(17).uniform(-1, 1, size=(B, 30)).astype(np.float32), device="cuda")
    o1, oo1 = eng.alloc_state(B), eng.alloc_outputs(B)
    eng.step(st, a, o1, oo1)
    o2, oo2 = eng.alloc_state(B), eng.alloc_outputs(B)
    eng.step_autoreset(st, a, o2, oo2, first, first_obs)
    torch.cuda.synchronize()
    done = oo1["done"] > 0
    assert done[::3].all() and done[1::3].all() and not done[2::3].any()
    for k in STATE_KEYS:
        want = torch.where(done.reshape((B,) + (1,) * (o1[k].dim() - 1)), first[k], o1[k])
        assert torch.equal(o2[k], want), k
    assert torch.equal(oo2["obs"], torch.where(done[:, None], first_obs, oo1["obs"]))
    for k in ("traj", "reward", "done", "metrics"):
        assert torch.equal(oo2[k], oo1[k]), k
    assert torch.equal(o2["cur_frame"], o1["cur_frame"]) and torch.equal(o2["sub_clip_frame"], o1["sub_clip_frame"])


def test_fused_training_wrappers_equal_brax_semantics(gpu_env, rodent):
    """vnl_step_training == AutoResetWrapper(EpisodeWrapper(env)).step of brax (envs/wrappers/training.py) as installed at
    ppo_imitation/train.py:204-214 with action_repeat 1: the wrappers restated in torch around vnl_step, step by step."""
    import torch
    B, K, EP = 64, 16, 7.0  # episode_length 7 < sub_clip_length 10: truncation fires before the env's own done
    eng = gpu_env.engine
    qpos, qvel, start = start_states(rodent, B, seed=41)
    s0 = gpu_env.reset_from(qpos, qvel, start)
    first, first_obs = dict(s0.pipeline_state), s0.obs
    mk = lambda: dict({k: v.clone() for k, v in first.items()}, cur_frame=s0.info["cur_frame"].clone(),
                      sub_clip_frame=s0.info["sub_clip_frame"].clone())
    ref, fus = mk(), mk()
    ref["qpos"][5::9, 2] = 0.6; fus["qpos"][5::9, 2] = 0.6  # a few envs start unhealthy: done on step 1, counter restarts
    z = lambda: torch.zeros(B, device="cuda")
    r_steps, r_done = z(), z()
    f_steps, f_done, f_trunc = z(), z(), z()
    rng = np.random.default_rng(42)
    seen_trunc = seen_restart = False
    for it in range(K):
        a = torch.tensor(rng.uniform(-1, 1, size=(B, 30)).astype(np.float32), device="cuda")
        # --- reference: AutoResetWrapper.step { steps = where(done, 0, steps); EpisodeWrapper.step { env.step } ; restore }
        r_steps = torch.where(r_done > 0, torch.zeros_like(r_steps), r_steps)
        nxt, out = eng.alloc_state(B), eng.alloc_outputs(B)
        eng.step(ref, a, nxt, out)
        r_steps = r_steps + 1
        over = r_steps >= EP
        trunc = torch.where(over, 1 - out["done"], torch.zeros_like(r_steps))
        done = torch.where(over, torch.ones_like(r_steps), out["done"])
        for k in STATE_KEYS:
            nxt[k] = torch.where(done.reshape((B,) + (1,) * (nxt[k].dim() - 1)) > 0, first[k], nxt[k])
        obs = torch.where(done[:, None] > 0, first_obs, out["obs"])
        ref, r_done = nxt, done
        # --- fused
        fn, fo = eng.alloc_state(B), eng.alloc_outputs(B)
        eng.step_training(fus, a, fn, fo, first, first_obs, f_steps, f_done, f_steps, f_trunc, EP)  # steps updated in place
        torch.cuda.synchronize()
        assert torch.equal(f_steps, r_steps) and torch.equal(f_trunc, trunc) and torch.equal(fo["done"], done), it
        assert torch.equal(fo["obs"], obs) and torch.equal(fo["reward"], out["reward"]) and torch.equal(fo["traj"], out["traj"])
        for k in STATE_KEYS + ("cur_frame", "sub_clip_frame"):
            assert torch.equal(fn[k], ref[k]), (it, k)
        fus, f_done = fn, fo["done"].clone()
        seen_trunc = seen_trunc or bool((trunc > 0).any())
        seen_restart = seen_restart or bool(((r_steps == 1) & (torch.tensor(it > 0, device="cuda"))).any())
    assert float(r_steps.max()) <= EP and seen_trunc and seen_restart  # both wrapper paths were exercised


def test_host_stepper_chunks_equal_one_launch(gpu_env, rodent):
    """hostio.HostStepper (host buffers, chunked launches, D2H overlapped on a second stream) returns bit-identical
    results to whole-batch vnl_step_autoreset launches; the chunking is invisible (envs are independent)."""
    import torch
    B, K = 300, 12
    eng = gpu_env.engine
    qpos, qvel, start = start_states(rodent, B, seed=21)
    s0 = gpu_env.reset_from(qpos, qvel, start)
    acts = torch.tensor(np.random.default_rng(22).uniform(-1, 1, size=(K, B, 30)).astype(np.float32)).pin_memory()
    stepper = pkg("hostio").HostStepper(gpu_env, s0, autoreset=True, chunk=128)  # 3 ragged chunks
    assert [b - a for a, b in stepper.chunks] == [128, 128, 44]
    first, first_obs = dict(s0.pipeline_state), s0.obs
    st = {k: v.clone() for k, v in first.items()}
    st["cur_frame"], st["sub_clip_frame"] = s0.info["cur_frame"].clone(), s0.info["sub_clip_frame"].clone()
    nxt, out = eng.alloc_state(B), eng.alloc_outputs(B)
    for i in range(K):  # crosses sub_clip_length = 10, so the AutoReset restore is exercised
        host = stepper.step(acts[i])
        eng.step_autoreset(st, acts[i].cuda(), nxt, out, first, first_obs)
        torch.cuda.synchronize()
        for k in ("obs", "traj", "reward", "done"):
            assert torch.equal(host[k], out[k].cpu()), (i, k)
        st, nxt = nxt, st
    for k in STATE_KEYS + ("cur_frame", "sub_clip_frame"):
        assert torch.equal(stepper.state[k], st[k]), k
    assert float(host["done"].min()) == 1.0  # Q7: every env is past its sub-clip by now
    # with the episode counter (vnl_step_training per chunk): same thing against whole-batch launches
    ep = pkg("hostio").HostStepper(gpu_env, s0, autoreset=True, chunk=128, episode_length=5)
    st = {k: v.clone() for k, v in first.items()}
    st["cur_frame"], st["sub_clip_frame"] = s0.info["cur_frame"].clone(), s0.info["sub_clip_frame"].clone()
    steps, done_prev, trunc = (torch.zeros(B, device="cuda") for _ in range(3))
    for i in range(K):
        host = ep.step(acts[i])
        eng.step_training(st, acts[i].cuda(), nxt, out, first, first_obs, steps, done_prev, steps, trunc, 5.0)
        torch.cuda.synchronize()
        done_prev = out["done"].clone()
        for k in ("obs", "traj", "reward", "done"):
            assert torch.equal(host[k], out[k].cpu()), (i, k)
        assert torch.equal(ep.steps, steps) and torch.equal(ep.truncation, trunc)
        st, nxt = nxt, st
    assert float(steps.max()) <= 5.0


def test_two_warps_per_env_build_matches_one_warp(gpu_env, rodent):
    """The -DVNL_EW=2 instantiation (two warps and a named barrier per env, 64-lane programs) computes the same env step
    as the default one-warp build: integer outputs identical, floats to rounding (partial sums are grouped differently)."""
    import torch
    rod, envs = pkg("envs.rodent"), pkg("envs")
    os_env = __import__("os").environ
    old = os_env.get("VNL_ENV_WARPS")
    os_env["VNL_ENV_WARPS"] = "2"
    try:
        env2 = envs.RodentTracking(reference_clip=rodent["clip"], model=rodent["model"], device="cuda:0", **rod.RODENT_ENV_ARGS)
    finally:
        if old is None:
            del os_env["VNL_ENV_WARPS"]
        else:
            os_env["VNL_ENV_WARPS"] = old
    assert env2.engine.dims["env_warps"] == 2 and gpu_env.engine.dims["env_warps"] == 1
    B = 40
    qpos, qvel, start = start_states(rodent, B, seed=31)
    a = torch.tensor(np.random.default_rng(32).uniform(-1, 1, size=(B, 30)).astype(np.float32), device="cuda")
    s1, s2 = gpu_env.reset_from(qpos, qvel, start), env2.reset_from(qpos, qvel, start)
    assert torch.equal(s1.info["cur_frame"], s2.info["cur_frame"]) and (s1.obs - s2.obs).abs().max() < 1e-5
    s1, s2 = gpu_env.step(s1, a), env2.step(s2, a)
    # one step from (nearly) identical states: same active sets, same flags, floats to rounding
    assert torch.equal(s1.info["cur_frame"], s2.info["cur_frame"]) and torch.equal(s1.done, s2.done)
    assert torch.equal(s1.info["solver_stats"][:, 2:], s2.info["solver_stats"][:, 2:])  # active contacts / limits
    for k in ("qpos", "qvel", "xpos"):  # the truncated CG solve (6 iterations) amplifies rounding in a few envs: bulk + tail bound
        x1, x2 = s1.pipeline_state[k].reshape(B, -1), s2.pipeline_state[k].reshape(B, -1)
        err = (x1 - x2).abs().max(1).values / x1.abs().max()
        assert err.median() < 2e-5 and err.max() < 2e-2, (k, float(err.median()), float(err.max()))
    assert (s1.reward - s2.reward).abs().median() < 1e-5
    for _ in range(3):  # further steps under a constant large action are chaotic (trajectories separate): only sanity here
        s2 = env2.step(s2, a)
    assert torch.isfinite(s2.pipeline_state["qpos"]).all() and torch.isfinite(s2.obs).all()
    assert (torch.linalg.norm(s2.pipeline_state["qpos"][:, 3:7], dim=1) - 1).abs().max() < 1e-5


def test_rodent_pair_physics_matches_oracle(oracle_mod):
    """BASELINE configs[4] model: rodent_pair.xml (<replicate count=2>: 131 bodies, nv 146, 114 contacts, nefc 590, two
    kinematic trees).  Physics only (the reference has no env for it): per-stage arrays and one pipeline step vs the oracle."""
    import os
    import torch
    from conftest import ROOT
    mj, mb, libm = pkg("mjcf"), pkg("model_blob"), pkg("_lib")
    model = mj.load_model(os.path.join(ROOT, "vnl-brax-imitation_b200", "data", "rodent_pair_model.npz"))
    blob = mb.build_model_blob(model)
    dims = mb.read_dims(blob)
    assert (dims["nbody"], dims["nv"], dims["nu"], dims["ncon"], dims["nefc"], dims["nM"], dims["nroot"]) == (131, 146, 60, 114, 590, 2238, 2)
    eng = libm.Engine(blob, None, device="cuda:0")
    B = 6
    rng = np.random.default_rng(20)
    qpos = np.tile(model.arrays["qpos0"], (B, 1)).astype(np.float32)
    qpos[:, 7:74] += (0.05 * rng.standard_normal((B, 67))).astype(np.float32)
    qpos[:, 81:] += (0.05 * rng.standard_normal((B, 67))).astype(np.float32)
    qpos[:, 2] -= 0.02; qpos[:, 76] -= 0.015  # press both animals into the floor: active contacts
    qvel = (0.1 * rng.standard_normal((B, 146))).astype(np.float32)
    ctrl = rng.uniform(-1, 1, size=(B, 60)).astype(np.float32)
    st = dict(qpos=torch.tensor(qpos, device="cuda"), qvel=torch.tensor(qvel, device="cuda"))
    g = oracle_mod.split_dump(dims, eng.forward_dump(st, torch.tensor(ctrl, device="cuda")).cpu().numpy().astype(np.float64))
    ost = dict(qpos=qpos.astype(np.float64), qvel=qvel.astype(np.float64))
    o32 = oracle_mod.forward_dump(blob, ost, ctrl.astype(np.float64), precision=32, dims=dims)
    assert (o32["counters"][:, 2] > 0).all()
    for name, tol in dict(xpos=3e-6, xquat=3e-6, cinert=3e-6, cdof=5e-6, crb=5e-6, qM=5e-6, cvel=5e-6, qfrc_bias=3e-5,
                          qfrc_smooth=3e-5, qacc_smooth=2e-4, con_dist=5e-6, qacc=1e-2, qfrc_constraint=1e-2).items():
        err = rel(g[name], o32[name])
        assert np.isfinite(g[name]).all() and err < tol, (name, err)
    assert np.array_equal(g["counters"][:, 2:4], o32["counters"][:, 2:4])
    for t in (1, 66):  # subtree_com of both tree roots
        assert rel(g["subtree_com"][:, t], o32["subtree_com"][:, t]) < 3e-6

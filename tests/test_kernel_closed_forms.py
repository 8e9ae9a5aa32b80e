"""The CUDA kernel itself (not only the oracle) against closed forms of the physics it implements: the small analytic models of
tests/test_oracle_pins.py run through `vnl_pipeline_step` (C ABI, one launch of n substeps) on the GPU.  These do not involve the
oracle at all, so they are parity evidence that is independent of the restatement: soft-contact equilibrium (impedance, `efc_D`,
`aref`, pyramid), the plane-capsule collider, implicit joint damping (second factorisation + solve), actuator force and ctrl clamp.
fp32 tolerances, written at each assertion."""
import math
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from conftest import pkg

pytestmark = pytest.mark.gpu


def _engine(xml, **kw):
    mjcf, mb, libm = pkg("mjcf"), pkg("model_blob"), pkg("_lib")
    model = mjcf.compile_model(ET.fromstring(xml), solver="cg", **kw)
    return model, libm.Engine(mb.build_model_blob(model), None, device="cuda:0")


def _run(eng, qpos, qvel, nsteps, ctrl=None):
    import torch
    st = dict(qpos=torch.tensor(np.asarray(qpos, dtype=np.float32)[None], device="cuda"),
              qvel=torch.tensor(np.asarray(qvel, dtype=np.float32)[None], device="cuda"))
    out = eng.alloc_state(1)
    c = None if ctrl is None else torch.tensor(np.asarray(ctrl, dtype=np.float32)[None], device="cuda")
    eng.pipeline_step(st, c, out, nsteps, None)
    torch.cuda.synchronize()
    return out["qpos"][0].cpu().numpy().astype(np.float64), out["qvel"][0].cpu().numpy().astype(np.float64)


def _pivot_inertia(model):
    mjcf = pkg("mjcf")
    A = model.arrays
    mass, l = A["body_mass"][1], abs(A["body_ipos"][1][2])
    R = mjcf.quat_to_mat(A["body_iquat"][1])
    return (R @ np.diag(A["body_inertia"][1]) @ R.T)[1, 1] + mass * l * l


def test_sphere_rests_at_the_closed_form_penetration():
    """mu = 1: |r| = g (1 - imp) / (k imp^2) (tests/test_oracle_pins.py derives it); 3.6718e-4 m for this model."""
    pins = __import__("test_oracle_pins")
    model, eng = _engine(pins.BALL.format(mu=1.0, rho=1000.0, tc=0.02, dr=1.0), iterations=100, ls_iterations=50)
    q, v = _run(eng, [0, 0, 0.0499, 1, 0, 0, 0], np.zeros(6), 600)
    r = q[2] - 0.05
    A = model.arrays
    k, _ = pins._kb(A["pair_solref"][0], A["pair_solimp"][0], model.timestep)
    imp = pins._impedance(A["pair_solimp"][0], r)
    assert r < 0 and abs(abs(r) - 9.81 * (1 - imp) / (k * imp * imp)) < 2e-6  # fp32: z = 0.05 resolves 4e-9; solver tolerance on top
    assert abs(r + 3.6718e-4) < 2e-6 and np.abs(v).max() < 1e-4 and np.abs(q[:2]).max() < 1e-6


def test_capsule_rests_on_two_contacts():
    pins = __import__("test_oracle_pins")
    geom = 'type="capsule" fromto="-0.08 0 0 0.08 0 0" size="0.03"'
    model, eng = _engine(pins.REST.format(z=0.0298, geom=geom, quat=""), iterations=100, ls_iterations=50)
    A = model.arrays
    q, v = _run(eng, A["qpos0"], np.zeros(6), 800)
    r = q[2] - 0.03
    mass, mu = A["body_mass"][1], 1.0
    k, _ = pins._kb(A["pair_solref"][0], A["pair_solimp"][0], model.timestep)
    w = A["body_invweight0"][1][0] * (1 + mu * mu) * 2 * mu * mu / model.impratio
    imp = pins._impedance(A["pair_solimp"][0], r)
    lhs = 8 * (imp / ((1 - imp) * w)) * k * imp * abs(r)  # two contacts x four pyramid rows
    assert r < 0 and abs(lhs - mass * 9.81) < 5e-3 * mass * 9.81, (r, lhs, mass * 9.81)  # |r| = 2e-4 m carries 1e-5 relative fp32 noise in z
    assert np.abs(q[3:7] - A["qpos0"][3:7]).max() < 1e-5 and np.abs(v).max() < 1e-3


@pytest.mark.parametrize("ed,damp,n", [("enable", 0.02, 200), ("disable", 0.02, 200), ("enable", 5.0, 3)])
def test_joint_damping_decay(ed, damp, n):
    """v' = v I / (I + dt d) per step with eulerdamp (also for dt d / I = 8), v' = v (1 - dt d / I) without."""
    pins = __import__("test_oracle_pins")
    model, eng = _engine(pins.PEND.format(ed=ed, damp=str(damp), limit=""), iterations=6, ls_iterations=6)
    # gravity is a model constant of the blob: rebuild with zero gravity
    mjcf, mb, libm = pkg("mjcf"), pkg("model_blob"), pkg("_lib")
    model.gravity = np.zeros(3)
    eng = libm.Engine(mb.build_model_blob(model), None, device="cuda:0")
    I, dt, v0 = _pivot_inertia(model), model.timestep, 1.5
    fac = I / (I + dt * damp) if ed == "enable" else 1.0 - dt * damp / I
    q, v = _run(eng, [0.0], [v0], n)
    assert abs(v[0] - v0 * fac ** n) < 3e-5 * v0 * fac ** n, (v[0], v0 * fac ** n)  # n fp32 multiplications
    theta = dt * v0 * fac * (1 - fac ** n) / (1 - fac)
    assert abs(q[0] - theta) < 3e-5 * abs(theta)


def test_motor_torque_and_ctrl_clamp():
    pins = __import__("test_oracle_pins")
    gear = 3.0
    mjcf, mb, libm = pkg("mjcf"), pkg("model_blob"), pkg("_lib")
    model = mjcf.compile_model(ET.fromstring(pins.MOTOR.format(gear=gear)), solver="cg", iterations=6, ls_iterations=6)
    model.gravity = np.zeros(3)
    eng = libm.Engine(mb.build_model_blob(model), None, device="cuda:0")
    I, dt, n = _pivot_inertia(model), model.timestep, 50
    for u, ueff in ((0.4, 0.4), (-0.25, -0.25), (2.5, 1.0)):
        q, v = _run(eng, [0.0], [0.0], n, ctrl=[u])
        a = gear * ueff / I
        assert abs(v[0] - n * dt * a) < 1e-5 * abs(n * dt * a), (u, v[0], n * dt * a)
        assert abs(q[0] - dt * dt * a * n * (n + 1) / 2) < 1e-5 * abs(dt * dt * a * n * n)


def test_pendulum_period():
    """T = 2 pi sqrt(I / (m g l)) from the zero crossings of 3000 single-substep launches."""
    pins = __import__("test_oracle_pins")
    import torch
    model, eng = _engine(pins.PEND.format(ed="disable", damp="0", limit=""), iterations=6, ls_iterations=6)
    A = model.arrays
    mass, l = A["body_mass"][1], abs(A["body_ipos"][1][2])
    T = 2 * math.pi * math.sqrt(_pivot_inertia(model) / (mass * 9.81 * l))
    a, b = eng.alloc_state(1), eng.alloc_state(1)
    a["qpos"].fill_(0.01); a["qvel"].zero_()
    th = []
    for _ in range(1500):
        eng.pipeline_step(a, None, b, 1, None)
        th.append(b["qpos"].clone())
        a, b = b, a
    th = torch.cat(th).reshape(-1).cpu().numpy().astype(np.float64)
    t = (np.arange(len(th)) + 1) * model.timestep
    zc = [i for i in range(1, len(th)) if th[i - 1] > 0 >= th[i]]
    tz = [t[i - 1] + (t[i] - t[i - 1]) * th[i - 1] / (th[i - 1] - th[i]) for i in zc]
    assert len(tz) >= 1 and abs(tz[0] - T / 4) / T < 1e-3, (tz, T)  # first downward crossing at a quarter period
    if len(tz) >= 2:
        assert abs((tz[1] - tz[0]) - T) / T < 5e-4


def test_free_body_tumbles_by_eulers_equations():
    """I w' = -(w x I w) in the body frame, q <- q * exp(dt w'), COM falling with g: 200 substeps of the kernel in ONE launch against
    200 steps of that recurrence in float64 (fp32 kernel: 2e-4 on the angular velocity after 0.4 s of tumbling)."""
    pins = __import__("test_oracle_pins")
    mjcf = pkg("mjcf")
    model, eng = _engine(pins.TUMBLE, iterations=6, ls_iterations=6)
    A = model.arrays
    Ri = mjcf.quat_to_mat(A["body_iquat"][1])
    Ib = Ri @ np.diag(A["body_inertia"][1]) @ Ri.T
    dt, g = model.timestep, np.array([0, 0, -9.81])
    q = np.array([0.1, -0.2, 1.0, 0.8, 0.2, -0.4, 0.4]); q[3:] /= np.linalg.norm(q[3:])
    v = np.array([0.3, 0.1, -0.2, 3.0, -2.0, 5.0])
    gq, gv = _run(eng, q, v, 200)
    for _ in range(200):
        w1 = v[3:] + dt * np.linalg.solve(Ib, -np.cross(v[3:], Ib @ v[3:]))
        vl1 = v[:3] + dt * g
        a = np.linalg.norm(w1) * dt
        dq = np.concatenate([[math.cos(a / 2)], math.sin(a / 2) * w1 / np.linalg.norm(w1)])
        qq = mjcf.quat_mul(q[3:], dq)
        q = np.concatenate([q[:3] + dt * vl1, qq / np.linalg.norm(qq)])
        v = np.concatenate([vl1, w1])
    assert np.abs(gv[3:] - v[3:]).max() < 2e-4 * np.abs(v[3:]).max(), (gv[3:], v[3:])
    assert np.abs(gv[:3] - v[:3]).max() < 1e-5 and np.abs(gq[:3] - q[:3]).max() < 1e-5
    assert min(np.abs(gq[3:] - q[3:]).max(), np.abs(gq[3:] + q[3:]).max()) < 2e-4


def test_spring_armature_and_filtered_actuator():
    """v += dt (gear clip(gain act) - k (q - ref)) / (I + armature); q += dt v; act += dt (ctrl - act) / tau -- 300 substeps in one launch."""
    pins = __import__("test_oracle_pins")
    mjcf, mb, libm = pkg("mjcf"), pkg("model_blob"), pkg("_lib")
    k, ref, arm, gear, gain, tau, fmax = 0.8, 15.0, 0.003, 2.0, 4.0, 0.05, 3.0
    model = mjcf.compile_model(ET.fromstring(pins.SPRING.format(k=k, ref=ref, arm=arm, gear=gear, gain=gain, tau=tau, fmax=fmax)), solver="cg",
                               iterations=6, ls_iterations=6)
    model.gravity = np.zeros(3)
    eng = libm.Engine(mb.build_model_blob(model), None, device="cuda:0")
    I, dt, qref = _pivot_inertia(model) + arm, model.timestep, math.radians(ref)
    for u in (0.0, 0.3, 1.0, -2.0):
        ue = min(max(u, -1.0), 1.0)
        q, v, act = 0.05, 0.0, 0.0
        gq, gv = _run(eng, [q], [v], 300, ctrl=[u])
        for _ in range(300):
            force = min(max(gain * act, -fmax), fmax)
            v += dt * (gear * force - k * (q - qref)) / I
            q += dt * v
            act += dt * (ue - act) / tau
        assert abs(gv[0] - v) < 5e-5 * max(1.0, abs(v)) and abs(gq[0] - q) < 5e-5 * max(1.0, abs(q)), (u, gv[0], v, gq[0], q)

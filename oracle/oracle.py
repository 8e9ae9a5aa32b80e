"""ctypes front end of the CPU oracle (oracle/vnl_oracle.cpp).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs, never
by the product package."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvnl_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vnl_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _State(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("qpos", "qvel", "act", "qacc_warmstart", "xpos", "xquat", "subtree_com",
                                               "qfrc_actuator", "cur_frame", "sub_clip_frame")]


class _Out(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("obs", "traj", "reward", "done", "metrics", "stats")]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.vnl_oracle_dump_size.restype = ctypes.c_size_t
    return _lib


def max_threads() -> int:
    return int(lib().vnl_oracle_max_threads())


STATE_KEYS = ("qpos", "qvel", "act", "qacc_warmstart", "xpos", "xquat", "subtree_com", "qfrc_actuator")


def alloc_state(dims: Dict[str, int], B: int) -> Dict[str, np.ndarray]:
    return dict(qpos=np.zeros((B, dims["nq"])), qvel=np.zeros((B, dims["nv"])), act=np.zeros((B, dims["na"])),
                qacc_warmstart=np.zeros((B, dims["nv"])), xpos=np.zeros((B, dims["nbody"], 3)),
                xquat=np.zeros((B, dims["nbody"], 4)), subtree_com=np.zeros((B, 3)), qfrc_actuator=np.zeros((B, dims["nv"])),
                cur_frame=np.zeros(B, dtype=np.int32), sub_clip_frame=np.zeros(B, dtype=np.int32))


def alloc_outputs(B: int, obs_size: int, traj_size: int) -> Dict[str, np.ndarray]:
    return dict(obs=np.zeros((B, obs_size)), traj=np.zeros((B, traj_size)), reward=np.zeros(B), done=np.zeros(B),
                metrics=np.zeros((B, 7)), stats=np.zeros((B, 4), dtype=np.int32))


def _as_state(s: Dict[str, np.ndarray]) -> _State:
    st = _State()
    for k in STATE_KEYS:
        a = s.get(k)
        if a is not None:
            assert a.dtype == np.float64 and a.flags.c_contiguous, k
        setattr(st, k, a.ctypes.data if a is not None else None)
    for k in ("cur_frame", "sub_clip_frame"):
        a = s.get(k)
        if a is not None:
            assert a.dtype == np.int32 and a.flags.c_contiguous, k
        setattr(st, k, a.ctypes.data if a is not None else None)
    return st


def _as_out(o: Dict[str, np.ndarray]) -> _Out:
    out = _Out()
    for k in ("obs", "traj", "reward", "done", "metrics"):
        assert o[k].dtype == np.float64 and o[k].flags.c_contiguous
        setattr(out, k, o[k].ctypes.data)
    out.stats = o["stats"].ctypes.data if o.get("stats") is not None else None
    return out


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def step(model_blob: np.ndarray, task_blob: np.ndarray, state: Dict[str, np.ndarray], action: np.ndarray, *,
         precision: int = 32, nthreads: int = 0, dims: Dict[str, int], obs_size: int, traj_size: int):
    """RodentTracking.step for a batch; returns (new_state, outputs).  All float arrays are float64
    containers (fp32 values round-trip exactly when precision == 32)."""
    B = state["qpos"].shape[0]
    s_in = {k: _f64(state[k]) for k in STATE_KEYS}
    s_in["cur_frame"] = np.ascontiguousarray(state["cur_frame"], dtype=np.int32)
    s_in["sub_clip_frame"] = np.ascontiguousarray(state["sub_clip_frame"], dtype=np.int32)
    s_out = alloc_state(dims, B)
    out = alloc_outputs(B, obs_size, traj_size)
    act = _f64(action)
    a, b, o = _as_state(s_in), _as_state(s_out), _as_out(out)
    rc = lib().vnl_oracle_step(model_blob.ctypes.data_as(ctypes.c_void_p), task_blob.ctypes.data_as(ctypes.c_void_p),
                               int(precision), int(B), ctypes.byref(a), act.ctypes.data_as(ctypes.c_void_p),
                               ctypes.byref(b), ctypes.byref(o), int(nthreads or max_threads()))
    assert rc == 0, rc
    return s_out, out


def reset(model_blob: np.ndarray, task_blob: np.ndarray, qpos: np.ndarray, qvel: np.ndarray, start_frame: np.ndarray, *,
          precision: int = 32, nthreads: int = 0, dims: Dict[str, int], obs_size: int, traj_size: int):
    B = qpos.shape[0]
    s_in = dict(qpos=_f64(qpos), qvel=_f64(qvel), cur_frame=np.ascontiguousarray(start_frame, dtype=np.int32))
    s_out = alloc_state(dims, B)
    out = alloc_outputs(B, obs_size, traj_size)
    a, b, o = _as_state(s_in), _as_state(s_out), _as_out(out)
    rc = lib().vnl_oracle_reset(model_blob.ctypes.data_as(ctypes.c_void_p), task_blob.ctypes.data_as(ctypes.c_void_p),
                                int(precision), int(B), ctypes.byref(a), ctypes.byref(b), ctypes.byref(o),
                                int(nthreads or max_threads()))
    assert rc == 0, rc
    return s_out, out


def pipeline_step(model_blob: np.ndarray, state: Dict[str, np.ndarray], ctrl: Optional[np.ndarray], nsteps: int, *,
                  precision: int = 32, nthreads: int = 0, dims: Dict[str, int]):
    B = state["qpos"].shape[0]
    s_in = {k: _f64(state[k]) for k in ("qpos", "qvel", "act", "qacc_warmstart") if state.get(k) is not None}
    s_out = alloc_state(dims, B)
    stats = np.zeros((B, 4), dtype=np.int32)
    c = _f64(ctrl) if ctrl is not None else None
    a, b = _as_state(s_in), _as_state(s_out)
    rc = lib().vnl_oracle_pipeline_step(model_blob.ctypes.data_as(ctypes.c_void_p), int(precision), int(B), int(nsteps),
                                        ctypes.byref(a), c.ctypes.data_as(ctypes.c_void_p) if c is not None else None,
                                        ctypes.byref(b), stats.ctypes.data_as(ctypes.c_void_p),
                                        int(nthreads or max_threads()))
    assert rc == 0, rc
    return s_out, stats


def dump_layout(dims: Dict[str, int]):
    """(name, shape) list of the stage dump, in order (mirrors forward_dump in vnl_oracle.cpp and
    the VNL_DUMP order of the CUDA test hook)."""
    nb, nv, nj, na, nc, ne = dims["nbody"], dims["nv"], dims["njnt"], dims["na"], dims["ncon"], dims["nefc"]
    return [("xpos", (nb, 3)), ("xquat", (nb, 4)), ("xmat", (nb, 9)), ("xipos", (nb, 3)), ("ximat", (nb, 9)),
            ("xanchor", (nj, 3)), ("xaxis", (nj, 3)), ("subtree_com", (nb, 3)), ("cinert", (nb, 10)), ("cdof", (nv, 6)),
            ("crb", (nb, 10)), ("qM", (nv, nv)), ("cvel", (nb, 6)), ("cdof_dot", (nv, 6)), ("qfrc_passive", (nv,)),
            ("qfrc_bias", (nv,)), ("qfrc_actuator", (nv,)), ("act_dot", (na,)), ("qfrc_smooth", (nv,)),
            ("qacc_smooth", (nv,)), ("con_dist", (nc,)), ("con_pos", (nc, 3)), ("con_frame", (nc, 9)),
            ("efc_pos", (ne,)), ("efc_D", (ne,)), ("efc_aref", (ne,)), ("efc_J", (ne, nv)), ("qacc", (nv,)),
            ("qfrc_constraint", (nv,)), ("efc_force", (ne,)), ("counters", (4,))]


def split_dump(dims: Dict[str, int], flat: np.ndarray) -> Dict[str, np.ndarray]:
    out, o = {}, 0
    for name, shape in dump_layout(dims):
        n = int(np.prod(shape))
        out[name] = flat[..., o:o + n].reshape(flat.shape[:-1] + shape)
        o += n
    assert o == flat.shape[-1], (o, flat.shape)
    return out


def forward_dump(model_blob: np.ndarray, state: Dict[str, np.ndarray], ctrl: Optional[np.ndarray], *, precision: int = 32,
                 dims: Dict[str, int]) -> Dict[str, np.ndarray]:
    B = state["qpos"].shape[0]
    n = int(lib().vnl_oracle_dump_size(model_blob.ctypes.data_as(ctypes.c_void_p)))
    s_in = {k: _f64(state[k]) for k in ("qpos", "qvel", "act", "qacc_warmstart") if state.get(k) is not None}
    c = _f64(ctrl) if ctrl is not None else None
    dump = np.zeros((B, n))
    a = _as_state(s_in)
    rc = lib().vnl_oracle_forward_dump(model_blob.ctypes.data_as(ctypes.c_void_p), int(precision), int(B), ctypes.byref(a),
                                       c.ctypes.data_as(ctypes.c_void_p) if c is not None else None,
                                       dump.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    return split_dump(dims, dump)

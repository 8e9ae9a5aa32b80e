"""ctypes binding of libvnl_b200.so (the C ABI of include/vnl_b200.h) over torch CUDA tensors.

torch is plumbing only: device memory, streams.  There is NO CPU fallback: if the CUDA library
is missing or no GPU is visible, construction raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VNL_B200_LIB") or os.path.join(_HERE, "libvnl_b200.so")  # override: developer A/B builds

STATE_F = ("qpos", "qvel", "act", "qacc_warmstart", "xpos", "xquat", "subtree_com", "qfrc_actuator")
STATE_I = ("cur_frame", "sub_clip_frame")
OUT_F = ("obs", "traj", "reward", "done", "metrics")


class VnlState(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in STATE_F + STATE_I]


class VnlEpisode(ctypes.Structure):
    _fields_ = [("steps_in", ctypes.c_void_p), ("done_in", ctypes.c_void_p), ("steps_out", ctypes.c_void_p),
                ("truncation_out", ctypes.c_void_p), ("episode_length", ctypes.c_float)]


class VnlOutputs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in OUT_F + ("stats",)]


EXPORTS = ("vnl_step", "vnl_reset", "vnl_pipeline_step", "vnl_forward_dump", "vnl_dump_size", "vnl_check_model",
           "vnl_check_task", "vnl_register_blob", "vnl_unregister_blob", "vnl_step_smem_bytes", "vnl_xla_step",
           "vnl_xla_reset", "vnl_version", "vnl_ffma_probe", "vnl_step_profiled", "vnl_step_autoreset", "vnl_envs_per_cta",
           "vnl_resident_envs", "vnl_workspace_bytes", "vnl_set_workspace", "vnl_step_training",
           "vnl_debug_layout")


def load_library() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() (nvcc, sm_100a). "
                           "There is no CPU fallback for the product path.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.vnl_version.restype = ctypes.c_char_p
    lib.vnl_dump_size.restype = ctypes.c_size_t
    lib.vnl_dump_size.argtypes = [ctypes.c_void_p]
    lib.vnl_step_smem_bytes.argtypes = [ctypes.c_void_p]
    lib.vnl_envs_per_cta.argtypes = [ctypes.c_void_p]
    lib.vnl_resident_envs.argtypes = [ctypes.c_void_p]
    lib.vnl_workspace_bytes.restype = ctypes.c_size_t
    lib.vnl_workspace_bytes.argtypes = [ctypes.c_void_p]
    lib.vnl_set_workspace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    lib.vnl_check_model.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    lib.vnl_check_task.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    lib.vnl_register_blob.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    lib.vnl_unregister_blob.argtypes = [ctypes.c_void_p]
    lib.vnl_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(VnlState), ctypes.c_void_p,
                             ctypes.POINTER(VnlState), ctypes.POINTER(VnlOutputs), ctypes.c_void_p]
    lib.vnl_step_autoreset.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(VnlState), ctypes.c_void_p,
                                       ctypes.POINTER(VnlState), ctypes.POINTER(VnlOutputs), ctypes.POINTER(VnlState),
                                       ctypes.c_void_p, ctypes.c_void_p]
    lib.vnl_step_training.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(VnlState), ctypes.c_void_p,
                                      ctypes.POINTER(VnlState), ctypes.POINTER(VnlOutputs), ctypes.POINTER(VnlState),
                                      ctypes.c_void_p, ctypes.POINTER(VnlEpisode), ctypes.c_void_p]
    lib.vnl_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(VnlState),
                              ctypes.POINTER(VnlState), ctypes.POINTER(VnlOutputs), ctypes.c_void_p]
    lib.vnl_pipeline_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(VnlState), ctypes.c_void_p,
                                      ctypes.POINTER(VnlState), ctypes.c_void_p, ctypes.c_void_p]
    lib.vnl_forward_dump.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(VnlState), ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p]
    lib.vnl_ffma_probe.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.vnl_step_profiled.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(VnlState), ctypes.c_void_p,
                                      ctypes.POINTER(VnlState), ctypes.POINTER(VnlOutputs), ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_int]
    return lib


def _ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    """Device-resident model / task blobs + typed wrappers of the C entry points.

    All tensors are torch CUDA tensors (fp32 / int32, contiguous, batch-major)."""

    def __init__(self, model_blob: np.ndarray, task_blob: Optional[np.ndarray] = None, device: str = "cuda:0"):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("vnl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.lib = load_library()
        self.device = torch.device(device)
        self.model_host = np.ascontiguousarray(model_blob, dtype=np.uint32)
        rc = self.lib.vnl_check_model(self.model_host.ctypes.data, self.model_host.nbytes)
        if rc:
            raise ValueError(f"bad model blob ({rc})")
        self.model_dev = torch.from_numpy(self.model_host.view(np.int32)).to(self.device)
        self._register(self.model_dev, self.model_host)
        self.task_host = self.task_dev = None
        if task_blob is not None:
            self.task_host = np.ascontiguousarray(task_blob, dtype=np.uint32)
            rc = self.lib.vnl_check_task(self.task_host.ctypes.data, self.task_host.nbytes)
            if rc:
                raise ValueError(f"bad task blob ({rc})")
            self.task_dev = torch.from_numpy(self.task_host.view(np.int32)).to(self.device)
            self._register(self.task_dev, self.task_host)
        from . import model_blob as mb
        self.dims = mb.read_dims(self.model_host)
        self.dump_size = int(self.lib.vnl_dump_size(self.model_host.ctypes.data))
        self.smem_bytes = int(self.lib.vnl_step_smem_bytes(self.model_host.ctypes.data))
        self.envs_per_cta = int(self.lib.vnl_envs_per_cta(self.model_host.ctypes.data))
        with torch.cuda.device(self.device):
            self.resident_envs = int(self.lib.vnl_resident_envs(self.model_host.ctypes.data))
            nbytes = int(self.lib.vnl_workspace_bytes(self.model_host.ctypes.data))
        # inertia workspace of the resident envs (L2-resident scratch the kernels address by CTA / env slot)
        self.workspace = torch.empty(max(nbytes, 4) // 4, dtype=torch.float32, device=self.device)
        rc = self.lib.vnl_set_workspace(self.model_dev.data_ptr(), self.workspace.data_ptr(), self.workspace.numel() * 4)
        if rc:
            raise RuntimeError(f"vnl_set_workspace failed ({rc})")
        if self.task_host is not None:
            self.obs_size = int(self.task_host[mb.C["VNL_TH_OBS_SIZE"]])
            self.traj_size = int(self.task_host[mb.C["VNL_TH_TRAJ_SIZE"]])
            self.n_frames = int(self.task_host[mb.C["VNL_TH_NFRAMES"]])
        self.launches = 0

    def _register(self, dev, host):
        rc = self.lib.vnl_register_blob(dev.data_ptr(), host.ctypes.data, host.nbytes)
        if rc:
            raise RuntimeError(f"vnl_register_blob failed ({rc})")

    def close(self):
        for t in (self.model_dev, self.task_dev):
            if t is not None:
                self.lib.vnl_unregister_blob(t.data_ptr())

    # ---- allocation helpers ---------------------------------------------------------------
    def alloc_state(self, B: int) -> Dict[str, "torch.Tensor"]:
        t, d, dev = self.torch, self.dims, self.device
        z = lambda *s: t.zeros(*s, dtype=t.float32, device=dev)
        return dict(qpos=z(B, d["nq"]), qvel=z(B, d["nv"]), act=z(B, d["na"]), qacc_warmstart=z(B, d["nv"]),
                    xpos=z(B, d["nbody"], 3), xquat=z(B, d["nbody"], 4), subtree_com=z(B, 3), qfrc_actuator=z(B, d["nv"]),
                    cur_frame=t.zeros(B, dtype=t.int32, device=dev), sub_clip_frame=t.zeros(B, dtype=t.int32, device=dev))

    def alloc_outputs(self, B: int) -> Dict[str, "torch.Tensor"]:
        t, dev = self.torch, self.device
        z = lambda *s: t.zeros(*s, dtype=t.float32, device=dev)
        return dict(obs=z(B, self.obs_size), traj=z(B, self.traj_size), reward=z(B), done=z(B), metrics=z(B, 7),
                    stats=t.zeros(B, 4, dtype=t.int32, device=dev))

    @staticmethod
    def _state(s: Dict) -> VnlState:
        st = VnlState()
        for k in STATE_F:
            v = s.get(k)
            if v is not None:
                assert v.is_contiguous() and v.dtype.is_floating_point and v.element_size() == 4, k
            setattr(st, k, _ptr(v))
        for k in STATE_I:
            v = s.get(k)
            if v is not None:
                assert v.is_contiguous() and v.element_size() == 4, k
            setattr(st, k, _ptr(v))
        return st

    @staticmethod
    def _outputs(o: Dict) -> VnlOutputs:
        out = VnlOutputs()
        for k in OUT_F:
            assert o[k].is_contiguous()
            setattr(out, k, _ptr(o[k]))
        out.stats = _ptr(o.get("stats"))
        return out

    def _stream(self) -> int:
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _check(self, rc: int, what: str):
        if rc:
            raise RuntimeError(f"{what} failed with code {rc}")
        self.launches += 1

    # ---- entry points -----------------------------------------------------------------------
    def step(self, state: Dict, action, out_state: Dict, outputs: Dict):
        B = state["qpos"].shape[0]
        a, b, o = self._state(state), self._state(out_state), self._outputs(outputs)
        assert action.is_contiguous() and action.shape == (B, self.dims["nu"])
        self._check(self.lib.vnl_step(self.model_dev.data_ptr(), self.task_dev.data_ptr(), B, ctypes.byref(a),
                                      action.data_ptr(), ctypes.byref(b), ctypes.byref(o), self._stream()), "vnl_step")

    def step_autoreset(self, state: Dict, action, out_state: Dict, outputs: Dict, first: Dict, first_obs):
        """`vnl_step` + brax AutoResetWrapper in the same launch (restore `first` / `first_obs` where done)."""
        B = state["qpos"].shape[0]
        a, b, o, f = self._state(state), self._state(out_state), self._outputs(outputs), self._state(first)
        assert action.is_contiguous() and action.shape == (B, self.dims["nu"]) and first_obs.is_contiguous()
        self._check(self.lib.vnl_step_autoreset(self.model_dev.data_ptr(), self.task_dev.data_ptr(), B, ctypes.byref(a),
                                                action.data_ptr(), ctypes.byref(b), ctypes.byref(o), ctypes.byref(f),
                                                first_obs.data_ptr(), self._stream()), "vnl_step_autoreset")

    def step_training(self, state: Dict, action, out_state: Dict, outputs: Dict, first: Dict, first_obs, steps, done_in,
                      steps_out, truncation, episode_length: float):
        """`vnl_step_training`: AutoResetWrapper(EpisodeWrapper(env)).step in one launch (action_repeat 1)."""
        B = state["qpos"].shape[0]
        a, b, o, f = self._state(state), self._state(out_state), self._outputs(outputs), self._state(first)
        assert action.is_contiguous() and action.shape == (B, self.dims["nu"]) and first_obs.is_contiguous()
        for t in (steps, done_in, steps_out, truncation):
            assert t.is_contiguous() and t.shape == (B,) and t.element_size() == 4 and t.dtype.is_floating_point
        ep = VnlEpisode(steps.data_ptr(), done_in.data_ptr(), steps_out.data_ptr(), truncation.data_ptr(), float(episode_length))
        self._check(self.lib.vnl_step_training(self.model_dev.data_ptr(), self.task_dev.data_ptr(), B, ctypes.byref(a),
                                               action.data_ptr(), ctypes.byref(b), ctypes.byref(o), ctypes.byref(f),
                                               first_obs.data_ptr(), ctypes.byref(ep), self._stream()), "vnl_step_training")

    def reset(self, state: Dict, out_state: Dict, outputs: Dict):
        B = state["qpos"].shape[0]
        a, b, o = self._state(state), self._state(out_state), self._outputs(outputs)
        self._check(self.lib.vnl_reset(self.model_dev.data_ptr(), self.task_dev.data_ptr(), B, ctypes.byref(a),
                                       ctypes.byref(b), ctypes.byref(o), self._stream()), "vnl_reset")

    def pipeline_step(self, state: Dict, ctrl, out_state: Dict, nsteps: int, stats=None):
        B = state["qpos"].shape[0]
        a, b = self._state(state), self._state(out_state)
        self._check(self.lib.vnl_pipeline_step(self.model_dev.data_ptr(), B, int(nsteps), ctypes.byref(a), _ptr(ctrl),
                                               ctypes.byref(b), _ptr(stats), self._stream()), "vnl_pipeline_step")

    def forward_dump(self, state: Dict, ctrl=None):
        t = self.torch
        B = state["qpos"].shape[0]
        dump = t.empty(B, self.dump_size, dtype=t.float32, device=self.device)
        a = self._state(state)
        self._check(self.lib.vnl_forward_dump(self.model_dev.data_ptr(), B, ctypes.byref(a), _ptr(ctrl), dump.data_ptr(),
                                              self._stream()), "vnl_forward_dump")
        return dump

"""Host side of the PPO data-side ops (C ABI in include/vnl_ppo.h): `compute_gae` of the reference
(ppo_imitation/intention_losses.py:26-89) as one launch.  Same argument names and order; time-major [T, B] tensors.
No CPU fallback."""
from __future__ import annotations

import ctypes

from . import _lib

PPO_EXPORTS = ("vnl_gae", "vnl_xla_gae")
_LIB = None


def _bind(lib):
    v = ctypes.c_void_p
    lib.vnl_gae.argtypes = [ctypes.c_int, ctypes.c_int, v, v, v, v, v, ctypes.c_float, ctypes.c_float, v, v, v]
    lib.vnl_xla_gae.argtypes = [v, ctypes.POINTER(v), ctypes.c_char_p, ctypes.c_size_t, v]
    lib.vnl_xla_gae.restype = None
    return lib


def compute_gae(truncation, termination, rewards, values, bootstrap_value, lambda_: float = 1.0, discount: float = 0.99):
    """(vs, advantages), both [T, B] — `compute_gae(truncation, termination, rewards, values, bootstrap_value, lambda_,
    discount)` of intention_losses.py:26-89 (defaults as there; the loss passes gae_lambda 0.95 and its discounting)."""
    import torch

    global _LIB
    if not torch.cuda.is_available():
        raise RuntimeError("vnl_b200 GAE needs a CUDA device (sm_100a); there is no CPU fallback")
    if _LIB is None:
        _LIB = _bind(_lib.load_library())
    T, B = rewards.shape
    for x in (truncation, termination, rewards, values):
        if x.shape != (T, B) or x.dtype != torch.float32 or not x.is_contiguous() or not x.is_cuda:
            raise ValueError("GAE operands must be contiguous fp32 [T, B] CUDA tensors")
    if bootstrap_value.shape != (B,) or bootstrap_value.dtype != torch.float32 or not bootstrap_value.is_contiguous():
        raise ValueError("bootstrap_value must be a contiguous fp32 [B] tensor")
    vs, adv = torch.empty_like(rewards), torch.empty_like(rewards)
    with torch.cuda.device(rewards.device):
        rc = _LIB.vnl_gae(T, B, truncation.data_ptr(), termination.data_ptr(), rewards.data_ptr(), values.data_ptr(),
                          bootstrap_value.data_ptr(), float(lambda_), float(discount), vs.data_ptr(), adv.data_ptr(),
                          torch.cuda.current_stream(rewards.device).cuda_stream)
    if rc:
        raise RuntimeError(f"vnl_gae failed ({rc})")
    return vs, adv

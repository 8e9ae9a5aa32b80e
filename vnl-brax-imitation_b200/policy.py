"""Host side of the intention-network policy forward (C ABI in include/vnl_policy.h).

Mirrors the reference's rollout policy (ppo_imitation/ppo_networks.py:45-83 `make_inference_fn(...).policy` over
ppo_imitation/intention_policy_network.py:20-105): same parameter tree (flax names), same call shape
`policy(traj, obs, key) -> (action, extras{log_prob, rand_log_prob, raw_action, logits})`, the random draws passed as
operands instead of a PRNG key.  The arithmetic runs in libvnl_b200.so (`vnl_policy_forward`, tcgen05 tensor cores);
there is no CPU or torch fallback on this path — `reference_forward` below is the fp32 torch restatement that the tests
compare against, never the product.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib

# flax parameter tree of IntentionNetwork, in the order vnl_policy_pack takes the arrays
PARAM_ORDER = (
    "encoder/hidden_0/kernel", "encoder/hidden_0/bias", "encoder/LayerNorm_0/scale", "encoder/LayerNorm_0/bias",
    "encoder/hidden_1/kernel", "encoder/hidden_1/bias", "encoder/LayerNorm_1/scale", "encoder/LayerNorm_1/bias",
    "encoder/fc2_mean/kernel", "encoder/fc2_mean/bias", "encoder/fc2_logvar/kernel", "encoder/fc2_logvar/bias",
    "decoder/hidden_0/kernel", "decoder/hidden_0/bias", "decoder/LayerNorm_0/scale", "decoder/LayerNorm_0/bias",
    "decoder/hidden_1/kernel", "decoder/hidden_1/bias", "decoder/LayerNorm_1/scale", "decoder/LayerNorm_1/bias",
    "decoder/hidden_2/kernel", "decoder/hidden_2/bias",
)


# the same entries as (module, layer, leaf) of the flax dict `params["params"]`
PARAM_ORDER_TUPLES = tuple(tuple(k.split("/")) for k in PARAM_ORDER)


class VnlPolicyDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("traj", "obs", "latent", "e1", "e2", "d1", "d2", "nu")]


POLICY_EXPORTS = ("vnl_policy_check", "vnl_policy_blob_bytes", "vnl_policy_pack", "vnl_policy_forward", "vnl_policy_debug",
                  "vnl_xla_policy_forward")


def _bind(lib):
    P = ctypes.POINTER(VnlPolicyDims)
    v = ctypes.c_void_p
    lib.vnl_policy_check.argtypes = [P]
    lib.vnl_policy_blob_bytes.argtypes = [P]
    lib.vnl_policy_blob_bytes.restype = ctypes.c_size_t
    lib.vnl_policy_pack.argtypes = [P, ctypes.POINTER(v), v, ctypes.c_size_t]
    lib.vnl_policy_forward.argtypes = [v, P, ctypes.c_int] + [v] * 15
    lib.vnl_policy_debug.argtypes = [v, P, ctypes.c_int, v, v, v, v, v, ctypes.c_int, v, v]
    lib.vnl_xla_policy_forward.argtypes = [v, ctypes.POINTER(v), ctypes.c_char_p, ctypes.c_size_t, v]
    lib.vnl_xla_policy_forward.restype = None
    return lib


def param_shapes(traj_size: int, obs_size: int, action_size: int, latent: int = 64, encoder_layer_sizes: Sequence[int] = (256, 128),
                 decoder_layer_sizes: Sequence[int] = (128, 256)) -> Dict[str, tuple]:
    """Shapes of the flax tree of `make_intention_policy` (intention_policy_network.py:108-139; sizes from
    configs/train_config.yaml:15-17; the decoder's last layer has NormalTanhDistribution.param_size = 2 * action_size)."""
    (e1, e2), (d1, d2) = encoder_layer_sizes, decoder_layer_sizes
    dense = {"encoder/hidden_0": (traj_size, e1), "encoder/hidden_1": (e1, e2), "encoder/fc2_mean": (e2, latent),
             "encoder/fc2_logvar": (e2, latent), "decoder/hidden_0": (latent + obs_size, d1), "decoder/hidden_1": (d1, d2),
             "decoder/hidden_2": (d2, 2 * action_size)}
    shapes = {}
    for k, (i, o) in dense.items():
        shapes[k + "/kernel"], shapes[k + "/bias"] = (i, o), (o,)
    for k, n in (("encoder/LayerNorm_0", e1), ("encoder/LayerNorm_1", e2), ("decoder/LayerNorm_0", d1), ("decoder/LayerNorm_1", d2)):
        shapes[k + "/scale"], shapes[k + "/bias"] = (n,), (n,)
    return shapes


def init_params(rng: np.random.Generator, shapes: Dict[str, tuple], perturb: float = 0.0) -> Dict[str, np.ndarray]:
    """Random init in the reference's scheme: lecun_uniform kernels (intention_policy_network.py:26,54), zero biases, unit
    LayerNorm scales.  `perturb` > 0 additionally jitters biases / scales so that tests exercise them."""
    out = {}
    for k, s in shapes.items():
        if k.endswith("/kernel"):
            lim = np.sqrt(3.0 / s[0])
            out[k] = rng.uniform(-lim, lim, size=s).astype(np.float32)
        elif k.endswith("/scale"):
            out[k] = (1.0 + perturb * rng.standard_normal(s)).astype(np.float32)
        else:
            out[k] = (perturb * rng.standard_normal(s)).astype(np.float32)
    return out


class IntentionPolicy:
    """Device-resident packed parameters + the `policy(traj, obs, draws)` call of the rollout."""

    def __init__(self, params: Dict[str, np.ndarray], device: str = "cuda:0", obs_mean=None, obs_std=None):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("vnl_b200 policy needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.lib = _bind(_lib.load_library())
        self.device = torch.device(device)
        sh = {k: params[k].shape for k in PARAM_ORDER}
        traj, e1 = sh["encoder/hidden_0/kernel"]
        e2, latent = sh["encoder/fc2_mean/kernel"]
        k4, d1 = sh["decoder/hidden_0/kernel"]
        d2, nlog = sh["decoder/hidden_2/kernel"]
        self.dims = VnlPolicyDims(traj, k4 - latent, latent, e1, e2, d1, d2, nlog // 2)
        rc = self.lib.vnl_policy_check(ctypes.byref(self.dims))
        if rc:
            raise ValueError(f"layer sizes not supported by the policy kernel ({rc})")
        self.traj_size, self.obs_size, self.latent, self.action_size = traj, k4 - latent, latent, nlog // 2
        self.load_params(params)
        self.set_normalizer(obs_mean, obs_std)
        self.launches = 0

    def load_params(self, params: Dict[str, np.ndarray]) -> None:
        """Pack the flax tree (host, C: vnl_policy_pack) and upload it; call again after every PPO update."""
        arrs = [np.ascontiguousarray(params[k], dtype=np.float32) for k in PARAM_ORDER]
        ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        n = int(self.lib.vnl_policy_blob_bytes(ctypes.byref(self.dims)))
        host = np.zeros((n + 3) // 4, dtype=np.int32)
        rc = self.lib.vnl_policy_pack(ctypes.byref(self.dims), ptrs, host.ctypes.data, host.nbytes)
        if rc:
            raise RuntimeError(f"vnl_policy_pack failed ({rc})")
        self.blob_host = host
        # The device blob keeps ONE address for the life of the policy: `rollout.Rollout` captures `blob_dev.data_ptr()` into
        # its CUDA graph, so a re-upload after a PPO update must land in the same buffer (stream-ordered copy).
        if getattr(self, "blob_dev", None) is None:
            self.blob_dev = self.torch.from_numpy(host).to(self.device)
        else:
            self.blob_dev.copy_(self.torch.from_numpy(host), non_blocking=False)

    def set_normalizer(self, mean, std) -> None:
        """brax running_statistics.normalize parameters for obs (ppo_imitation/train.py:220-229); None = identity."""
        t = self.torch
        # like the blob, the mean / std operands keep their device addresses once they exist (captured by CUDA graphs);
        # switching between identity (None) and a real normaliser after a capture is refused instead of silently ignored
        for name, val in (("obs_mean", mean), ("obs_std", std)):
            cur = getattr(self, name, None)
            if val is None:
                if cur is not None and getattr(self, "_norm_pinned", False):
                    raise ValueError("the normaliser operands were captured; pass tensors (e.g. zeros / ones), not None")
                setattr(self, name, None)
                continue
            src = t.as_tensor(val, dtype=t.float32)
            if cur is None:
                if getattr(self, "_norm_pinned", False):
                    raise ValueError("the policy was captured without a normaliser; build it with obs_mean / obs_std instead")
                # a contiguous fp32 tensor already on the device is ADOPTED (shared with e.g. normalizer.RunningStatistics,
                # whose in-place updates then reach the next launch); anything else is uploaded once
                setattr(self, name, src.to(self.device).contiguous())
            elif isinstance(val, t.Tensor) and val.data_ptr() == cur.data_ptr():
                pass  # the shared tensor itself
            else:
                if tuple(src.shape) != tuple(cur.shape):
                    raise ValueError("normaliser shape changed")
                cur.copy_(src)

    def pin_operands(self) -> None:
        """Called by a capturer (rollout.Rollout) once the operand addresses are baked into a CUDA graph."""
        self._norm_pinned = True

    def alloc_outputs(self, B: int, heads: bool = False):
        t, nu = self.torch, self.action_size
        f = lambda *s: t.empty(*s, dtype=t.float32, device=self.device)
        out = {"action": f(B, nu), "raw_action": f(B, nu), "logits": f(B, 2 * nu), "log_prob": f(B), "rand_log_prob": f(B)}
        if heads:
            out["z_mean"], out["z_logvar"] = f(B, self.latent), f(B, self.latent)
        return out

    def __call__(self, traj, obs, eps_z, eps_a, rand_action=None, out: Optional[dict] = None, heads: bool = False):
        """One launch: (action [B,nu], extras) as `policy(trajectories, observations, key)` of ppo_networks.py:55-83.
        eps_z [B,latent], eps_a [B,nu]: standard-normal draws; rand_action [B,nu]: the uniform(-1,1) draw (optional).
        eps_a=None gives the deterministic policy of evaluation (`mode(logits)`; the latent is still sampled, as there)."""
        t = self.torch
        B = traj.shape[0]
        for x, w in ((traj, self.traj_size), (obs, self.obs_size), (eps_z, self.latent), (eps_a, self.action_size)):
            if x is None and w == self.action_size:
                continue  # eps_a=None: `deterministic=True` of make_policy (ppo_networks.py:63-64), action = mode = tanh(loc)
            if x.dtype != t.float32 or not x.is_contiguous() or x.shape != (B, w) or x.device != self.device:
                raise ValueError("policy operands must be contiguous fp32 [B, width] tensors on the policy's device")
        if out is None:
            out = self.alloc_outputs(B, heads)
        p = lambda x: None if x is None else x.data_ptr()
        stream = t.cuda.current_stream(self.device).cuda_stream
        with t.cuda.device(self.device):
            rc = self.lib.vnl_policy_forward(self.blob_dev.data_ptr(), ctypes.byref(self.dims), B, p(traj), p(obs), p(self.obs_mean),
                                             p(self.obs_std), p(eps_z), p(eps_a), p(rand_action), p(out["action"]),
                                             p(out["raw_action"]), p(out["logits"]), p(out["log_prob"]),
                                             p(out["rand_log_prob"]) if rand_action is not None else None,
                                             p(out.get("z_mean")), p(out.get("z_logvar")), stream)
        if rc:
            raise RuntimeError(f"vnl_policy_forward failed ({rc})")
        self.launches += 1
        return out["action"], out

    def debug_layer(self, traj, obs, eps_z, layer: int):
        """Raw tensor-core accumulators of layer 0..5 for the first 128 envs (test hook)."""
        t = self.torch
        n = (self.dims.e1, self.dims.e2, 2 * self.dims.latent, self.dims.d1, self.dims.d2, (2 * self.dims.nu + 31) // 32 * 32)[layer]
        dump = t.zeros(128, n, dtype=t.float32, device=self.device)
        p = lambda x: None if x is None else x.data_ptr()
        with t.cuda.device(self.device):
            rc = self.lib.vnl_policy_debug(self.blob_dev.data_ptr(), ctypes.byref(self.dims), traj.shape[0], p(traj), p(obs),
                                           p(self.obs_mean), p(self.obs_std), p(eps_z), layer, dump.data_ptr(),
                                           t.cuda.current_stream(self.device).cuda_stream)
        if rc:
            raise RuntimeError(f"vnl_policy_debug failed ({rc})")
        return dump


class PrecisePolicy:
    """The same `policy(traj, obs, draws)` call as IntentionPolicy at the REFERENCE's precision (fp32 network, VERDICT r1 / ADVICE r1:
    the bf16 kernel's behaviour log-probs are 0.05-0.35 away from the fp32 network the learner re-evaluates, the size of the PPO
    clip range).  Runs the learner's forward kernels: `vnl_gemm_tf32` on the tensor cores -- `precision="3xtf32"` (default, fp32-class:
    log-probs within ~1e-4 of an fp32 / float64 evaluation) or `"tf32"` (one pass, what XLA runs for the reference on a GPU) -- plus
    relu + LayerNorm, reparameterisation and `vnl_policy_sample`.  ~25 launches instead of one (0.25 ms at 8192 envs beside a 7.7 ms
    env step), all CUDA-graph capturable: every buffer is allocated once per batch size, weights and their 3xTF32 splits are
    refreshed in place by `load_params`.  Drop-in for IntentionPolicy in rollout.Rollout (same attributes and call)."""

    def __init__(self, params: Dict[str, np.ndarray], device: str = "cuda:0", obs_mean=None, obs_std=None, precision: str = "3xtf32"):
        import torch

        from . import train_kernels as tk
        if not torch.cuda.is_available():
            raise RuntimeError("vnl_b200 policy needs a CUDA device (sm_100a); there is no CPU fallback")
        if precision not in ("3xtf32", "tf32"):
            raise ValueError("precision must be '3xtf32' or 'tf32'")
        self.torch, self.tk, self.x3 = torch, tk, precision == "3xtf32"
        self.device = torch.device(device)
        traj, e1 = params["encoder/hidden_0/kernel"].shape
        e2, latent = params["encoder/fc2_mean/kernel"].shape
        k4, d1 = params["decoder/hidden_0/kernel"].shape
        d2, nlog = params["decoder/hidden_2/kernel"].shape
        self.traj_size, self.obs_size, self.latent, self.action_size = traj, k4 - latent, latent, nlog // 2
        self.widths = (e1, e2, 2 * latent, d1, d2, nlog)
        if max(self.widths) > 1024 or self.action_size > 32 or any(w % 4 for w in self.widths) or self.obs_size % 4 or latent % 4:
            raise ValueError("layer sizes not supported by the fp32 policy path")
        f = lambda *sh: torch.zeros(*sh, dtype=torch.float32, device=self.device)
        self.W = {"e0": f(traj, e1), "e1": f(e1, e2), "eh": f(e2, 2 * latent), "d0": f(k4, d1), "d1": f(d1, d2), "d2": f(d2, nlog)}
        self.Wsplit = {k: (torch.empty_like(v), torch.empty_like(v)) for k, v in self.W.items()}
        self.b = {"e0": f(e1), "e1": f(e2), "eh": f(2 * latent), "d0": f(d1), "d1": f(d2), "d2": f(nlog)}
        self.ln = {k: (f(n), f(n)) for k, n in (("e0", e1), ("e1", e2), ("d0", d1), ("d1", d2))}
        self.blob_dev = self.W["e0"]  # an address that identifies "the parameters" for capturers (rollout.Rollout)
        self.obs_mean, self.obs_std = f(self.obs_size), torch.ones(self.obs_size, dtype=torch.float32, device=self.device)
        self._identity = True
        self.bufs = {}
        self.launches = 0
        self.load_params(params)
        self.set_normalizer(obs_mean, obs_std)

    def load_params(self, params) -> None:
        t = self.torch
        up = lambda a: t.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=self.device)
        self.W["e0"].copy_(up(params["encoder/hidden_0/kernel"])); self.b["e0"].copy_(up(params["encoder/hidden_0/bias"]))
        self.W["e1"].copy_(up(params["encoder/hidden_1/kernel"])); self.b["e1"].copy_(up(params["encoder/hidden_1/bias"]))
        self.W["eh"].copy_(t.cat([up(params["encoder/fc2_mean/kernel"]), up(params["encoder/fc2_logvar/kernel"])], 1))
        self.b["eh"].copy_(t.cat([up(params["encoder/fc2_mean/bias"]), up(params["encoder/fc2_logvar/bias"])]))
        for k, n in (("d0", "decoder/hidden_0"), ("d1", "decoder/hidden_1"), ("d2", "decoder/hidden_2")):
            self.W[k].copy_(up(params[n + "/kernel"])); self.b[k].copy_(up(params[n + "/bias"]))
        for k, n in (("e0", "encoder/LayerNorm_0"), ("e1", "encoder/LayerNorm_1"), ("d0", "decoder/LayerNorm_0"), ("d1", "decoder/LayerNorm_1")):
            self.ln[k][0].copy_(up(params[n + "/scale"])); self.ln[k][1].copy_(up(params[n + "/bias"]))
        if self.x3:
            for k, w in self.W.items():
                self.tk.split_into(w, *self.Wsplit[k])

    def set_normalizer(self, mean, std) -> None:
        """None = identity.  Tensors already on the device are ADOPTED when first given (shared with normalizer.RunningStatistics),
        later calls copy into the same storage."""
        t = self.torch
        for name, val, ident in (("obs_mean", mean, 0.0), ("obs_std", std, 1.0)):
            cur = getattr(self, name)
            if val is None:
                cur.fill_(ident)
            elif isinstance(val, t.Tensor) and val.is_cuda and val.dtype == t.float32 and val.is_contiguous() and self._identity and not getattr(self, "_norm_pinned", False):
                setattr(self, name, val)
            elif not (isinstance(val, t.Tensor) and val.data_ptr() == cur.data_ptr()):
                cur.copy_(t.as_tensor(val, dtype=t.float32))
        self._identity = mean is None and std is None and self._identity

    def pin_operands(self) -> None:
        self._norm_pinned = True

    def alloc_outputs(self, B: int, heads: bool = False):
        t, nu = self.torch, self.action_size
        f = lambda *s: t.empty(*s, dtype=t.float32, device=self.device)
        out = {"action": f(B, nu), "raw_action": f(B, nu), "logits": f(B, 2 * nu), "log_prob": f(B), "rand_log_prob": f(B)}
        if heads:
            out["z_mean"], out["z_logvar"] = f(B, self.latent), f(B, self.latent)
        return out

    def _buffers(self, B: int):
        if B not in self.bufs:
            t = self.torch
            f = lambda *s: t.zeros(*s, dtype=t.float32, device=self.device)
            e1, e2, h2, d1, d2, nlog = self.widths
            ld = (self.traj_size + 3) // 4 * 4
            b = dict(traj=f(B, ld), idx=t.arange(B, dtype=t.int32, device=self.device), pre0=f(B, e1), h0=f(B, e1), pre1=f(B, e2), h1=f(B, e2),
                     heads=f(B, h2), dec_in=f(B, self.latent + self.obs_size), pre2=f(B, d1), h2=f(B, d1), pre3=f(B, d2), h3=f(B, d2), st=f(B, 2))
            if self.x3:
                for k in ("traj", "h0", "h1", "dec_in", "h2", "h3"):
                    b[k + "_hi"], b[k + "_lo"] = t.empty_like(b[k]), t.empty_like(b[k])
            self.bufs[B] = b
        return self.bufs[B]

    def _dense(self, b, xname, rows, wname, out):
        tk = self.tk
        W = self.W[wname]
        K, N = W.shape
        parts, sk = None, 1
        if self.x3:
            tk.split_into(b[xname], b[xname + "_hi"], b[xname + "_lo"])
            parts = ((b[xname + "_hi"], b[xname + "_lo"]), self.Wsplit[wname])
            sk = max(1, ((K + 31) // 32 + 7) // 8)  # accumulation chains of <= 8 K blocks (the tensor core's accumulator truncates)
        tk.gemm(b[xname], 0, W, 1, out, rows, N, K, bias=self.b[wname], x3=self.x3, splitk=sk, parts=parts)

    def __call__(self, traj, obs, eps_z, eps_a, rand_action=None, out: Optional[dict] = None, heads: bool = False):
        """(action [B, nu], extras) as IntentionPolicy.__call__; rand_action: the reference's ONE uniform draw of shape (nu,)
        (a [B, nu] tensor is accepted for compatibility: its first row is used)."""
        t, tk = self.torch, self.tk
        L_ = tk.lib()
        B = traj.shape[0]
        for x, w in ((traj, self.traj_size), (obs, self.obs_size), (eps_z, self.latent)):
            if x.dtype != t.float32 or not x.is_contiguous() or x.shape != (B, w) or x.device != self.device:
                raise ValueError("policy operands must be contiguous fp32 [B, width] tensors on the policy's device")
        if out is None:
            out = self.alloc_outputs(B, heads)
        if B == 0:
            return out["action"], out
        b, st, ptr = self._buffers(B), tk.stream(traj), (lambda x: None if x is None else x.data_ptr())
        chk = tk.check
        with t.cuda.device(self.device):
            chk(L_.vnl_gather_rows(ptr(traj), 1, B, self.traj_size, ptr(b["idx"]), B, ptr(b["traj"]), b["traj"].shape[1], st), "pad traj")
            Ld = self.latent + self.obs_size
            chk(L_.vnl_obs_normalize(ptr(obs), self.obs_size, B, self.obs_size, ptr(self.obs_mean), ptr(self.obs_std), ptr(b["dec_in"]) + 4 * self.latent, Ld, st), "normalize")
            relu_ln = lambda pre, name, dst: chk(L_.vnl_relu_ln_fwd(ptr(pre), pre.shape[1], B, pre.shape[1], ptr(self.ln[name][0]), ptr(self.ln[name][1]), ptr(dst), dst.shape[1], ptr(b["st"]), st), "relu_ln")
            self._dense(b, "traj", B, "e0", b["pre0"]); relu_ln(b["pre0"], "e0", b["h0"])
            self._dense(b, "h0", B, "e1", b["pre1"]); relu_ln(b["pre1"], "e1", b["h1"])
            self._dense(b, "h1", B, "eh", b["heads"])
            chk(L_.vnl_reparam_fwd(ptr(b["heads"]), ptr(eps_z), B, self.latent, ptr(b["dec_in"]), Ld, st), "reparam")
            self._dense(b, "dec_in", B, "d0", b["pre2"]); relu_ln(b["pre2"], "d0", b["h2"])
            self._dense(b, "h2", B, "d1", b["pre3"]); relu_ln(b["pre3"], "d1", b["h3"])
            self._dense(b, "h3", B, "d2", out["logits"])
            if rand_action is not None and rand_action.dim() == 2:
                rand_action = rand_action[0].contiguous()
            chk(L_.vnl_policy_sample(ptr(out["logits"]), 2 * self.action_size, ptr(eps_a), ptr(rand_action), B, self.action_size, ptr(out["action"]),
                                     ptr(out["raw_action"]), ptr(out["log_prob"]), ptr(out["rand_log_prob"]) if rand_action is not None else None, st),
                "policy_sample")
            if heads:
                out["z_mean"].copy_(b["heads"][:, :self.latent]); out["z_logvar"].copy_(b["heads"][:, self.latent:])
        self.launches += 1
        return out["action"], out


def reference_forward(params, traj, obs, eps_z, eps_a, rand_action=None, obs_mean=None, obs_std=None, operand_dtype=None):
    """fp32 torch restatement of the reference policy (checker for tests / tools; works on CPU tensors too).
    `operand_dtype=torch.bfloat16` rounds the operands of every dense layer the way the kernel does (fp32 accumulation)."""
    import torch

    P = {k: torch.as_tensor(v, dtype=torch.float32, device=traj.device) for k, v in params.items()}
    rnd = (lambda x: x) if operand_dtype is None else (lambda x: x.to(operand_dtype).to(torch.float32))

    def dense(x, name):
        return torch.matmul(rnd(x).double(), rnd(P[name + "/kernel"]).double()).float() + P[name + "/bias"]

    def ln(x, name):  # flax LayerNorm: fast variance, eps 1e-6
        m = x.mean(-1, keepdim=True)
        var = torch.clamp((x * x).mean(-1, keepdim=True) - m * m, min=0.0)
        return (x - m) * torch.rsqrt(var + 1e-6) * P[name + "/scale"] + P[name + "/bias"]

    pre = {}
    h = traj
    for i in range(2):  # Encoder, intention_policy_network.py:31-41
        pre[i] = dense(h, f"encoder/hidden_{i}")
        h = ln(torch.relu(pre[i]), f"encoder/LayerNorm_{i}")
    mean, logvar = dense(h, "encoder/fc2_mean"), dense(h, "encoder/fc2_logvar")
    pre[2] = torch.cat([mean, logvar], -1)
    z = mean + eps_z * torch.exp(0.5 * logvar)  # reparameterize, :76-79
    o = obs if obs_mean is None else (obs - obs_mean) / obs_std  # the `apply` closure normalises obs only, :124-126
    h = torch.cat([z, o], -1)
    for i in range(2):  # Decoder, :58-73
        pre[3 + i] = dense(h, f"decoder/hidden_{i}")
        h = ln(torch.relu(pre[3 + i]), f"decoder/LayerNorm_{i}")
    logits = dense(h, "decoder/hidden_2")
    pre[5] = logits
    nu = logits.shape[-1] // 2
    loc, scale = logits[..., :nu], torch.nn.functional.softplus(logits[..., nu:]) + 1e-3  # brax NormalTanhDistribution
    raw = loc + scale * eps_a
    log_det = lambda x: 2.0 * (np.log(2.0) - x - torch.nn.functional.softplus(-2.0 * x))
    normal_lp = lambda x: -0.5 * ((x - loc) / scale) ** 2 - torch.log(scale) - 0.5 * np.log(2 * np.pi)
    out = {"action": torch.tanh(raw), "raw_action": raw, "logits": logits, "log_prob": (normal_lp(raw) - log_det(raw)).sum(-1),
           "z_mean": mean, "z_logvar": logvar, "pre": pre}
    if rand_action is not None:
        out["rand_log_prob"] = (normal_lp(rand_action) - log_det(rand_action)).sum(-1)
    return out

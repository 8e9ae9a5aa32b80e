"""The PPO update on the GPU: one `minibatch_step` of the reference (ppo_imitation/train.py:251-268) =
`jax.value_and_grad(compute_ppo_intention_loss)` (ppo_imitation/intention_losses.py:91-202) over the intention policy network
(intention_policy_network.py:20-105) and the brax value MLP (ppo_networks.py:114-118: hidden (1024, 1024), swish), the gradient
`pmean` over devices (one NCCL all-reduce of the flat gradient buffer, overlapped: the policy bucket is reduced while the value
network is still in its backward pass) and `optax.adam` (train.py:231-232).

Everything numeric runs in libvnl_b200.so (include/vnl_train.h): the dense contractions on the tensor cores
(`vnl_gemm_tf32`: tcgen05 kind::tf32, TMA tensor maps, K-major / MN-major operands so that forward, dgrad and wgrad need no
transposed copies), the rest as fp32 row kernels.  torch is device memory, streams and `torch.distributed`; there is no torch
or CPU fallback on this path.  `reference_loss` below is the torch-autograd restatement of the reference loss the tests compare
against (never the product).

Precision: `x3=False` runs one TF32 pass per product -- what XLA executes for the reference's f32 dots on an NVIDIA GPU
(jax default matmul precision); `x3=True` runs the 3xTF32 split (hi.hi + hi.lo + lo.hi, split-K chains of <= 8 K blocks because
the tensor core's accumulator truncates) and reproduces fp32 autograd to ~1e-5: the parity mode.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np

from . import policy as pol
from . import ppo as gae_mod
from . import train_kernels as tk

VALUE_ORDER = ("hidden_0/kernel", "hidden_0/bias", "hidden_1/kernel", "hidden_1/bias", "hidden_2/kernel", "hidden_2/bias")
METRIC_NAMES = ("total_loss", "policy_loss", "v_loss", "entropy_loss", "kl_loss_intention", "mean_rho", "clip_fraction")


def value_param_shapes(obs_size: int, hidden=(1024, 1024)) -> Dict[str, tuple]:
    """flax tree of brax `networks.make_value_network` (MLP layer_sizes = hidden + [1]; ppo_networks.py:114-118)."""
    sizes = [obs_size] + list(hidden) + [1]
    out = {}
    for i in range(len(sizes) - 1):
        out[f"hidden_{i}/kernel"], out[f"hidden_{i}/bias"] = (sizes[i], sizes[i + 1]), (sizes[i + 1],)
    return out


def init_value_params(rng: np.random.Generator, shapes, perturb: float = 0.0):
    """brax MLP init: lecun_uniform kernels, zero biases (`perturb` jitters the biases for tests)."""
    out = {}
    for k, s in shapes.items():
        if k.endswith("/kernel"):
            lim = math.sqrt(3.0 / s[0])
            out[k] = rng.uniform(-lim, lim, size=s).astype(np.float32)
        else:
            out[k] = (perturb * rng.standard_normal(s)).astype(np.float32)
    return out


# the two encoder heads are ONE tensor in the learner ([e2, 2 L]: mean | logvar columns), as in the rollout kernel
_POLICY_TENSORS = ("encoder/hidden_0/kernel", "encoder/hidden_0/bias", "encoder/LayerNorm_0/scale", "encoder/LayerNorm_0/bias",
                   "encoder/hidden_1/kernel", "encoder/hidden_1/bias", "encoder/LayerNorm_1/scale", "encoder/LayerNorm_1/bias",
                   "encoder/heads/kernel", "encoder/heads/bias",
                   "decoder/hidden_0/kernel", "decoder/hidden_0/bias", "decoder/LayerNorm_0/scale", "decoder/LayerNorm_0/bias",
                   "decoder/hidden_1/kernel", "decoder/hidden_1/bias", "decoder/LayerNorm_1/scale", "decoder/LayerNorm_1/bias",
                   "decoder/hidden_2/kernel", "decoder/hidden_2/bias")


class PPOLearner:
    """Flat parameter / gradient / Adam buffers + the workspaces of one minibatch shape (T x Bm rows).

    `loss_and_grads(batch)` -> metrics; `apply_gradients()` = all-reduce (if a process group exists) + Adam; `update(batch)` =
    both.  `policy_params()` / `value_params()` export the flax trees (the rollout policy re-packs from them)."""

    def __init__(self, policy_params: Dict[str, np.ndarray], value_params: Dict[str, np.ndarray], T: int, Bm: int, device: str = "cuda:0",
                 learning_rate: float = 6e-4, entropy_cost: float = 1e-4, discounting: float = 0.9, reward_scaling: float = 1.0,
                 gae_lambda: float = 0.95, clipping_epsilon: float = 0.3, normalize_advantage: bool = True, kl_weight: float = 1e-4,
                 x3: bool = False, adam_b1: float = 0.9, adam_b2: float = 0.999, adam_eps: float = 1e-8):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("vnl_b200 learner needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = t = torch
        self.device = dev = t.device(device)
        self.T, self.Bm, self.R = int(T), int(Bm), int(T) * int(Bm)
        self.hp = dict(learning_rate=learning_rate, entropy_cost=entropy_cost, discounting=discounting, reward_scaling=reward_scaling,
                       gae_lambda=gae_lambda, clipping_epsilon=clipping_epsilon, normalize_advantage=normalize_advantage, kl_weight=kl_weight,
                       b1=adam_b1, b2=adam_b2, eps=adam_eps)
        self.x3 = bool(x3)
        P = policy_params
        self.traj, e1 = P["encoder/hidden_0/kernel"].shape
        e2, self.L = P["encoder/fc2_mean/kernel"].shape
        k4, d1 = P["decoder/hidden_0/kernel"].shape
        d2, nlog = P["decoder/hidden_2/kernel"].shape
        self.obs, self.nu = k4 - self.L, nlog // 2
        self.widths = dict(e1=e1, e2=e2, d1=d1, d2=d2)
        self.vh = [value_params["hidden_0/kernel"].shape[1], value_params["hidden_1/kernel"].shape[1]]
        if max(e1, e2, d1, d2) > 256 or self.nu > 32 or any(w % 4 for w in (e1, e2, d1, d2, self.obs, 2 * self.nu, self.L)):
            raise ValueError("layer sizes not supported by the update kernels")
        # ---- flat buffers -------------------------------------------------------------------------------------------------
        shapes = {}
        for k in _POLICY_TENSORS:
            if k == "encoder/heads/kernel":
                shapes["policy/" + k] = (e2, 2 * self.L)
            elif k == "encoder/heads/bias":
                shapes["policy/" + k] = (2 * self.L,)
            else:
                shapes["policy/" + k] = tuple(P[k].shape)
        for k in VALUE_ORDER:
            shapes["value/" + k] = tuple(value_params[k].shape)
        self.shapes, self.offsets, off = shapes, {}, 0
        for k, s in shapes.items():
            self.offsets[k] = off
            off += (int(np.prod(s)) + 3) // 4 * 4  # 16-byte aligned tensors (TMA operands)
        self.nparams = off
        self.n_policy = self.offsets["value/hidden_0/kernel"]  # the policy bucket = [0, n_policy)
        z = lambda n, dt=t.float32: t.zeros(n, dtype=dt, device=dev)
        self.params, self.grads, self.m, self.v = z(off), z(off), z(off), z(off)
        self.step_dev, self.bc_dev = z(1, t.int32), z(2)
        self.p = {k: self.params[self.offsets[k]:self.offsets[k] + int(np.prod(s))].view(*s) for k, s in shapes.items()}
        self.g = {k: self.grads[self.offsets[k]:self.offsets[k] + int(np.prod(s))].view(*s) for k, s in shapes.items()}
        self.load_params(policy_params, value_params)
        # ---- workspaces of one minibatch ---------------------------------------------------------------------------------
        R, Bm, L, nu = self.R, self.Bm, self.L, self.nu
        f = lambda *s: t.zeros(*s, dtype=t.float32, device=dev)
        self.ld_traj = (self.traj + 3) // 4 * 4
        W = self.widths
        self.ws = dict(
            h0pre=f(R, e1), h0=f(R, e1), s0=f(R, 2), h1pre=f(R, e2), h1=f(R, e2), s1=f(R, 2), heads=f(R, 2 * L),
            dec_in=f(R, L + self.obs), d0pre=f(R, d1), d0=f(R, d1), s2=f(R, 2), d1pre=f(R, d2), d1=f(R, d2), s3=f(R, 2), logits=f(R, 2 * nu),
            vin=f(R + Bm, self.obs), v0pre=f(R + Bm, self.vh[0]), v0=f(R + Bm, self.vh[0]), v1pre=f(R + Bm, self.vh[1]), v1=f(R + Bm, self.vh[1]),
            val=f(R + Bm), target_lp=f(R), ent=f(R), termination=f(R), rewards_s=f(R), vs=f(R), adv=f(R),
            dlogits=f(R, 2 * nu), dval=f(R), dd1=f(R, d2), dd1pre=f(R, d2), dd0=f(R, d1), dd0pre=f(R, d1), ddec_in=f(R, L + self.obs),
            dheads=f(R, 2 * L), dh1=f(R, e2), dh1pre=f(R, e2), dh0=f(R, e1), dh0pre=f(R, e1),
            dv1pre=f(R, self.vh[1]), dv0=f(R, self.vh[0]), dv0pre=f(R, self.vh[0]),
            metrics=f(8), scratch2=f(2))
        self.obs_mean, self.obs_std = f(self.obs), t.ones(self.obs, dtype=t.float32, device=dev)
        self.updates = 0
        self.launches = 0
        self.side = t.cuda.Stream(device=dev)    # gradient exchange of the policy bucket
        self.branch = t.cuda.Stream(device=dev)  # the value network's forward / backward, beside the policy's
        self.leaf_p, self.leaf_v = t.cuda.Stream(device=dev), t.cuda.Stream(device=dev)  # weight / bias gradients: leaves of the backward chains
        self._gae = gae_mod._bind(tk.lib())
        self._pending = None
        del W

    # ---- parameters -----------------------------------------------------------------------------------------------------
    def load_params(self, policy_params, value_params) -> None:
        t = self.torch
        up = lambda a: t.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=self.device)
        for k in _POLICY_TENSORS:
            if k == "encoder/heads/kernel":
                self.p["policy/" + k].copy_(t.cat([up(policy_params["encoder/fc2_mean/kernel"]), up(policy_params["encoder/fc2_logvar/kernel"])], 1))
            elif k == "encoder/heads/bias":
                self.p["policy/" + k].copy_(t.cat([up(policy_params["encoder/fc2_mean/bias"]), up(policy_params["encoder/fc2_logvar/bias"])]))
            else:
                self.p["policy/" + k].copy_(up(policy_params[k]))
        for k in VALUE_ORDER:
            self.p["value/" + k].copy_(up(value_params[k]))

    def set_normalizer(self, mean, std) -> None:
        self.obs_mean.copy_(self.torch.as_tensor(mean, dtype=self.torch.float32))
        self.obs_std.copy_(self.torch.as_tensor(std, dtype=self.torch.float32))

    def _export(self, src, prefix) -> Dict[str, np.ndarray]:
        out = {}
        L = self.L
        for k in (_POLICY_TENSORS if prefix == "policy/" else VALUE_ORDER):
            a = src[prefix + k].detach().cpu().numpy()
            if k == "encoder/heads/kernel":
                out["encoder/fc2_mean/kernel"], out["encoder/fc2_logvar/kernel"] = a[:, :L].copy(), a[:, L:].copy()
            elif k == "encoder/heads/bias":
                out["encoder/fc2_mean/bias"], out["encoder/fc2_logvar/bias"] = a[:L].copy(), a[L:].copy()
            else:
                out[k] = a.copy()
        return out

    def policy_params(self):
        return self._export(self.p, "policy/")

    def value_params(self):
        return self._export(self.p, "value/")

    def policy_grads(self):
        return self._export(self.g, "policy/")

    def value_grads(self):
        return self._export(self.g, "value/")

    # ---- the three products of a dense layer -------------------------------------------------------------------------------
    def _splitk(self, K: int) -> int:
        kb = (K + 31) // 32
        return max(1, (kb + 7) // 8) if self.x3 else 1  # 3xTF32: accumulation chains of <= 8 K blocks (the accumulator truncates)

    def _fwd(self, x, rows, W, b, out, act=None):  # out[rows, out] = x[rows, in] . W[in, out] + b;  act = swish(out) from the same epilogue
        K, N = W.shape
        sk = self._splitk(K)
        fused = act is not None and sk <= 1 and N % 32 == 0
        tk.gemm(x, 0, W, 1, out, rows, N, K, bias=b, x3=self.x3, splitk=sk, epilogue=1 if fused else 0, aux=act if fused else None)
        self.launches += 1
        if act is not None and not fused:
            tk.check(tk.lib().vnl_swish_fwd(out.data_ptr(), rows * N, act.data_ptr(), tk.stream(out)), "swish")
            self.launches += 1

    def _dgrad(self, dy, rows, W, out, pre=None, tmp=None):  # out[rows, in] = dy[rows, out] . W[in, out]^T  (* swish'(pre) in the epilogue)
        N, K = W.shape
        sk = self._splitk(K)
        if pre is None:
            tk.gemm(dy, 0, W, 0, out, rows, N, K, x3=self.x3, splitk=sk)
        elif sk <= 1 and N % 32 == 0:
            tk.gemm(dy, 0, W, 0, out, rows, N, K, x3=self.x3, splitk=sk, epilogue=2, aux=pre)
        else:
            tk.gemm(dy, 0, W, 0, tmp, rows, N, K, x3=self.x3, splitk=sk)
            tk.check(tk.lib().vnl_swish_bwd(tmp.data_ptr(), pre.data_ptr(), rows * N, out.data_ptr(), tk.stream(out)), "swish_bwd")
            self.launches += 1
        self.launches += 1

    def _wgrad(self, x, dy, rows, gW):  # gW[in, out] = x[rows, in]^T . dy[rows, out]
        M, N = gW.shape
        tiles = ((M + 127) // 128) * ((N + (127 if N > 64 else 63)) // (128 if N > 64 else 64))
        kb = (rows + 31) // 32
        sk = max(self._splitk(rows), min(max(1, 148 // tiles), max(1, kb // 4)))
        if N % 256 == 0 and M >= 256:  # 256 x 256 tiles (vnl_gemm.cu picks them when tiles x splits still fill the machine)
            tb = ((M + 255) // 256) * (N // 256)
            skb = min(max(1, 148 // tb), max(1, kb // 4))
            if tb * skb >= 96:
                sk = max(self._splitk(rows), skb)
        tk.gemm(x, 1, dy, 1, gW, M, N, rows, x3=self.x3, splitk=sk, zero=False)  # self.grads was zeroed at the top of the evaluation
        self.launches += 1

    # ---- one loss + gradient evaluation ------------------------------------------------------------------------------------
    def loss_and_grads(self, batch: Dict[str, "torch.Tensor"], exchange: bool = True):
        """batch (time-major, contiguous fp32 CUDA tensors, R = T * Bm rows):
        traj [R, ld_traj] (row stride padded to a multiple of 4: vnl_gather_rows does it), observation [R, obs],
        next_observation_last [Bm, obs], reward / discount / truncation / log_prob [R], raw_action [R, nu],
        eps_z [R, L] (the encoder's reparameterisation noise, `policy_rng`), eps_ent [R, nu] (the entropy sample, `rng`).
        Returns the metrics tensor (device; names in METRIC_NAMES); gradients are left in self.grads."""
        t, L_, ws, p, g, R, Bm, nu = self.torch, tk.lib(), self.ws, self.p, self.g, self.R, self.Bm, self.nu
        st = tk.stream(self.params)
        hp = self.hp
        P = lambda k: p["policy/" + k]
        G = lambda k: g["policy/" + k]
        chk = tk.check
        ptr = lambda x: x.data_ptr()
        self.grads.zero_()
        ws["metrics"].zero_()
        traj, obs = batch["traj"], batch["observation"]
        assert traj.shape == (R, self.ld_traj) and obs.shape == (R, self.obs) and traj.is_contiguous() and obs.is_contiguous()
        # ---- forward: policy --------------------------------------------------------------------------------------------
        Ld = self.L + self.obs
        # The two networks are independent up to the loss rows and again after them: the value network runs on `self.branch`
        # (fork / join by stream events, which a CUDA-graph capture records as two parallel chains), so the policy's small,
        # latency-bound GEMMs (40-80 CTAs each) fill the SMs the value GEMMs leave idle between their waves.
        main = t.cuda.current_stream(self.device)
        V = lambda k: p["value/" + k]
        GV = lambda k: g["value/" + k]
        RB = R + Bm
        self.branch.wait_stream(main)
        with t.cuda.stream(self.branch):
            sb = tk.stream(self.params)
            chk(L_.vnl_obs_normalize(ptr(obs), self.obs, R, self.obs, ptr(self.obs_mean), ptr(self.obs_std), ptr(ws["vin"]), self.obs, sb), "normalize")
            chk(L_.vnl_obs_normalize(ptr(batch["next_observation_last"]), self.obs, Bm, self.obs, ptr(self.obs_mean), ptr(self.obs_std),
                                     ptr(ws["vin"]) + 4 * R * self.obs, self.obs, sb), "normalize")
            # value forward (baseline rows + the Bm bootstrap rows in one pass)
            self._fwd(ws["vin"], RB, V("hidden_0/kernel"), V("hidden_0/bias"), ws["v0pre"], act=ws["v0"])  # swish in the GEMM epilogue
            self._fwd(ws["v0"], RB, V("hidden_1/kernel"), V("hidden_1/bias"), ws["v1pre"], act=ws["v1"])
            chk(L_.vnl_rowdot(ptr(ws["v1"]), self.vh[1], RB, self.vh[1], ptr(V("hidden_2/kernel")), ptr(V("hidden_2/bias")), ptr(ws["val"]), sb), "rowdot")
        chk(L_.vnl_obs_normalize(ptr(obs), self.obs, R, self.obs, ptr(self.obs_mean), ptr(self.obs_std), ptr(ws["dec_in"]) + 4 * self.L, Ld, st), "normalize")

        def relu_ln(pre, name, out, stats):
            n = pre.shape[1]
            chk(L_.vnl_relu_ln_fwd(ptr(pre), n, R, n, ptr(P(name + "/scale")), ptr(P(name + "/bias")), ptr(out), n, ptr(stats), st), "relu_ln_fwd")
        self._fwd(traj, R, P("encoder/hidden_0/kernel"), P("encoder/hidden_0/bias"), ws["h0pre"])
        relu_ln(ws["h0pre"], "encoder/LayerNorm_0", ws["h0"], ws["s0"])
        self._fwd(ws["h0"], R, P("encoder/hidden_1/kernel"), P("encoder/hidden_1/bias"), ws["h1pre"])
        relu_ln(ws["h1pre"], "encoder/LayerNorm_1", ws["h1"], ws["s1"])
        self._fwd(ws["h1"], R, P("encoder/heads/kernel"), P("encoder/heads/bias"), ws["heads"])
        chk(L_.vnl_reparam_fwd(ptr(ws["heads"]), ptr(batch["eps_z"]), R, self.L, ptr(ws["dec_in"]), Ld, st), "reparam")
        self._fwd(ws["dec_in"], R, P("decoder/hidden_0/kernel"), P("decoder/hidden_0/bias"), ws["d0pre"])
        relu_ln(ws["d0pre"], "decoder/LayerNorm_0", ws["d0"], ws["s2"])
        self._fwd(ws["d0"], R, P("decoder/hidden_1/kernel"), P("decoder/hidden_1/bias"), ws["d1pre"])
        relu_ln(ws["d1pre"], "decoder/LayerNorm_1", ws["d1"], ws["s3"])
        self._fwd(ws["d1"], R, P("decoder/hidden_2/kernel"), P("decoder/hidden_2/bias"), ws["logits"])
        # ---- loss -------------------------------------------------------------------------------------------------------------
        chk(L_.vnl_ppo_rows(ptr(ws["logits"]), 2 * nu, ptr(batch["raw_action"]), ptr(batch["eps_ent"]), R, nu, ptr(batch["discount"]),
                            ptr(batch["truncation"]), ptr(batch["reward"]), float(hp["reward_scaling"]), ptr(ws["target_lp"]), ptr(ws["ent"]),
                            ptr(ws["termination"]), ptr(ws["rewards_s"]), st), "ppo_rows")
        main.wait_stream(self.branch)  # join: GAE needs the values
        chk(self._gae.vnl_gae(self.T, Bm, ptr(batch["truncation"]), ptr(ws["termination"]), ptr(ws["rewards_s"]), ptr(ws["val"]), ptr(ws["val"]) + 4 * R,
                         float(hp["gae_lambda"]), float(hp["discounting"]), ptr(ws["vs"]), ptr(ws["adv"]), st), "gae")
        chk(L_.vnl_ppo_loss_bwd(ptr(ws["logits"]), 2 * nu, ptr(batch["raw_action"]), ptr(batch["eps_ent"]), R, nu, ptr(ws["target_lp"]),
                                ptr(batch["log_prob"]), ptr(ws["ent"]), ptr(ws["adv"]), ptr(ws["vs"]), ptr(ws["val"]), float(hp["clipping_epsilon"]),
                                float(hp["entropy_cost"]), int(bool(hp["normalize_advantage"])), ptr(ws["dlogits"]), 2 * nu, ptr(ws["dval"]),
                                ptr(ws["metrics"]), ptr(ws["scratch2"]), st), "ppo_loss_bwd")
        self.launches += 13  # row kernels of the forward pass and the loss (GEMMs count themselves)
        colsum = lambda x, n, out, w=None: chk(L_.vnl_colsum(ptr(x), x.shape[1] if x.dim() == 2 else n, R, n, None if w is None else ptr(w), ptr(out),
                                                             tk.stream(self.params)), "colsum")
        # ---- backward ------------------------------------------------------------------------------------------------------------
        # Each network's backward is a serial CHAIN (dgrad -> activation backward -> dgrad ...) with LEAVES hanging off it (the weight
        # and bias gradients of every layer, which nothing downstream reads).  Chains run on `main` (policy) and `self.branch`
        # (value); leaves go to `leaf_p` / `leaf_v` as soon as their operand exists, so a chain never queues behind a leaf.
        def leaf(stream, after, fn):
            stream.wait_stream(after)
            with t.cuda.stream(stream):
                fn()

        # value (only the R baseline rows carry gradient; the bootstrap rows feed the stop-gradient GAE)
        self.branch.wait_stream(main)

        def value_head_leaves():
            sl = tk.stream(self.params)
            chk(L_.vnl_colsum(ptr(ws["v1"]), self.vh[1], R, self.vh[1], ptr(ws["dval"]), ptr(GV("hidden_2/kernel")), sl), "colsum")
            chk(L_.vnl_colsum(ptr(ws["dval"]), 1, R, 1, None, ptr(GV("hidden_2/bias")), sl), "colsum")
        leaf(self.leaf_v, main, value_head_leaves)
        with t.cuda.stream(self.branch):
            sb = tk.stream(self.params)
            chk(L_.vnl_outer_swish_bwd(ptr(ws["dval"]), R, ptr(V("hidden_2/kernel")), self.vh[1], ptr(ws["v1pre"]), ptr(ws["dv1pre"]), sb), "outer_swish_bwd")
        leaf(self.leaf_v, self.branch, lambda: (self._wgrad(ws["v0"], ws["dv1pre"], R, GV("hidden_1/kernel")),
                                                colsum(ws["dv1pre"], self.vh[1], GV("hidden_1/bias"))))
        with t.cuda.stream(self.branch):
            self._dgrad(ws["dv1pre"], R, V("hidden_1/kernel"), ws["dv0pre"], pre=ws["v0pre"], tmp=ws["dv0"])  # swish' in the GEMM epilogue
            self._wgrad(ws["vin"], ws["dv0pre"], R, GV("hidden_0/kernel"))
            colsum(ws["dv0pre"], self.vh[0], GV("hidden_0/bias"))

        # policy (its gradient bucket goes to the exchange as soon as it is complete)
        def relu_ln_bwd(dy, pre, stats, name, dpre, dense):  # also leaves the bias gradient of the dense layer in front (column sums of dpre)
            n = pre.shape[1]
            chk(L_.vnl_relu_ln_bwd(ptr(dy), n, ptr(pre), n, ptr(stats), ptr(P(name + "/scale")), R, n, ptr(dpre), n, ptr(G(name + "/scale")),
                                   ptr(G(name + "/bias")), ptr(G(dense + "/bias")), st), "relu_ln_bwd")
        lp = lambda fn: leaf(self.leaf_p, main, fn)
        lp(lambda: (self._wgrad(ws["d1"], ws["dlogits"], R, G("decoder/hidden_2/kernel")), colsum(ws["dlogits"], 2 * nu, G("decoder/hidden_2/bias"))))
        self._dgrad(ws["dlogits"], R, P("decoder/hidden_2/kernel"), ws["dd1"])
        relu_ln_bwd(ws["dd1"], ws["d1pre"], ws["s3"], "decoder/LayerNorm_1", ws["dd1pre"], "decoder/hidden_1")
        lp(lambda: self._wgrad(ws["d0"], ws["dd1pre"], R, G("decoder/hidden_1/kernel")))
        self._dgrad(ws["dd1pre"], R, P("decoder/hidden_1/kernel"), ws["dd0"])
        relu_ln_bwd(ws["dd0"], ws["d0pre"], ws["s2"], "decoder/LayerNorm_0", ws["dd0pre"], "decoder/hidden_0")
        lp(lambda: self._wgrad(ws["dec_in"], ws["dd0pre"], R, G("decoder/hidden_0/kernel")))
        self._dgrad(ws["dd0pre"], R, P("decoder/hidden_0/kernel"), ws["ddec_in"])
        kl_coef = float(hp["kl_weight"]) / float(R * self.L)
        chk(L_.vnl_heads_bwd(ptr(ws["ddec_in"]), Ld, ptr(ws["heads"]), ptr(batch["eps_z"]), R, self.L, kl_coef, ptr(ws["dheads"]),
                             ptr(ws["metrics"]) + 16, st), "heads_bwd")
        lp(lambda: (self._wgrad(ws["h1"], ws["dheads"], R, G("encoder/heads/kernel")), colsum(ws["dheads"], 2 * self.L, G("encoder/heads/bias"))))
        self._dgrad(ws["dheads"], R, P("encoder/heads/kernel"), ws["dh1"])
        relu_ln_bwd(ws["dh1"], ws["h1pre"], ws["s1"], "encoder/LayerNorm_1", ws["dh1pre"], "encoder/hidden_1")
        lp(lambda: self._wgrad(ws["h0"], ws["dh1pre"], R, G("encoder/hidden_1/kernel")))
        self._dgrad(ws["dh1pre"], R, P("encoder/hidden_1/kernel"), ws["dh0"])
        relu_ln_bwd(ws["dh0"], ws["h0pre"], ws["s0"], "encoder/LayerNorm_0", ws["dh0pre"], "encoder/hidden_0")
        self._wgrad(traj, ws["dh0pre"], R, G("encoder/hidden_0/kernel"))
        main.wait_stream(self.leaf_p)  # the policy bucket is complete
        self._pending = None
        if exchange:
            self._policy_bucket_ready()
        main.wait_stream(self.branch)  # join: all gradients are in self.grads
        main.wait_stream(self.leaf_v)
        self.launches += 12  # row kernels of the backward pass
        return ws["metrics"]

    # ---- gradient exchange + optimiser ---------------------------------------------------------------------------------------
    def _dist(self):
        import torch.distributed as dist
        return dist if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None

    def _policy_bucket_ready(self):
        """`pmean` of the policy gradients (gradients.gradient_update_fn, ppo_imitation/train.py:251) starts on a side stream as soon
        as the policy backward has been enqueued; the value backward runs beside it."""
        self._pending = None
        dist = self._dist()
        if dist is None:
            return
        t = self.torch
        self.side.wait_stream(t.cuda.current_stream(self.device))
        with t.cuda.stream(self.side):
            self._pending = start_bucket(dist, self.grads, 0, self.n_policy)

    def apply_gradients(self, exchange: bool = True, scale: float = 1.0):
        """value bucket all-reduce + join the policy bucket, then optax.adam on the flat buffers (mean over ranks folded into Adam).
        `exchange=False`: the caller has already all-reduced self.grads (trainer.Trainer: eager NCCL call between two CUDA graphs)
        and passes `scale` = 1 / world."""
        t, L_ = self.torch, tk.lib()
        dist = self._dist() if exchange else None
        if dist is not None:
            scale = finish_buckets(dist, self.grads, self.n_policy, self._pending)
            t.cuda.current_stream(self.device).wait_stream(self.side)
            self._pending = None
        st = tk.stream(self.params)
        hp = self.hp
        tk.check(L_.vnl_adam_tick(self.step_dev.data_ptr(), float(hp["b1"]), float(hp["b2"]), self.bc_dev.data_ptr(), st), "adam_tick")
        tk.check(L_.vnl_adam(self.params.data_ptr(), self.grads.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.nparams, float(hp["learning_rate"]),
                             float(hp["b1"]), float(hp["b2"]), float(hp["eps"]), 0, float(scale), self.bc_dev.data_ptr(), st), "adam")
        self.updates += 1
        self.launches += 2

    def update(self, batch):
        m = self.loss_and_grads(batch)
        self.apply_gradients()
        return m

    def metrics_dict(self, m=None) -> Dict[str, float]:
        a = (self.ws["metrics"] if m is None else m).detach().cpu().numpy()
        out = {k: float(a[i]) for i, k in enumerate(METRIC_NAMES)}
        out["total_loss"] = float(a[0] + a[4])  # the KL term is accumulated by its own kernel
        return out


# -------------------------------------------------------------------------------------------------------------------------------
# gradient exchange (device-agnostic: the 2-rank gloo test on CPU drives the same two functions)
# -------------------------------------------------------------------------------------------------------------------------------
def start_bucket(dist, grads, lo: int, hi: int):
    """Asynchronous SUM all-reduce of grads[lo:hi] in place (the policy bucket, started while the value backward still runs)."""
    return dist.all_reduce(grads[lo:hi], op=dist.ReduceOp.SUM, async_op=True)


def finish_buckets(dist, grads, n_first: int, pending) -> float:
    """All-reduces the remaining bucket grads[n_first:], joins the pending one and returns the factor that turns the SUMs into
    `lax.pmean` (gradients.gradient_update_fn with pmap_axis_name): Adam applies it (vnl_adam grad_scale)."""
    dist.all_reduce(grads[n_first:], op=dist.ReduceOp.SUM)
    if pending is not None:
        pending.wait()
    return 1.0 / dist.get_world_size()


# -------------------------------------------------------------------------------------------------------------------------------
# checker: torch-autograd restatement of the reference loss (tests / tools only)
# -------------------------------------------------------------------------------------------------------------------------------
def reference_loss(policy_params, value_params, batch, T: int, Bm: int, obs_mean, obs_std, *, entropy_cost=1e-4, discounting=0.9,
                   reward_scaling=1.0, gae_lambda=0.95, clipping_epsilon=0.3, normalize_advantage=True, kl_weight=1e-4, dtype=None):
    """`compute_ppo_intention_loss` (ppo_imitation/intention_losses.py:91-202) restated in torch with autograd, line by line:
    policy_apply (intention_policy_network.py:82-105 with the reparameterisation noise `eps_z`), value_apply (brax MLP, swish),
    NormalTanhDistribution.log_prob / entropy (sample from `eps_ent`), compute_gae (:26-89, stop-gradient), advantage
    normalisation, clipped surrogate, v_loss * 0.5 * 0.5, entropy_cost, kl_weight * kl_divergence.
    Returns (total, metrics dict, policy grads dict, value grads dict) with parameters as leaf tensors."""
    import torch

    dtype = dtype or torch.float64
    dev = batch["observation"].device
    Pp = {k: torch.as_tensor(v, dtype=dtype, device=dev).clone().requires_grad_(True) for k, v in policy_params.items()}
    Vp = {k: torch.as_tensor(v, dtype=dtype, device=dev).clone().requires_grad_(True) for k, v in value_params.items()}
    R = T * Bm
    c = lambda x: x.to(dtype)
    mean, std = c(torch.as_tensor(obs_mean, device=dev)), c(torch.as_tensor(obs_std, device=dev))
    traj = c(batch["traj"])[:, :Pp["encoder/hidden_0/kernel"].shape[0]]
    obs_n = (c(batch["observation"]) - mean) / std
    nobs_n = (c(batch["next_observation_last"]) - mean) / std

    def ln(x, name):
        m = x.mean(-1, keepdim=True)
        var = torch.clamp((x * x).mean(-1, keepdim=True) - m * m, min=0.0)
        return (x - m) * torch.rsqrt(var + 1e-6) * Pp[name + "/scale"] + Pp[name + "/bias"]
    dense = lambda x, name: x @ Pp[name + "/kernel"] + Pp[name + "/bias"]
    h = traj
    pres = []
    for i in range(2):
        pres.append(dense(h, f"encoder/hidden_{i}"))
        h = ln(torch.relu(pres[-1]), f"encoder/LayerNorm_{i}")
    zm, zlv = dense(h, "encoder/fc2_mean"), dense(h, "encoder/fc2_logvar")
    z = zm + c(batch["eps_z"]) * torch.exp(0.5 * zlv)
    h = torch.cat([z, obs_n], -1)
    for i in range(2):
        pres.append(dense(h, f"decoder/hidden_{i}"))
        h = ln(torch.relu(pres[-1]), f"decoder/LayerNorm_{i}")
    logits = dense(h, "decoder/hidden_2")
    logits.retain_grad()

    def value(x):
        nl = len(Vp) // 2
        for i in range(nl):
            x = x @ Vp[f"hidden_{i}/kernel"] + Vp[f"hidden_{i}/bias"]
            if i < nl - 1:
                x = x * torch.sigmoid(x)
        return x.squeeze(-1)
    baseline = value(obs_n).reshape(T, Bm)
    baseline.retain_grad()
    bootstrap = value(nobs_n)
    tm = lambda k: c(batch[k]).reshape(T, Bm)
    rewards = tm("reward") * reward_scaling
    truncation = tm("truncation")
    termination = (1 - tm("discount")) * (1 - truncation)
    nu = logits.shape[-1] // 2
    loc, scale = logits[..., :nu], torch.nn.functional.softplus(logits[..., nu:]) + 0.001
    ldj = lambda x: 2.0 * (math.log(2.0) - x - torch.nn.functional.softplus(-2.0 * x))
    raw = c(batch["raw_action"])
    target_lp = (-0.5 * ((raw - loc) / scale) ** 2 - torch.log(scale) - 0.5 * math.log(2 * math.pi) - ldj(raw)).sum(-1).reshape(T, Bm)
    behaviour_lp = tm("log_prob")
    # compute_gae (stop-gradient)
    with torch.no_grad():
        vals = baseline.detach()
        tmask = 1 - truncation
        v_tp1 = torch.cat([vals[1:], bootstrap.detach()[None]], 0)
        deltas = (rewards + discounting * (1 - termination) * v_tp1 - vals) * tmask
        acc = torch.zeros_like(bootstrap)
        out = []
        for t_ in range(T - 1, -1, -1):
            acc = deltas[t_] + discounting * (1 - termination[t_]) * tmask[t_] * gae_lambda * acc
            out.append(acc)
        vs = torch.stack(out[::-1]) + vals
        vs_tp1 = torch.cat([vs[1:], bootstrap.detach()[None]], 0)
        adv = (rewards + discounting * (1 - termination) * vs_tp1 - vals) * tmask
        if normalize_advantage:
            adv = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)
    rho = torch.exp(target_lp - behaviour_lp)
    policy_loss = -torch.minimum(rho * adv, torch.clamp(rho, 1 - clipping_epsilon, 1 + clipping_epsilon) * adv).mean()
    v_loss = ((vs - baseline) ** 2).mean() * 0.5 * 0.5
    ent = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(scale) + ldj(loc + scale * c(batch["eps_ent"]))).sum(-1)
    entropy_loss = entropy_cost * -ent.mean()
    kl = kl_weight * (-0.5 * torch.mean(1 + zlv - zm ** 2 - torch.exp(zlv)))
    total = policy_loss + v_loss + entropy_loss + kl
    total.backward()
    f_ = lambda x: float(x.detach())
    metrics = dict(total_loss=f_(total), policy_loss=f_(policy_loss), v_loss=f_(v_loss), entropy_loss=f_(entropy_loss),
                   kl_loss_intention=f_(kl), mean_rho=f_(rho.mean()))
    aux = dict(logits=logits.detach(), baseline=baseline.detach(), vs=vs, adv=adv, target_lp=target_lp.detach(), dlogits=logits.grad,
               dbaseline=baseline.grad, rho=rho.detach(), pre=[x.detach() for x in pres])
    return total, metrics, {k: v.grad for k, v in Pp.items()}, {k: v.grad for k, v in Vp.items()}, aux


def reference_adam(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam update restated (scale_by_adam + scale(-lr)); returns (p', m', v')."""
    m2, v2 = b1 * m + (1 - b1) * g, b2 * v + (1 - b2) * g * g
    return p - lr * (m2 / (1 - b1 ** step)) / ((v2 / (1 - b2 ** step)) ** 0.5 + eps), m2, v2

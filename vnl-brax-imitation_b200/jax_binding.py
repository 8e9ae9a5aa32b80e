"""JAX / brax side of the boundary: `RodentTracking` as a brax `Env` whose `reset` / `step` lower to the XLA custom calls
`vnl_xla_reset` / `vnl_xla_step` of libvnl_b200.so (status-returning legacy ABI, include/vnl_b200.h).

This is the module a maintainer of the reference imports INSTEAD of `envs/rodent.py`:

    from vnl_b200.jax_binding import RodentTracking            # same constructor arguments (envs/rodent.py:15-38)
    envs.register_environment("rodent", RodentTracking)         # reference train.py:65-68, unchanged
    env = envs.get_environment("rodent", reference_clip=clip, **env_args)
    ppo.train(environment=env, ...)                             # ppo_imitation/train.py, unchanged

What it provides (line by line):
  * un-batched `reset(rng) -> State`, `step(state, action) -> State` with the reference's pytree: `State(pipeline_state,
    obs, reward, done, metrics{rcom..termination_error}, info{cur_frame, sub_clip_frame, traj, termination_error})`;
    `pipeline_state` is a flax struct of batch-leading arrays exposing `.qpos/.qvel/.q/.qd/.xpos/...` (train.py:293,
    AutoResetWrapper's `tree_map(where(done, first, cur))`);
  * primitives `vnl_step_p` / `vnl_reset_p` with abstract evaluation, CUDA lowering and a BATCHING RULE that flattens any
    number of nested `vmap`s into the one leading env axis B of the C ABI (`VmapWrapper` + the outer `vmap` of
    ppo_imitation/train.py:215); the un-batched call (reference train.py:160,176) is B = 1; `lax.scan`, `jit` and `pmap`
    need nothing more (the call is a pure function of its operands);
  * the model / task blobs are ordinary device operands (constants closed over by the env), the launch geometry travels in
    the custom call's `opaque` (VnlXlaOpaque), the inertia workspace is a scratch RESULT: the library keeps no state, so
    XLA may copy / donate / replicate anything;
  * failures surface as XLA errors (`XlaCustomCallStatusSetFailure`), never as silently uninitialised results.

JAX (>= 0.4.14; the reference's era is 0.4.26-0.4.28, brax 0.10.x) is NOT installed in the build image of this repo, so
this module is import-guarded and its jax-dependent part has never been executed there (INTEGRATION.md says so too).  The
pieces that do not need jax -- operand order, result specs, the batch-flattening plan, the opaque bytes -- are plain
functions below and are covered by tests/test_jax_binding.py; the custom calls themselves are driven through the very
same (stream, buffers, opaque, status) ABI by tests/test_xla_boundary.py on the GPU.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from . import model_blob as mb
from .envs import rodent as _rodent

try:  # pragma: no cover - jax is absent in the build image
    import jax
    import jax.numpy as jp
    HAVE_JAX = True
except Exception:  # noqa: BLE001
    jax = jp = None
    HAVE_JAX = False

# ---------------------------------------------------------------------------------------------------------------------
# jax-free part: the operand / result plan of the custom calls (exactly the buffer list of include/vnl_b200.h)
# ---------------------------------------------------------------------------------------------------------------------
STATE_FIELDS = _lib.STATE_F + _lib.STATE_I  # qpos .. qfrc_actuator (f32), cur_frame, sub_clip_frame, clip_id (i32)
OUTPUT_FIELDS = ("obs", "traj", "reward", "done", "metrics", "stats")
OPERANDS = ("model", "task") + STATE_FIELDS + ("action",)
RESULTS = tuple(k + "_out" for k in STATE_FIELDS) + OUTPUT_FIELDS + ("workspace",)
assert len(OPERANDS) + len(RESULTS) == _lib.VNL_XLA_STEP_NBUF


def field_shapes(dims: Dict[str, int], obs_size: int, traj_size: int) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """Per-env (un-batched) shape and dtype of every state / output leaf."""
    nq, nv, na, nb = dims["nq"], dims["nv"], dims["na"], dims["nbody"]
    f, i = "float32", "int32"
    return {"qpos": ((nq,), f), "qvel": ((nv,), f), "act": ((na,), f), "qacc_warmstart": ((nv,), f), "xpos": ((nb, 3), f),
            "xquat": ((nb, 4), f), "subtree_com": ((3,), f), "qfrc_actuator": ((nv,), f), "cur_frame": ((), i),
            "sub_clip_frame": ((), i), "clip_id": ((), i), "action": ((dims["nu"],), f), "obs": ((obs_size,), f),
            "traj": ((traj_size,), f), "reward": ((), f), "done": ((), f), "metrics": ((7,), f), "stats": ((4,), i)}


def result_specs(B: int, dims: Dict[str, int], obs_size: int, traj_size: int, workspace_bytes: int) -> List[Tuple[Tuple[int, ...], str]]:
    """(shape, dtype) of every result buffer of one call over B envs, in RESULTS order; the workspace is a flat f32 scratch."""
    sh = field_shapes(dims, obs_size, traj_size)
    out = [((B,) + sh[k][0], sh[k][1]) for k in STATE_FIELDS] + [((B,) + sh[k][0], sh[k][1]) for k in OUTPUT_FIELDS]
    return out + [((max(int(workspace_bytes), 4) // 4,), "float32")]


def flatten_plan(shapes: Sequence[Tuple[int, ...]], batch_dims: Sequence[Optional[int]], core_ndims: Sequence[int]):
    """The batching rule as data.  Operand i has shape `shapes[i]` = [B, *core] (already carrying the primitive's own env
    axis), is mapped by the enclosing `vmap` along `batch_dims[i]` (None = not mapped) and has `core_ndims[i]` trailing
    per-env dims.  Returns (N, per-operand recipe): move the mapped axis to the front and merge it with the env axis
    ([N, B, *core] -> [N * B, *core]); an unmapped operand is broadcast to N first.  Results come back as [N * B, *core] and
    are reshaped to [N, B, *core] with the new batch axis at 0.  Nested vmaps apply the rule repeatedly, which is what
    flattens `vmap(vmap(env.step))` (ppo_imitation/train.py:215 over VmapWrapper) into one launch."""
    sizes = {s[bd] for s, bd in zip(shapes, batch_dims) if bd is not None}
    if len(sizes) != 1:
        raise ValueError(f"inconsistent mapped axis sizes {sizes}")
    N = sizes.pop()
    plan = []
    for s, bd, cn in zip(shapes, batch_dims, core_ndims):
        if len(s) != 1 + cn + (0 if bd is None else 1):
            raise ValueError(f"operand of shape {s} is not [B, core({cn})] (+ mapped axis)")
        if bd is None:
            plan.append(("broadcast", N, (N * s[0],) + tuple(s[1:])))
        else:
            rest = tuple(d for a, d in enumerate(s) if a != bd)
            plan.append(("move", bd, (N * rest[0],) + rest[1:]))
    return N, plan


class Binding:
    """Host-side constants of one env: blobs, call context, opaque factory.  jax-free (tested on CPU)."""

    def __init__(self, model_blob: np.ndarray, task_blob: np.ndarray, lib=None):
        self.lib = lib or _lib.load_library()
        self.model_host = np.ascontiguousarray(model_blob, dtype=np.uint32)
        self.task_host = np.ascontiguousarray(task_blob, dtype=np.uint32)
        self.ctx = _lib.VnlContext()
        rc = self.lib.vnl_context_init(ctypes.byref(self.ctx), self.model_host.ctypes.data, self.model_host.nbytes,
                                       self.task_host.ctypes.data, self.task_host.nbytes)
        if rc:
            raise ValueError(f"vnl_context_init failed ({rc})")
        self.dims = mb.read_dims(self.model_host)
        self.obs_size = int(self.task_host[mb.C["VNL_TH_OBS_SIZE"]])
        self.traj_size = int(self.task_host[mb.C["VNL_TH_TRAJ_SIZE"]])
        # vnl_workspace_bytes asks the CUDA runtime for the SM count of the current device (148 when there is none)
        self.workspace_bytes = int(self.lib.vnl_workspace_bytes(self.model_host.ctypes.data))
        self.ctx.workspace_bytes = self.workspace_bytes

    def opaque(self, B: int) -> bytes:
        op = _lib.VnlXlaOpaque()
        rc = self.lib.vnl_xla_make_opaque(ctypes.byref(self.ctx), int(B), ctypes.byref(op))
        if rc:
            raise ValueError(f"vnl_xla_make_opaque failed ({rc})")
        return bytes(op)

    def targets(self) -> Dict[str, int]:
        """name -> function address of the custom-call targets to register with XLA (platform CUDA, api_version
        STATUS_RETURNING for the legacy registration call)."""
        names = ("vnl_xla_step", "vnl_xla_reset", "vnl_xla_policy_forward", "vnl_xla_gae", "vnl_xla_obs_stats_partial",
                 "vnl_xla_obs_stats_finish")
        return {n: ctypes.cast(getattr(self.lib, n), ctypes.c_void_p).value for n in names}


# ---------------------------------------------------------------------------------------------------------------------
# jax part
# ---------------------------------------------------------------------------------------------------------------------
if HAVE_JAX:  # pragma: no cover - not executable in the build image (no jax); see the module docstring
    import functools

    from flax import struct
    from jax.interpreters import batching, mlir
    from jax.interpreters.mlir import ir

    try:
        from jax.extend.core import Primitive  # jax >= 0.4.36
    except Exception:  # noqa: BLE001
        from jax.core import Primitive
    from jax.core import ShapedArray

    from brax.envs.base import Env, State

    _REGISTERED = False

    def _capsule(addr: int):
        ctypes.pythonapi.PyCapsule_New.restype = ctypes.py_object
        ctypes.pythonapi.PyCapsule_New.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        return ctypes.pythonapi.PyCapsule_New(addr, b"xla._CUSTOM_CALL_TARGET", None)

    def register_custom_calls(binding: Binding) -> None:
        """Registers every `vnl_xla_*` symbol as a CUDA custom-call target (legacy ABI; the lowering marks the calls
        api_version = 2, STATUS_RETURNING, so XLA passes the status pointer)."""
        global _REGISTERED
        if _REGISTERED:
            return
        for name, addr in binding.targets().items():
            cap = _capsule(addr)
            try:  # jax >= 0.4.31
                jax.ffi.register_ffi_target(name, cap, platform="CUDA", api_version=0)
            except Exception:  # noqa: BLE001  (reference era: jaxlib xla_client)
                from jax.lib import xla_client
                xla_client.register_custom_call_target(name.encode(), cap, platform="CUDA")
        _REGISTERED = True

    @struct.dataclass
    class VnlPipelineState:
        """The leaves of `mjx.Data` that the reference's callers read, batch-leading like every brax pytree.
        `xpos / xquat / subtree_com / qfrc_actuator` lag qpos by one substep exactly as `mjx.step` leaves them."""
        qpos: jax.Array
        qvel: jax.Array
        act: jax.Array
        qacc_warmstart: jax.Array
        xpos: jax.Array
        xquat: jax.Array
        subtree_com: jax.Array
        qfrc_actuator: jax.Array

        @property
        def q(self):  # brax aliases (envs/rodent.py:314, train.py:293)
            return self.qpos

        @property
        def qd(self):
            return self.qvel

    def _make_primitive(name: str, target: str, binding: Binding, has_action: bool):
        prim = Primitive(name)
        prim.multiple_results = True
        sh = field_shapes(binding.dims, binding.obs_size, binding.traj_size)
        operand_names = OPERANDS if has_action else OPERANDS[:-1]
        core_nd = [1, 1] + [len(sh[k][0]) for k in STATE_FIELDS] + ([len(sh["action"][0])] if has_action else [])

        def abstract(*args):
            B = args[2].shape[0]
            return [ShapedArray(s, np.dtype(d)) for s, d in result_specs(B, binding.dims, binding.obs_size, binding.traj_size,
                                                                         binding.workspace_bytes)]

        def lowering(ctx, *args):
            B = ctx.avals_in[2].shape[0]
            operands = list(args)
            if not has_action:  # the buffer list is fixed: the reset call passes a dummy action operand
                operands.append(mlir.ir_constant(np.zeros((B, binding.dims["nu"]), np.float32)))
            avals = list(ctx.avals_in) + ([] if has_action else [ShapedArray((B, binding.dims["nu"]), np.float32)])
            row_major = lambda a: tuple(range(a.ndim - 1, -1, -1))
            call = mlir.custom_call(target, result_types=[mlir.aval_to_ir_type(a) for a in ctx.avals_out], operands=operands,
                                    backend_config=binding.opaque(B), api_version=2,
                                    operand_layouts=[row_major(a) for a in avals], result_layouts=[row_major(a) for a in ctx.avals_out])
            return call.results

        def batch(args, dims_):
            # model / task blobs (operands 0, 1) are closed-over constants: never mapped
            N, plan = flatten_plan([a.shape for a in args[2:]], list(dims_[2:]), core_nd[2:])
            new = [args[0], args[1]]
            for a, (kind, arg, shape) in zip(args[2:], plan):
                if kind == "move":
                    new.append(jp.reshape(jp.moveaxis(a, arg, 0), shape))
                else:
                    new.append(jp.reshape(jp.broadcast_to(a[None], (N,) + a.shape), shape))
            outs = prim.bind(*new)
            res = [jp.reshape(o, (N, o.shape[0] // N) + o.shape[1:]) for o in outs[:-1]] + [outs[-1]]
            return res, [0] * (len(outs) - 1) + [None]

        prim.def_impl(functools.partial(jax.interpreters.xla.apply_primitive, prim))
        prim.def_abstract_eval(abstract)
        mlir.register_lowering(prim, lowering, platform="cuda")
        batching.primitive_batchers[prim] = batch
        del operand_names
        return prim

    class RodentTracking(Env):
        """Drop-in for the reference's `RodentTracking` (envs/rodent.py:14-470): same constructor, same `State`."""

        def __init__(self, reference_clip, end_eff_names, appendage_names, walker_body_names, joint_names, center_of_mass,
                     mjcf_path: str = "./assets/rodent.xml", scale_factor: float = 0.9, solver: str = "cg", iterations: int = 6,
                     ls_iterations: int = 6, healthy_z_range=(0.05, 0.5), reset_noise_scale=1e-3, clip_length: int = 250,
                     sub_clip_length: int = 10, ref_traj_length: int = 5, termination_threshold: float = 5,
                     body_error_multiplier: float = 1.0, **kwargs):
            from . import mjcf
            if sub_clip_length > clip_length:
                raise ValueError("episode_length cannot be greater than clip_length!")
            model = kwargs.pop("model", None) or mjcf.load_rodent(mjcf_path, scale_factor, solver, iterations, ls_iterations)
            self._n_frames = int(kwargs.get("n_frames", 5))
            host_clip = jax.tree_util.tree_map(np.asarray, reference_clip)
            task_blob, fclip, idx, self._obs_size, self._traj_size = _rodent.rodent_task_tables(
                model, host_clip, end_eff_names=end_eff_names, appendage_names=appendage_names,
                walker_body_names=walker_body_names, joint_names=joint_names, center_of_mass=center_of_mass,
                clip_length=clip_length, sub_clip_length=sub_clip_length, ref_traj_length=ref_traj_length,
                termination_threshold=termination_threshold, body_error_multiplier=body_error_multiplier,
                healthy_z_range=healthy_z_range, n_frames=self._n_frames)
            self.binding = Binding(mb.build_model_blob(model), task_blob)
            register_custom_calls(self.binding)
            self._model_dev = jp.asarray(self.binding.model_host.view(np.int32))
            self._task_dev = jp.asarray(self.binding.task_host.view(np.int32))
            self._step_p = _make_primitive("vnl_step", "vnl_xla_step", self.binding, True)
            self._reset_p = _make_primitive("vnl_reset", "vnl_xla_reset", self.binding, False)
            self._ref_traj = jax.tree_util.tree_map(jp.asarray, fclip)  # train.py:287-291 reads position / quaternion / joints
            self._model = model
            self._clip_length, self._sub_clip_length, self._ref_traj_length = clip_length, sub_clip_length, ref_traj_length
            self._reset_noise_scale = reset_noise_scale
            self.sys = _rodent._Sys(model, self._n_frames)  # .nq / .nv / .nu (rodent.py:132,156)

        # ---- brax Env surface -------------------------------------------------------------------------------------
        @property
        def dt(self):
            return self.sys.dt

        @property
        def action_size(self) -> int:
            return self.sys.nu

        @property
        def observation_size(self) -> int:
            return self._obs_size

        @property
        def backend(self) -> str:
            return "vnl_b200"

        def _call(self, prim, leaves: Dict[str, "jax.Array"], action=None):
            """Un-batched leaves -> [1, ...] operands -> primitive -> un-batched results (vmap adds the real env axis)."""
            sh = field_shapes(self.binding.dims, self._obs_size, self._traj_size)
            ops = [self._model_dev, self._task_dev]
            for k in STATE_FIELDS:
                v = leaves.get(k)
                v = jp.zeros(sh[k][0], sh[k][1]) if v is None else jp.asarray(v, sh[k][1])
                ops.append(v[None])
            if action is not None:
                ops.append(jp.asarray(action, jp.float32)[None])
            outs = prim.bind(*ops)
            named = dict(zip(RESULTS, outs))
            return {k: (v if k == "workspace" else v[0]) for k, v in named.items()}

        def _state(self, r, info_extra=None) -> State:
            ps = VnlPipelineState(**{k: r[k + "_out"] for k in _lib.STATE_F})
            m = r["metrics"]
            metrics = {k: m[i] for i, k in enumerate(_rodent.METRIC_KEYS)}
            info = {"cur_frame": r["cur_frame_out"], "sub_clip_frame": r["sub_clip_frame_out"], "traj": r["traj"],
                    "termination_error": m[6]}
            info.update(info_extra or {})
            return State(ps, r["obs"], r["reward"], r["done"], metrics, info)

        def reset(self, rng) -> State:
            """envs/rodent.py:119-176: the same two draws from the same keys, then ONE launch (`pipeline_init` + traj / obs /
            termination error)."""
            start_frame = jax.random.randint(rng, (), 0, self._clip_length - self._sub_clip_length - self._ref_traj_length)
            _, rng = jax.random.split(rng)
            noise = self._reset_noise_scale * jax.random.normal(rng, shape=(self.sys.nq,))
            rt = self._ref_traj
            qpos = jp.hstack([rt.position[start_frame, :], rt.quaternion[start_frame, :], rt.joints[start_frame, :]])
            qvel = jp.hstack([rt.velocity[start_frame, :], rt.angular_velocity[start_frame, :], rt.joints_velocity[start_frame, :]])
            r = self._call(self._reset_p, {"qpos": qpos + noise, "qvel": qvel, "cur_frame": start_frame})
            return self._state(r)

        def step(self, state: State, action) -> State:
            """envs/rodent.py:178-239: one launch."""
            ps = state.pipeline_state
            leaves = {k: getattr(ps, k) for k in _lib.STATE_F}
            leaves["cur_frame"], leaves["sub_clip_frame"] = state.info["cur_frame"], state.info["sub_clip_frame"]
            r = self._call(self._step_p, leaves, action)
            new = self._state(r)
            # the wrappers' own info entries (steps, truncation, first_pipeline_state, ...) ride along untouched
            info = dict(state.info)
            info.update(new.info)
            state.metrics.update(new.metrics)
            return state.replace(pipeline_state=new.pipeline_state, obs=new.obs, reward=new.reward, done=new.done, info=info)

    def register(name: str = "rodent") -> None:
        """`envs.register_environment("rodent", RodentTracking)` of reference train.py:65-68 with this class."""
        from brax import envs
        envs.register_environment(name, RodentTracking)

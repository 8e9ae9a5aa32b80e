"""jax-free part of vnl_b200.jax_binding (the module a reference maintainer imports instead of envs/rodent.py): it imports
without jax, its operand / result plan is the buffer list of include/vnl_b200.h, the batching rule flattens nested vmaps
into the one env axis of the C ABI, and the opaque bytes round-trip.  The jax-dependent part cannot run in this image
(no jax); the custom calls it lowers to are driven through the same ABI by tests/test_xla_boundary.py on the GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, pkg

jb = pkg("jax_binding")
libm = pkg("_lib")


def test_imports_without_jax_and_plan_matches_header():
    assert jb.HAVE_JAX in (True, False)
    hdr = open(os.path.join(ROOT, "include", "vnl_b200.h")).read()
    m = re.search(r"buffers: \[(.*?)\]\n \* vnl_xla_reset", hdr, flags=re.S)
    names = [n.strip() for n in re.sub(r"\(results\)|\*|\n", " ", m.group(1)).split(",")]
    alias = {"warm": "qacc_warmstart", "warm'": "qacc_warmstart_out"}
    want = [alias.get(n, n[:-1] + "_out" if n.endswith("'") else n) for n in names]
    assert want == list(jb.OPERANDS) + list(jb.RESULTS) and len(want) == libm.VNL_XLA_STEP_NBUF


def test_result_specs_and_opaque(rodent):
    b = jb.Binding(rodent["model_blob"], rodent["task_blob"])
    specs = jb.result_specs(6, b.dims, b.obs_size, b.traj_size, b.workspace_bytes)
    assert len(specs) == len(jb.RESULTS)
    d = dict(zip(jb.RESULTS, specs))
    assert d["qpos_out"] == ((6, 74), "float32") and d["xquat_out"] == ((6, 66, 4), "float32") and d["cur_frame_out"] == ((6,), "int32")
    assert d["obs"] == ((6, 232), "float32") and d["traj"] == ((6, 795), "float32") and d["stats"] == ((6, 4), "int32")
    assert d["workspace"][0][0] * 4 == b.workspace_bytes > 0
    op = libm.VnlXlaOpaque.from_buffer_copy(b.opaque(6))
    assert op.B == 6 and op.version == 2 and op.workspace_bytes == b.workspace_bytes
    assert list(op.model_hdr[:8]) == list(rodent["model_blob"][:8]) and op.task_hdr[0] == rodent["task_blob"][0]
    t = b.targets()
    assert set(t) >= {"vnl_xla_step", "vnl_xla_reset"} and all(v for v in t.values())


def test_batching_rule_flattens_nested_vmaps():
    # operands as the primitive sees them: [B, *core]; vmap over N adds one mapped axis somewhere
    core = [1, 1, 2, 0]                      # qpos-like, qvel-like, xpos-like, cur_frame-like
    shapes = [(1, 5, 74), (5, 1, 73), (1, 66, 5, 3), (1, 5)]
    bdims = [1, 0, 2, 1]
    N, plan = jb.flatten_plan(shapes, bdims, core)
    assert N == 5
    assert [p[0] for p in plan] == ["move"] * 4 and [p[2] for p in plan] == [(5, 74), (5, 73), (5, 66, 3), (5,)]
    # numpy emulation of the rule on real data: moving + merging keeps (n, b) -> n * B + b order
    x = np.arange(3 * 4 * 2).reshape(4, 3, 2)          # [B=4, N=3 mapped at axis 1, core 2]
    N, plan = jb.flatten_plan([x.shape], [1], [1])
    y = np.moveaxis(x, plan[0][1], 0).reshape(plan[0][2])
    assert y.shape == (12, 2) and np.array_equal(y.reshape(3, 4, 2)[2, 1], x[1, 2])
    # second (outer) vmap over the already flattened operand: [N2 mapped at 0, N*B, core]
    N2, plan2 = jb.flatten_plan([(7, 12, 2)], [0], [1])
    assert N2 == 7 and plan2[0][2] == (84, 2)
    # an unmapped operand (e.g. a broadcast action) is tiled
    N, plan = jb.flatten_plan([(2, 5, 30), (2, 30)], [1, None], [1, 1])
    assert plan[1] == ("broadcast", 5, (10, 30))
    with pytest.raises(ValueError):
        jb.flatten_plan([(1, 5, 74), (1, 6, 73)], [1, 1], [1, 1])   # inconsistent mapped sizes
    with pytest.raises(ValueError):
        jb.flatten_plan([(5, 74)], [0], [1])                        # no env axis left under the mapped one

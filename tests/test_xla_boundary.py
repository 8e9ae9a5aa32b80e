"""The XLA custom-call entry points of the hot path: `vnl_xla_step` / `vnl_xla_reset` driven exactly the way
XLA's status-returning legacy ABI drives them -- `(stream, buffers, opaque, opaque_len, status)` -- with

  * UNINITIALISED result buffers (XLA never zero-fills results),
  * the model / task operands RELOCATED between calls (XLA copies, donates and replicates operands: nothing may be keyed
    on their device address),
  * the workspace as the last buffer (a scratch result), one per calling thread,
  * two host threads calling concurrently on their own streams (what `pmap` does, one thread per device / replica),
  * failures reported through `XlaCustomCallStatusSetFailure` (a stand-in compiled by this test and loaded RTLD_GLOBAL, the
    way jaxlib's copy is visible to a custom call) instead of being dropped.
"""
import ctypes
import os
import subprocess
import threading

import numpy as np
import pytest

from conftest import pkg, start_states

pytestmark = pytest.mark.gpu

libm = pkg("_lib")
STATE_KEYS = libm.STATE_F + libm.STATE_I

FAKE_XLA = r"""
#include <stddef.h>
#include <string.h>
/* stand-in for xla/service/custom_call_status.h: the status object is opaque to the callee */
typedef struct { int failed; char msg[128]; } XlaCustomCallStatus;
void XlaCustomCallStatusSetFailure(XlaCustomCallStatus* s, const char* m, size_t n) {
  s->failed = 1; if (n > 127) n = 127; memcpy(s->msg, m, n); s->msg[n] = 0;
}
"""


class _Status(ctypes.Structure):
    _fields_ = [("failed", ctypes.c_int), ("msg", ctypes.c_char * 128)]


@pytest.fixture(scope="module")
def fake_xla(tmp_path_factory):
    d = tmp_path_factory.mktemp("fakexla")
    src, so = d / "fake_xla.c", d / "libfake_xla.so"
    src.write_text(FAKE_XLA)
    subprocess.check_call(["gcc", "-shared", "-fPIC", "-o", str(so), str(src)])
    return ctypes.CDLL(str(so), mode=ctypes.RTLD_GLOBAL)


def _buffers(eng, model_dev, task_dev, st_in, action, st_out, out, work):
    """The 32-entry buffer list of include/vnl_b200.h: operands, results, workspace last."""
    ptr = lambda t: None if t is None else t.data_ptr()
    bufs = [ptr(model_dev), ptr(task_dev)] + [ptr(st_in.get(k)) for k in STATE_KEYS] + [ptr(action)]
    bufs += [ptr(st_out[k]) for k in STATE_KEYS] + [ptr(out[k]) for k in ("obs", "traj", "reward", "done", "metrics", "stats")]
    bufs += [ptr(work)]
    assert len(bufs) == libm.VNL_XLA_STEP_NBUF
    return (ctypes.c_void_p * len(bufs))(*bufs)


def _garbage_like(d):
    """Result buffers as XLA hands them out: allocated, never initialised (here: filled with NaN / junk bits)."""
    import torch
    out = {}
    for k, v in d.items():
        out[k] = torch.full_like(v, float("nan")) if v.dtype.is_floating_point else torch.full_like(v, -12345)
    return out


def test_xla_step_and_reset_match_the_direct_calls(gpu_env, rodent, fake_xla):
    import torch
    eng = gpu_env.engine
    B = 300
    qpos, qvel, start = start_states(rodent, B, seed=71)
    s0 = gpu_env.reset_from(qpos, qvel, start)
    a = torch.tensor(np.random.default_rng(72).uniform(-1, 1, size=(B, 30)).astype(np.float32), device="cuda")
    st_in = dict(s0.pipeline_state, cur_frame=s0.info["cur_frame"], sub_clip_frame=s0.info["sub_clip_frame"], clip_id=s0.info["clip_idx"])
    ref_st, ref_out = eng.alloc_state(B), eng.alloc_outputs(B)
    eng.step(st_in, a, ref_st, ref_out)
    torch.cuda.synchronize()
    opaque = eng.xla_opaque(B)
    stream = torch.cuda.current_stream().cuda_stream
    for trial in range(3):
        # the blobs move to a fresh address on every call; the workspace is a fresh uninitialised scratch buffer
        model2, task2 = eng.model_dev.clone(), eng.task_dev.clone()
        work = torch.full_like(eng.workspace, float("nan"))
        st, out = _garbage_like(eng.alloc_state(B)), _garbage_like(eng.alloc_outputs(B))
        status = _Status()
        eng.lib.vnl_xla_step(stream, _buffers(eng, model2, task2, st_in, a, st, out, work), opaque, len(opaque), ctypes.byref(status))
        torch.cuda.synchronize()
        assert status.failed == 0, status.msg
        for k in STATE_KEYS:
            assert torch.equal(st[k], ref_st[k]), (trial, k)
        for k in ("obs", "traj", "reward", "done", "metrics", "stats"):
            assert torch.equal(out[k], ref_out[k]), (trial, k)
    # reset: same buffer list, `action` ignored
    rin = dict(qpos=torch.tensor(qpos, device="cuda"), qvel=torch.tensor(qvel, device="cuda"), cur_frame=torch.tensor(start, device="cuda"))
    st, out = _garbage_like(eng.alloc_state(B)), _garbage_like(eng.alloc_outputs(B))
    status = _Status()
    eng.lib.vnl_xla_reset(stream, _buffers(eng, eng.model_dev.clone(), eng.task_dev.clone(), rin, None, st, out,
                                           torch.empty_like(eng.workspace)), opaque, len(opaque), ctypes.byref(status))
    torch.cuda.synchronize()
    assert status.failed == 0
    for k in libm.STATE_F + ("cur_frame", "sub_clip_frame"):
        assert torch.equal(st[k], st_in[k]), k
    assert torch.equal(out["obs"], s0.obs) and torch.equal(out["traj"], s0.info["traj"])


def test_xla_failure_is_reported_not_dropped(gpu_env, rodent, fake_xla):
    """Round 1 dropped the return code (a lookup miss became a silent no-op on uninitialised results).  Now: a call that
    cannot launch sets the XLA status with the code; results are untouched."""
    import torch
    eng = gpu_env.engine
    B = 64
    st_in, a = eng.alloc_state(B), torch.zeros(B, 30, device="cuda")
    st, out = _garbage_like(eng.alloc_state(B)), _garbage_like(eng.alloc_outputs(B))
    stream = torch.cuda.current_stream().cuda_stream
    # (1) workspace declared too small for this batch
    op = libm.VnlXlaOpaque.from_buffer_copy(eng.xla_opaque(B))
    op.workspace_bytes = 1024
    status = _Status()
    eng.lib.vnl_xla_step(stream, _buffers(eng, eng.model_dev, eng.task_dev, st_in, a, st, out, eng.workspace), bytes(op), ctypes.sizeof(op),
                         ctypes.byref(status))
    assert status.failed == 1 and b"-21" in status.msg and b"vnl_xla_step" in status.msg
    # (2) truncated opaque
    status = _Status()
    eng.lib.vnl_xla_reset(stream, _buffers(eng, eng.model_dev, eng.task_dev, st_in, a, st, out, eng.workspace), bytes(op)[:64], 64,
                          ctypes.byref(status))
    assert status.failed == 1 and b"-30" in status.msg
    # (3) opaque built for another blob version
    op2 = libm.VnlXlaOpaque.from_buffer_copy(eng.xla_opaque(B))
    op2.model_hdr[1] += 1
    assert eng.lib.vnl_xla_step_rc(stream, _buffers(eng, eng.model_dev, eng.task_dev, st_in, a, st, out, eng.workspace), bytes(op2),
                                   ctypes.sizeof(op2)) == -11
    torch.cuda.synchronize()
    assert torch.isnan(out["reward"]).all()  # nothing ran


def test_two_host_threads_two_streams(gpu_env, rodent, fake_xla):
    """Re-entrancy (SURVEY 8b threading): two host threads, each with its own stream, workspace and relocated blobs, step
    different batches concurrently through the trampoline; both match the single-threaded results.  The library holds no
    lock and no global, so there is nothing to contend on."""
    import torch
    eng = gpu_env.engine
    jobs = []
    for i, B in enumerate((777, 1200)):
        qpos, qvel, start = start_states(rodent, B, seed=80 + i)
        s0 = gpu_env.reset_from(qpos, qvel, start)
        a = torch.tensor(np.random.default_rng(90 + i).uniform(-1, 1, size=(4, B, 30)).astype(np.float32), device="cuda")
        st_in = dict(s0.pipeline_state, cur_frame=s0.info["cur_frame"], sub_clip_frame=s0.info["sub_clip_frame"], clip_id=s0.info["clip_idx"])
        # single-threaded expectation: 4 steps in a row
        cur, outs = st_in, None
        for t in range(4):
            nxt, outs = eng.alloc_state(B), eng.alloc_outputs(B)
            eng.step(cur, a[t], nxt, outs)
            cur = nxt
        torch.cuda.synchronize()
        jobs.append(dict(B=B, st_in=st_in, a=a, want=cur, want_out=outs))
    results = [None, None]

    def worker(i):
        j = jobs[i]
        B = j["B"]
        stream = torch.cuda.Stream()
        ctx, work = eng.context_for_stream()
        opaque = eng.xla_opaque(B, ctx)
        with torch.cuda.stream(stream):
            cur, out = j["st_in"], None
            for t in range(4):
                model2, task2 = eng.model_dev.clone(), eng.task_dev.clone()
                nxt, out = _garbage_like(eng.alloc_state(B)), _garbage_like(eng.alloc_outputs(B))
                status = _Status()
                eng.lib.vnl_xla_step(stream.cuda_stream, _buffers(eng, model2, task2, cur, j["a"][t], nxt, out, work), opaque, len(opaque),
                                     ctypes.byref(status))
                assert status.failed == 0, status.msg
                cur = nxt
            stream.synchronize()
        results[i] = (cur, out)

    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for i in range(2):
        got, got_out = results[i]
        for k in STATE_KEYS:
            assert torch.equal(got[k], jobs[i]["want"][k]), (i, k)
        assert torch.equal(got_out["reward"], jobs[i]["want_out"]["reward"]) and torch.equal(got_out["obs"], jobs[i]["want_out"]["obs"])


def test_xla_call_is_graph_capturable(gpu_env, rodent, fake_xla):
    """XLA command buffers capture custom calls into CUDA graphs: the trampoline only enqueues."""
    import torch
    eng = gpu_env.engine
    B = 128
    qpos, qvel, start = start_states(rodent, B, seed=75)
    s0 = gpu_env.reset_from(qpos, qvel, start)
    a = torch.zeros(B, 30, device="cuda")
    st_in = dict(s0.pipeline_state, cur_frame=s0.info["cur_frame"], sub_clip_frame=s0.info["sub_clip_frame"], clip_id=s0.info["clip_idx"])
    ref_st, ref_out = eng.alloc_state(B), eng.alloc_outputs(B)
    eng.step(st_in, a, ref_st, ref_out)
    st, out = _garbage_like(eng.alloc_state(B)), _garbage_like(eng.alloc_outputs(B))
    opaque = eng.xla_opaque(B)
    bufs = _buffers(eng, eng.model_dev, eng.task_dev, st_in, a, st, out, eng.workspace)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        status = _Status()
        eng.lib.vnl_xla_step(torch.cuda.current_stream().cuda_stream, bufs, opaque, len(opaque), ctypes.byref(status))
        assert status.failed == 0
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["reward"], ref_out["reward"]) and torch.equal(st["qpos"], ref_st["qpos"])

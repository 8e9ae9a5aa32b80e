"""The per-env shared-memory layout recycles storage between the phases of a substep (csrc/vnl_kernels.cu make_layout).
This test states WHEN every array is live (first writer .. last reader, phases P1 .. P10 of the substep as numbered in
make_layout's comment) and checks, for every packaged model and both inertia homes, that two arrays that are live in the
same phase never overlap, that everything stays inside the slice, and that the slice still buys the residency the
benchmarks rely on.  compute-sanitizer is not available on the GPU pool: this is the hazard check of the overlays."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT, pkg

libm, mb, mj = pkg("_lib"), pkg("model_blob"), pkg("mjcf")

# phases: 1 kinematics, 2 com / cdof, 3 velocity + rne passes, 4 smooth forces, 5 inertia build (+ spill), 6 factorise +
# invert (+ spill), 7 smooth solve, 8 constraint rows, 9 solver, 10 integrator (incl. its factorisation and solve).
# "warm" etc. cross substeps: (1, 10).  A phase range is inclusive.
ALWAYS = (1, 10)
LIVE = {
    "qpos": ALWAYS, "qvel": ALWAYS, "act": ALWAYS, "ctrl": ALWAYS, "warm": ALWAYS, "ints": ALWAYS, "rcom": ALWAYS,
    "cdof": (2, 9), "Mdiag": (5, 10), "Kdiag": (6, 9), "qfrc_smooth": (4, 10), "qacc_smooth": (7, 9), "act_dot": (4, 10),
    "Ms": (5, 10), "Ks": (6, 9), "Mn": (5, 9), "H": (9, 9), "jr": (9, 9),
    "xpos": (1, 8), "xquat": (1, 8), "cvel": (3, 8),
    "Jaref": (9, 9), "qacc": (9, 10), "Ma": (9, 9), "grad": (9, 10), "Mgrad": (9, 10), "search": (9, 9), "Mv": (9, 9), "qfrc_con": (9, 10),
    "xipos": (2, 2), "xanchor": (1, 2), "xaxis": (1, 2), "cacc": (3, 5), "t16": (2, 5),
    "part": (7, 9), "tmpv": (7, 10),
    "lim_dof": (8, 9), "limrow_of_dof": (8, 9), "cbody": (8, 9), "crel": (8, 9), "cframe": (8, 9), "cmu": (8, 9), "efcD": (8, 9), "Jv": (8, 9),
    "K": (5, 6),       # factorisation workspace: M build .. inversion; again inside phase 10 (checked separately below)
}
# inside phase 10 only these are touched next to the (restored) factorisation workspace
PHASE10 = ("K", "tmpv", "grad", "Mgrad", "qacc", "qfrc_con", "qfrc_smooth", "act_dot", "Mdiag", "Ms", "qpos", "qvel", "act", "ctrl", "warm", "ints", "rcom")
# M build (phase 5) writes K while it reads fd = cacc and cdof; t16 is dead once fd is computed, so K may sit on it
K_MAY_SHARE_IN_5 = {"t16"}


def _layout(lib, blob):
    names = (ctypes.c_char_p * 64)()
    offs = (ctypes.c_int32 * 64)()
    sizes = (ctypes.c_int32 * 64)()
    n = lib.vnl_debug_layout(blob.ctypes.data, names, offs, sizes, 64)
    ent = {names[i].decode(): (offs[i], sizes[i]) for i in range(n)}
    total = ent.pop("total")[0]
    return ent, total


def _overlap(a, b):
    return a[1] > 0 and b[1] > 0 and a[0] < b[0] + b[1] and b[0] < a[0] + a[1]


@pytest.mark.parametrize("name", ["rodent", "humanoid", "ant", "rodent_pair"])
@pytest.mark.parametrize("stream", ["0", "1"])
def test_live_arrays_never_share_storage(name, stream, monkeypatch):
    monkeypatch.setenv("VNL_STREAM", stream)
    # VNL_STREAM is read once per process by the library: load a private copy so that both homes can be checked
    import shutil, tempfile
    tmp = tempfile.NamedTemporaryFile(suffix="_%s.so" % stream, delete=False).name
    shutil.copy(libm.LIB_PATH, tmp)
    lib = ctypes.CDLL(tmp)
    lib.vnl_debug_layout.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_int32),
                                     ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
    lib.vnl_envs_per_cta.argtypes = [ctypes.c_void_p]
    model = mj.load_model(os.path.join(ROOT, "vnl-brax-imitation_b200", "data", name + "_model.npz"))
    blob = mb.build_model_blob(model)
    ent, total = _layout(lib, blob)
    assert set(ent) == set(LIVE), set(ent) ^ set(LIVE)
    for k, (o, n) in ent.items():
        assert o >= 0 and o + n <= total, (k, o, n, total)
        assert o % 4 == 0, k  # 16-byte alignment of every array
    keys = sorted(ent)
    for i, a in enumerate(keys):
        for b in keys[i + 1:]:
            la, lb = LIVE[a], LIVE[b]
            both = range(max(la[0], lb[0]), min(la[1], lb[1]) + 1)
            if not len(both) or not _overlap(ent[a], ent[b]):
                continue
            if "K" in (a, b) and set(both) == {5} and ({a, b} - {"K"}) <= K_MAY_SHARE_IN_5:
                continue
            raise AssertionError("%s %s and %s %s share storage while both live in phases %s" % (a, ent[a], b, ent[b], list(both)))
    for a in PHASE10:  # the integrator's factorisation workspace against everything else it touches
        if a != "K":
            assert not _overlap(ent["K"], ent[a]), (a, ent["K"], ent[a])
    os.unlink(tmp)
    if name == "rodent" and stream == "1":
        assert lib.vnl_envs_per_cta(blob.ctypes.data) == 14 and total * 4 <= 14.6 * 1024 + 256

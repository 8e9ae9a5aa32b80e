"""Serialise a compiled `mjcf.Model` / imitation task into the flat blobs of `include/vnl_b200.h`.

The enum values (header slots, field ids) are parsed out of the C header at import, so the
Python writer and the C / CUDA readers cannot drift apart.
"""
from __future__ import annotations

import os
import re
from typing import Dict, List

import numpy as np

from . import mjcf

_HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "vnl_b200.h")


def _parse_header(path: str):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    consts: Dict[str, int] = {}
    for mm in re.finditer(r"#define\s+(VNL_\w+)\s+(0x[0-9a-fA-F]+u?|\d+)\s*$", txt, flags=re.M):
        consts[mm.group(1)] = int(mm.group(2).rstrip("u"), 0)
    for mm in re.finditer(r"enum\s+\w+\s*\{(.*?)\}", txt, flags=re.S):
        val = -1
        for item in mm.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                nm, v = [s.strip() for s in item.split("=")]
                val = int(v, 0)
            else:
                nm, val = item, val + 1
            consts[nm] = val
    consts["VNL_DATA_OFF"] = consts["VNL_TABLE_OFF"] + 2 * consts["VNL_MAX_FIELDS"]
    return consts


C = _parse_header(_HEADER)
DEFAULT_ENV_WARPS = 1


class _BlobWriter:
    def __init__(self, magic: int, nfields: int):
        self.hdr = np.zeros(C["VNL_DATA_OFF"], dtype=np.uint32)
        self.hdr[0] = magic
        self.hdr[1] = C["VNL_BLOB_VERSION"]
        self.hdr[3] = nfields
        self.chunks: List[np.ndarray] = []
        self.off = C["VNL_DATA_OFF"]

    def set_i(self, slot: str, v: int):
        self.hdr[C[slot]] = np.uint32(np.int32(v).view(np.uint32))

    def set_f(self, slot: str, v: float):
        self.hdr[C[slot]] = np.float32(v).view(np.uint32)

    def add(self, field: str, arr: np.ndarray, dtype):
        a = np.ascontiguousarray(np.asarray(arr).astype(dtype)).reshape(-1)
        f = C[field]
        self.hdr[C["VNL_TABLE_OFF"] + 2 * f] = self.off
        self.hdr[C["VNL_TABLE_OFF"] + 2 * f + 1] = a.size
        words = a.view(np.uint32)
        pad = (-words.size) % 4
        if pad:
            words = np.concatenate([words, np.zeros(pad, dtype=np.uint32)])
        self.chunks.append(words)
        self.off += words.size

    def finish(self) -> np.ndarray:
        self.hdr[2] = self.off
        return np.concatenate([self.hdr] + self.chunks)


def default_env_warps() -> int:
    """Warps cooperating on one env in the kernel (1 or 2); `VNL_ENV_WARPS` overrides the default."""
    ev = int(os.environ.get("VNL_ENV_WARPS", "0") or 0)
    return ev if ev in (1, 2) else DEFAULT_ENV_WARPS


def choose_env_warps(m: "mjcf.Model") -> int:
    """One warp per env unless shared memory leaves so few envs on an SM (<= 8: the two-warp kernel's register budget) that a second
    warp per env costs no residency -- then its extra latency hiding is free: rodent_pair (nv 146, 6 envs per SM) runs 31 % faster with
    two (0.346 -> 0.453 M env-steps/s at 16384 envs); rodent / humanoid / ant (14 / 16 / 16 envs per SM) keep one.  `VNL_ENV_WARPS`
    overrides.  The residency is asked of the library (`vnl_envs_per_cta`, host-side arithmetic on the blob header: no GPU needed)."""
    ev = int(os.environ.get("VNL_ENV_WARPS", "0") or 0)
    if ev in (1, 2):
        return ev
    import ctypes
    from . import _lib
    lib = _lib.load_library()
    lib.vnl_envs_per_cta.argtypes = [ctypes.c_void_p]
    lib.vnl_envs_per_cta.restype = ctypes.c_int
    blob = build_model_blob(m, 1)
    return 2 if 0 < lib.vnl_envs_per_cta(blob.ctypes.data) <= 8 else 1


def derived_tables(m: mjcf.Model, env_warps: int = 0) -> Dict[str, np.ndarray]:
    """Host-precomputed index tables for the CUDA kernels (tree levels, sparse-inertia
    pattern, emitted-contact list, static geom frames)."""
    A = m.arrays
    nb, nv = m.nbody, m.nv
    parent = A["body_parentid"]
    depth = np.zeros(nb, dtype=np.int64)
    for b in range(1, nb):
        depth[b] = 0 if parent[b] == 0 else depth[parent[b]] + 1
    order = sorted(range(1, nb), key=lambda b: (depth[b], b))
    nlevel = int(depth[1:].max()) + 1 if nb > 1 else 0
    level_start = [0]
    for lv in range(nlevel):
        level_start.append(level_start[-1] + int(np.sum(depth[1:] == lv)))
    madr, mcol, ddepth = [], [], []
    for i in range(nv):
        madr.append(len(mcol))
        j, d = i, 0
        while j >= 0:
            mcol.append(j)
            j = A["dof_parentid"][j]
            d += 1
        ddepth.append(d - 1)
    madr.append(len(mcol))
    sub_end = np.arange(nb) + 1
    for b in range(nb - 1, 0, -1):
        sub_end[parent[b]] = max(sub_end[parent[b]], sub_end[b])
    mrow = []
    for i in range(nv):
        mrow += [i] * (madr[i + 1] - madr[i])
    lastdof = np.full(nb, -1, dtype=np.int64)
    for b in range(1, nb):
        if A["body_dofnum"][b] > 0:
            lastdof[b] = A["body_dofadr"][b] + A["body_dofnum"][b] - 1
        else:
            lastdof[b] = lastdof[parent[b]]
    desc = [[] for _ in range(nv)]
    for e, (i, j) in enumerate(zip(mrow, mcol)):
        if i != j:
            desc[j].append(e)
    desc_adr = [0]
    for j in range(nv):
        desc_adr.append(desc_adr[-1] + len(desc[j]))
    desc_entry = [e for j in range(nv) for e in desc[j]]
    maxd = int(max(ddepth)) if ddepth else 0
    dl_start, dl_dof = [0], []
    for dpt in range(maxd + 1):
        dl_dof += [i for i in range(nv) if ddepth[i] == dpt]
        dl_start.append(len(dl_dof))
    limit_jnt = [j for j in range(m.njnt) if A["jnt_limited"][j] and A["jnt_type"][j] == mjcf.JNT_HINGE]
    con_pair, con_sign = [], []
    for p, t in enumerate(A["pair_type"]):
        if t == mjcf.GEOM_CAPSULE:
            con_pair += [p, p]
            con_sign += [1.0, -1.0]
        else:
            con_pair.append(p)
            con_sign.append(0.0)
    g2 = A["pair_geom2"]
    g1 = A["pair_geom1"]
    if len(g1) and np.any(A["geom_bodyid"][g1] != 0):
        raise NotImplementedError("planes must be attached to the world body")
    plane_mat = np.array([mjcf.quat_to_mat(A["geom_quat"][g]).reshape(9) for g in g1]).reshape(-1, 9)
    geom_mat = np.array([mjcf.quat_to_mat(A["geom_quat"][g]).reshape(9) for g in g2]).reshape(-1, 9)
    # ---- shared-memory index tables of the warp-per-env kernel (VNL_F_KTAB) -------------------------------------
    if nb > 255 or nv > 255 or len(mcol) > 65535:
        raise NotImplementedError("model too large for the packed kernel tables")
    children = [[] for _ in range(nb)]
    for b in range(nb - 1, 0, -1):
        children[parent[b]].append(b)  # descending id: the order `crb[parent] += crb[b]` visits them
    child_adr = [0]
    for b in range(nb):
        child_adr.append(child_adr[-1] + len(children[b]))
    roots = [b for b in range(1, nb) if parent[b] == 0]
    tree = np.zeros(nb, dtype=np.int64)
    for b in range(1, nb):
        tree[b] = roots.index(int(A["body_rootid"][b]))
    dofnum_flag = np.array(A["body_dofnum"], dtype=np.int64).copy()
    for b in range(1, nb):
        if A["body_jntnum"][b] > 0 and A["jnt_type"][A["body_jntadr"][b]] == mjcf.JNT_FREE:
            dofnum_flag[b] |= 0x80
    # lane programs of the sparse mat-vecs (see VnlKtab in include/vnl_b200.h)
    nM_ = len(mcol)
    if 4 * (nM_ + 1) >= (1 << 14):
        raise NotImplementedError("too many inertia entries for the packed mat-vec program")

    env_warps = env_warps or default_env_warps()
    NL = 32 * env_warps  # lanes of the mat-vec programs = threads cooperating on one env

    def pack_program(rows):
        """rows: list of (slot, [(entry, xindex), ...]) with non-empty term lists -> ([T*NL] words, T)."""
        lanes = [[] for _ in range(NL)]
        lane_load = [0] * NL
        for slot, terms in sorted(rows, key=lambda r: -len(r[1])):
            l = min(range(NL), key=lambda l: lane_load[l])
            lanes[l].append((slot, terms))
            lane_load[l] += len(terms)
        T = 8 * ((max(1, max(lane_load)) + 7) // 8)  # the kernel walks the program eight terms at a time
        prog = np.zeros((T, NL), dtype=np.uint32)
        pad = np.uint32((4 * nM_) | (0xFF << 24))  # entry nM is a zero slot, never flushed
        prog[:, :] = pad
        for l in range(NL):
            t = 0
            for slot, terms in lanes[l]:
                for k, (e, xi) in enumerate(terms):
                    flush = slot if k == len(terms) - 1 else 0xFF
                    prog[t, l] = np.uint32((4 * int(e)) | ((4 * int(xi)) << 14) | (int(flush) << 24))
                    t += 1
        # [T / 4][lanes][4]: the four consecutive steps of a lane form one 16-byte word
        return np.ascontiguousarray(prog.reshape(T // 4, 4, NL).transpose(0, 2, 1)).reshape(-1), T

    # long rows / columns are cut into chunks so that the lanes can be balanced; one warp per env keeps whole rows
    CH = 24 if env_warps == 1 else 12
    CHA = 1 << 20 if env_warps == 1 else 12
    rows_a, apart_adr = [], [0]
    for i in range(nv):
        terms = [(madr[i] + a, mcol[madr[i] + a]) for a in range(1, madr[i + 1] - madr[i])]
        for k in range(0, len(terms), CHA):
            rows_a.append((len(rows_a), terms[k:k + CHA]))
        apart_adr.append(len(rows_a))
    if len(rows_a) > 254:
        raise NotImplementedError("too many partial-sum slots for the packed mat-vec program")
    prog_a, TA = pack_program(rows_a) if rows_a else (np.full(8 * NL, np.uint32((4 * nM_) | (0xFF << 24)), dtype=np.uint32), 8)
    naslot = len(rows_a)
    rows_d, dpart_adr = [], [0]
    for j in range(nv):
        terms = [(e, mrow[e]) for e in desc[j]]
        for k in range(0, len(terms), CH):
            rows_d.append((len(rows_d), terms[k:k + CH]))
        dpart_adr.append(len(rows_d))
    if len(rows_d) > 254 or nv > 254:
        raise NotImplementedError("too many partial-sum slots for the packed mat-vec program")
    prog_d, TD = pack_program(rows_d) if rows_d else (np.full(8 * NL, np.uint32((4 * nM_) | (0xFF << 24)), dtype=np.uint32), 8)
    ndslot = len(rows_d)
    # Level schedules of the factorisation and of the triangular solves (see VnlKtab): rows by dof HEIGHT (leaves = 0),
    # each row with the list of its descendant dofs; dofs by DEPTH for the gather-from-ancestors sweep.
    lg2c = lambda n: 0 if n <= 1 else min(5, int(np.ceil(np.log2(n))))
    dpar = [int(mcol[madr[i] + 1]) if ddepth[i] > 0 else -1 for i in range(nv)]
    desc_of = [[] for _ in range(nv)]
    height = [0] * nv
    for k in range(nv - 1, -1, -1):
        j, a = dpar[k], 1
        while j >= 0:
            desc_of[j].append((k, a))  # k descending inside each list: the order a right-looking elimination visits
            j, a = dpar[j], a + 1
        if dpar[k] >= 0:
            height[dpar[k]] = max(height[dpar[k]], height[k] + 1)
    maxh = max(height) if nv else 0
    erow, elvl, desc_adr, desc_src, desc_k = [], [], [0], [], []
    for j in range(nv):
        for k, a in desc_of[j]:
            desc_src.append(madr[k] + a)
            desc_k.append(k)
        # the elimination walks a row's list four entries per lane group at a time: pad it to a multiple of 4 * groups with
        # the zero slot behind the inertia entries (entry nM, and the 39 zeros the kernel keeps behind it), descendant 0
        groups = 32 >> lg2c(ddepth[j] + 1) if ddepth[j] + 1 <= 32 else 1
        while desc_of[j] and (len(desc_src) - desc_adr[-1]) % (4 * groups):
            desc_src.append(nM_)
            desc_k.append(0)
        desc_adr.append(len(desc_src))
    if nM_ >= (1 << 13) or len(desc_src) >= (1 << 16):
        raise NotImplementedError("model too large for the packed factorisation schedule")
    for h in range(0, maxh + 1):
        rows = sorted([j for j in range(nv) if height[j] == h], reverse=True)
        elvl.append(len(erow) // 2 | (lg2c(max(len(desc_of[j]) for j in rows)) << 8))
        for j in rows:  # two words per row: entry base | (row length - 1) << 13 | lg2ceil(row length) << 19 | dof << 22; desc range
            erow += [madr[j] | (int(ddepth[j]) << 13) | (lg2c(ddepth[j] + 1) << 19) | (j << 22), desc_adr[j] | (desc_adr[j + 1] << 16)]
    elvl.append(len(erow) // 2)
    ddof, dlvl = [], []
    for dpt in range(1, maxd + 1):
        dlvl.append(len(ddof) | (lg2c(dpt) << 8))
        ddof += [i for i in range(nv) if ddepth[i] == dpt]
    dlvl.append(len(ddof))
    kitem, klvl = [], [0, 0]
    for dpt in range(1, maxd + 1):
        items = [(c, i) for i in range(nv) if ddepth[i] == dpt for c in range(1, dpt + 1)]
        items.sort(key=lambda t: (-t[0], t[1]))
        kitem += [c | (i << 8) for c, i in items]
        klvl.append(len(kitem))
    u8 = lambda a: np.asarray(a, dtype=np.int64).astype(np.uint8)
    u16 = lambda a: np.asarray(a, dtype=np.int64).astype(np.uint16)
    lastdof_u8 = np.where(lastdof < 0, 0xFF, lastdof)
    kt = [None] * C["VNL_KT_COUNT"]
    kt[C["VNL_KT_LVL_START"]] = u8(level_start)
    kt[C["VNL_KT_LVL_BP"]] = u16([b | (int(parent[b]) << 8) for b in order])
    kt[C["VNL_KT_PARENT"]] = u8(parent)
    kt[C["VNL_KT_CHILD_ADR"]] = u8(child_adr)
    kt[C["VNL_KT_CHILD_LIST"]] = u8([ch for b in range(nb) for ch in children[b]])
    kt[C["VNL_KT_BODY_DOFADR"]] = u8(np.maximum(A["body_dofadr"], 0))
    kt[C["VNL_KT_BODY_DOFNUM"]] = u8(dofnum_flag)
    kt[C["VNL_KT_BODY_TREE"]] = u8(tree)
    kt[C["VNL_KT_BODY_LASTDOF"]] = u8(lastdof_u8)
    kt[C["VNL_KT_SUB_END"]] = u8(sub_end)
    kt[C["VNL_KT_ROOTS"]] = u8(roots)
    kt[C["VNL_KT_MROW"]] = u8(mrow)
    kt[C["VNL_KT_MCOL"]] = u8(mcol)
    kt[C["VNL_KT_DOF_BODY"]] = u8(A["dof_bodyid"])
    kt[C["VNL_KT_DPART_ADR"]] = u8(dpart_adr)
    kt[C["VNL_KT_APART_ADR"]] = u8(apart_adr)
    kt[C["VNL_KT_MADR"]] = u16(madr)
    kt[C["VNL_KT_EROW"]] = np.asarray(erow, dtype=np.int64).astype(np.uint32)
    kt[C["VNL_KT_ELVL"]] = u16(elvl)
    kt[C["VNL_KT_DESC_ADR"]] = u16(desc_adr)
    kt[C["VNL_KT_DESC_SRC"]] = u16(desc_src)
    kt[C["VNL_KT_DESC_K"]] = u8(desc_k)
    kt[C["VNL_KT_DDOF"]] = u8(ddof)
    kt[C["VNL_KT_DLVL"]] = u16(dlvl)
    kt[C["VNL_KT_ANC_START"]] = u16([madr[j] for j in mcol])
    kt[C["VNL_KT_KITEM"]] = u16(kitem)
    kt[C["VNL_KT_KLVL"]] = u16(klvl)
    kt[C["VNL_KT_PROG_A"]] = prog_a
    kt[C["VNL_KT_PROG_D"]] = prog_d
    nkt, nks = C["VNL_KT_COUNT"], C["VNL_KT_NSCALAR"]
    ndir = (nkt + nks + 3) // 4 * 4  # directory padded to a multiple of four words: tables start 16-byte aligned
    off = 4 * ndir
    dirw = np.zeros(ndir, dtype=np.uint32)
    blobs = []
    for t, a in enumerate(kt):
        raw = a.tobytes()
        raw += b"\0" * ((-len(raw)) % 16)  # every table 16-byte aligned (the lane programs are read as 128-bit words)
        dirw[t] = off
        off += len(raw)
        blobs.append(raw)
    # the kernel streams PROG_A and PROG_D as one program
    assert dirw[C["VNL_KT_PROG_D"]] == dirw[C["VNL_KT_PROG_A"]] + 4 * TA * NL
    dirw[nkt + C["VNL_KS_TA"]] = TA
    dirw[nkt + C["VNL_KS_TD"]] = TD
    dirw[nkt + C["VNL_KS_NDSLOT"]] = ndslot
    dirw[nkt + C["VNL_KS_NHEIGHT"]] = maxh
    ktab = np.frombuffer(dirw.tobytes() + b"".join(blobs), dtype=np.uint32).copy()
    act_of_dof = [[] for _ in range(nv)]
    for u, dadr in enumerate(A["actuator_dofadr"]):
        act_of_dof[int(dadr)].append(u)
    dof_actadr = [0]
    for i in range(nv):
        dof_actadr.append(dof_actadr[-1] + len(act_of_dof[i]))
    return dict(ktab=ktab, TA=TA, TD=TD, env_warps=env_warps, nroot=len(roots), ndslot=ndslot, naslot=naslot, dof_actadr=np.array(dof_actadr),
                dof_actlist=np.array([u for i in range(nv) for u in act_of_dof[i]], dtype=np.int64),
                level_start=np.array(level_start), level_body=np.array(order), dof_madr=np.array(madr),
                m_col=np.array(mcol), dof_depth=np.array(ddepth), body_subtree_end=sub_end,
                limit_jnt=np.array(limit_jnt, dtype=np.int64), con_pair=np.array(con_pair, dtype=np.int64),
                con_sign=np.array(con_sign), geomc_body=A["geom_bodyid"][g2], geomc_pos=A["geom_pos"][g2],
                geomc_mat=geom_mat, plane_pos=A["geom_pos"][g1], plane_mat=plane_mat,
                body_imat=np.array([mjcf.quat_to_mat(q).reshape(9) for q in A["body_iquat"]]),
                m_row=np.array(mrow), body_lastdof=lastdof, desc_adr=np.array(desc_adr),
                desc_entry=np.array(desc_entry, dtype=np.int64), doflevel_start=np.array(dl_start),
                doflevel_dof=np.array(dl_dof, dtype=np.int64),
                nlevel=nlevel, maxdepth=int(max(ddepth)) if ddepth else 0)


_MODEL_FIELDS = [
    ("VNL_F_BODY_PARENTID", "body_parentid", np.int32), ("VNL_F_BODY_ROOTID", "body_rootid", np.int32),
    ("VNL_F_BODY_JNTADR", "body_jntadr", np.int32), ("VNL_F_BODY_JNTNUM", "body_jntnum", np.int32),
    ("VNL_F_BODY_DOFADR", "body_dofadr", np.int32), ("VNL_F_BODY_DOFNUM", "body_dofnum", np.int32),
    ("VNL_F_BODY_POS", "body_pos", np.float32), ("VNL_F_BODY_QUAT", "body_quat", np.float32),
    ("VNL_F_BODY_IPOS", "body_ipos", np.float32), ("VNL_F_BODY_IQUAT", "body_iquat", np.float32),
    ("VNL_F_BODY_MASS", "body_mass", np.float32), ("VNL_F_BODY_INERTIA", "body_inertia", np.float32),
    ("VNL_F_BODY_INVWEIGHT0", "body_invweight0", np.float32),
    ("VNL_F_JNT_TYPE", "jnt_type", np.int32), ("VNL_F_JNT_QPOSADR", "jnt_qposadr", np.int32),
    ("VNL_F_JNT_DOFADR", "jnt_dofadr", np.int32), ("VNL_F_JNT_BODYID", "jnt_bodyid", np.int32),
    ("VNL_F_JNT_LIMITED", "jnt_limited", np.int32), ("VNL_F_JNT_POS", "jnt_pos", np.float32),
    ("VNL_F_JNT_AXIS", "jnt_axis", np.float32), ("VNL_F_JNT_STIFFNESS", "jnt_stiffness", np.float32),
    ("VNL_F_JNT_RANGE", "jnt_range", np.float32), ("VNL_F_JNT_MARGIN", "jnt_margin", np.float32),
    ("VNL_F_JNT_SOLREF", "jnt_solref", np.float32), ("VNL_F_JNT_SOLIMP", "jnt_solimp", np.float32),
    ("VNL_F_DOF_BODYID", "dof_bodyid", np.int32), ("VNL_F_DOF_JNTID", "dof_jntid", np.int32),
    ("VNL_F_DOF_PARENTID", "dof_parentid", np.int32), ("VNL_F_DOF_ARMATURE", "dof_armature", np.float32),
    ("VNL_F_DOF_DAMPING", "dof_damping", np.float32), ("VNL_F_DOF_INVWEIGHT0", "dof_invweight0", np.float32),
    ("VNL_F_QPOS0", "qpos0", np.float32), ("VNL_F_QPOS_SPRING", "qpos_spring", np.float32),
    ("VNL_F_GEOM_TYPE", "geom_type", np.int32), ("VNL_F_GEOM_BODYID", "geom_bodyid", np.int32),
    ("VNL_F_GEOM_POS", "geom_pos", np.float32), ("VNL_F_GEOM_QUAT", "geom_quat", np.float32),
    ("VNL_F_GEOM_SIZE", "geom_size", np.float32),
    ("VNL_F_ACT_DOFADR", "actuator_dofadr", np.int32), ("VNL_F_ACT_CTRLLIMITED", "actuator_ctrllimited", np.int32),
    ("VNL_F_ACT_FORCELIMITED", "actuator_forcelimited", np.int32), ("VNL_F_ACT_DYNTYPE", "actuator_dyntype", np.int32),
    ("VNL_F_ACT_ACTADR", "actuator_actadr", np.int32), ("VNL_F_ACT_GAIN", "actuator_gain", np.float32),
    ("VNL_F_ACT_GEAR", "actuator_gear", np.float32), ("VNL_F_ACT_CTRLRANGE", "actuator_ctrlrange", np.float32),
    ("VNL_F_ACT_FORCERANGE", "actuator_forcerange", np.float32), ("VNL_F_ACT_DYNPRM", "actuator_dynprm", np.float32),
    ("VNL_F_PAIR_GEOM1", "pair_geom1", np.int32), ("VNL_F_PAIR_GEOM2", "pair_geom2", np.int32),
    ("VNL_F_PAIR_TYPE", "pair_type", np.int32), ("VNL_F_PAIR_FRICTION", "pair_friction", np.float32),
    ("VNL_F_PAIR_SOLREF", "pair_solref", np.float32), ("VNL_F_PAIR_SOLIMP", "pair_solimp", np.float32),
    ("VNL_F_PAIR_INCLUDEMARGIN", "pair_includemargin", np.float32),
]
_DERIVED_FIELDS = [
    ("VNL_F_LEVEL_START", "level_start", np.int32), ("VNL_F_LEVEL_BODY", "level_body", np.int32),
    ("VNL_F_DOF_MADR", "dof_madr", np.int32), ("VNL_F_M_COL", "m_col", np.int32),
    ("VNL_F_DOF_DEPTH", "dof_depth", np.int32), ("VNL_F_BODY_SUBTREE_END", "body_subtree_end", np.int32),
    ("VNL_F_LIMIT_JNT", "limit_jnt", np.int32), ("VNL_F_CON_PAIR", "con_pair", np.int32),
    ("VNL_F_CON_SIGN", "con_sign", np.float32), ("VNL_F_GEOMC_BODY", "geomc_body", np.int32),
    ("VNL_F_GEOMC_POS", "geomc_pos", np.float32), ("VNL_F_GEOMC_MAT", "geomc_mat", np.float32),
    ("VNL_F_PLANE_POS", "plane_pos", np.float32), ("VNL_F_PLANE_MAT", "plane_mat", np.float32),
    ("VNL_F_BODY_IMAT", "body_imat", np.float32),
    ("VNL_F_M_ROW", "m_row", np.int32), ("VNL_F_BODY_LASTDOF", "body_lastdof", np.int32),
    ("VNL_F_DESC_ADR", "desc_adr", np.int32), ("VNL_F_DESC_ENTRY", "desc_entry", np.int32),
    ("VNL_F_DOFLEVEL_START", "doflevel_start", np.int32), ("VNL_F_DOFLEVEL_DOF", "doflevel_dof", np.int32),
    ("VNL_F_DOF_ACTADR", "dof_actadr", np.int32), ("VNL_F_DOF_ACTLIST", "dof_actlist", np.int32),
    ("VNL_F_KTAB", "ktab", np.uint32),
]


def model_dims(m: mjcf.Model, d=None) -> Dict[str, int]:
    d = d or derived_tables(m)
    ncon, nlimit = len(d["con_pair"]), len(d["limit_jnt"])
    return dict(nq=m.nq, nv=m.nv, nu=m.nu, na=m.na, nbody=m.nbody, njnt=m.njnt, ngeom=m.ngeom,
                npair=len(m.arrays["pair_geom1"]), ncon=ncon, nlimit=nlimit, nefc=nlimit + 4 * ncon,
                nM=len(d["m_col"]), nlevel=d["nlevel"], maxdepth=d["maxdepth"], nroot=d["nroot"], ndslot=d["ndslot"], naslot=d["naslot"], TA=d["TA"], TD=d["TD"])


def build_model_blob(m: mjcf.Model, env_warps: int = 0) -> np.ndarray:
    d = derived_tables(m, env_warps or choose_env_warps(m))
    dims = model_dims(m, d)
    w = _BlobWriter(C["VNL_MAGIC_MODEL"], C["VNL_F_MODEL_COUNT"])
    for slot, key in [("VNL_MH_NQ", "nq"), ("VNL_MH_NV", "nv"), ("VNL_MH_NU", "nu"), ("VNL_MH_NA", "na"),
                      ("VNL_MH_NBODY", "nbody"), ("VNL_MH_NJNT", "njnt"), ("VNL_MH_NGEOM", "ngeom"),
                      ("VNL_MH_NPAIR", "npair"), ("VNL_MH_NCON", "ncon"), ("VNL_MH_NLIMIT", "nlimit"),
                      ("VNL_MH_NEFC", "nefc"), ("VNL_MH_NM", "nM"), ("VNL_MH_NLEVEL", "nlevel"),
                      ("VNL_MH_MAXDEPTH", "maxdepth"), ("VNL_MH_NROOT", "nroot"), ("VNL_MH_NDSLOT", "ndslot"),
                      ("VNL_MH_NASLOT", "naslot"), ("VNL_MH_TA", "TA"), ("VNL_MH_TD", "TD")]:
        w.set_i(slot, dims[key])
    w.set_i("VNL_MH_ENV_WARPS", d["env_warps"])
    w.set_i("VNL_MH_SOLVER", m.solver)
    w.set_i("VNL_MH_ITERATIONS", m.iterations)
    w.set_i("VNL_MH_LS_ITERATIONS", m.ls_iterations)
    w.set_i("VNL_MH_EULERDAMP", int(m.eulerdamp))
    w.set_f("VNL_MH_TIMESTEP", m.timestep)
    w.set_f("VNL_MH_GRAVITY_X", m.gravity[0])
    w.set_f("VNL_MH_GRAVITY_Y", m.gravity[1])
    w.set_f("VNL_MH_GRAVITY_Z", m.gravity[2])
    w.set_f("VNL_MH_TOLERANCE", m.tolerance)
    w.set_f("VNL_MH_LS_TOLERANCE", m.ls_tolerance)
    w.set_f("VNL_MH_IMPRATIO", m.impratio)
    w.set_f("VNL_MH_MEANINERTIA", m.meaninertia)
    for fid, key, dt in _MODEL_FIELDS:
        w.add(fid, m.arrays[key], dt)
    for fid, key, dt in _DERIVED_FIELDS:
        w.add(fid, d[key], dt)
    return w.finish()


def read_dims(blob: np.ndarray) -> Dict[str, int]:
    g = lambda s: int(blob[C[s]].view(np.int32))
    return dict(nq=g("VNL_MH_NQ"), nv=g("VNL_MH_NV"), nu=g("VNL_MH_NU"), na=g("VNL_MH_NA"),
                nbody=g("VNL_MH_NBODY"), njnt=g("VNL_MH_NJNT"), ngeom=g("VNL_MH_NGEOM"), npair=g("VNL_MH_NPAIR"),
                ncon=g("VNL_MH_NCON"), nlimit=g("VNL_MH_NLIMIT"), nefc=g("VNL_MH_NEFC"), nM=g("VNL_MH_NM"),
                nlevel=g("VNL_MH_NLEVEL"), maxdepth=g("VNL_MH_MAXDEPTH"), nroot=g("VNL_MH_NROOT"), solver=g("VNL_MH_SOLVER"),
                iterations=g("VNL_MH_ITERATIONS"), ls_iterations=g("VNL_MH_LS_ITERATIONS"),
                eulerdamp=g("VNL_MH_EULERDAMP"), env_warps=g("VNL_MH_ENV_WARPS"))


def read_field(blob: np.ndarray, field: str, dtype) -> np.ndarray:
    f = C[field]
    off = int(blob[C["VNL_TABLE_OFF"] + 2 * f])
    n = int(blob[C["VNL_TABLE_OFF"] + 2 * f + 1])
    return blob[off:off + n].view(dtype)


def read_hdr_f(blob: np.ndarray, slot: str) -> float:
    return float(blob[C[slot]:C[slot] + 1].view(np.float32)[0])


def build_task_blob(clip, *, body_idxs, end_eff_idx, app_idx, joint_idxs, com_idx: int, njoint_cols: int,
                    clip_length: int = 250, ref_traj_length: int = 5, sub_clip_length: int = 10,
                    healthy_z_range=(0.05, 0.5), termination_threshold: float = 5.0,
                    body_error_multiplier: float = 1.0, n_frames: int = 5, torso_body: int = 1,
                    obs_size: int = 0, traj_size: int = 0, kind: int = 0, reward_old_state: bool = False,
                    term_mean: bool = False, use_subclip: bool = True, obs_qfrc: bool = True, com_from_field: bool = False,
                    done_rtrunk: float = 0.0, rot_body: int = -1, traj_old_frame: bool = False, ract_action: bool = False,
                    metrics_raw: bool = False, weights=(0.01, 0.01, 0.01, 0.01, 0.0001, 0.01)) -> np.ndarray:
    """Rodent imitation task tables.  `clip.body_positions` must already be filtered to
    `body_idxs` (`envs/rodent.py:114-115`).  The reference indexes that filtered table with
    MODEL body ids and relies on JAX clamping out-of-range gathers; the
    clamped indices are baked here so kernels only gather."""
    ntrack = len(body_idxs)
    # multi-clip tables: fields with a leading clip axis [nclips, T, ...] are stored clip-major, flattened
    # to [nclips * T, ...]; the kernel offsets every row by clip_id * T
    nclips = 1
    if np.asarray(clip.position).ndim == 3:
        nclips = int(np.asarray(clip.position).shape[0])
        flat = lambda a: None if a is None else np.asarray(a).reshape((-1,) + np.asarray(a).shape[2:])
        clip = clip.replace(**{k: flat(getattr(clip, k)) for k in ("position", "quaternion", "joints", "body_positions", "velocity",
                                                                  "angular_velocity", "joints_velocity", "body_quaternions",
                                                                  "center_of_mass") if getattr(clip, k, None) is not None})
    w = _BlobWriter(C["VNL_MAGIC_TASK"], C["VNL_TASK_COUNT"])
    w.set_i("VNL_TH_NCLIPS", nclips)
    w.set_i("VNL_TH_KIND", kind)
    w.set_i("VNL_TH_REWARD_OLD_STATE", int(reward_old_state))
    w.set_i("VNL_TH_TERM_MEAN", int(term_mean))
    w.set_i("VNL_TH_USE_SUBCLIP", int(use_subclip))
    w.set_i("VNL_TH_OBS_QFRC", int(obs_qfrc))
    w.set_i("VNL_TH_COM_FROM_FIELD", int(com_from_field))
    w.set_f("VNL_TH_DONE_RTRUNK", done_rtrunk)
    w.set_i("VNL_TH_ROT_BODY", torso_body if rot_body < 0 else rot_body)
    w.set_i("VNL_TH_TRAJ_OLD_FRAME", int(traj_old_frame))
    w.set_i("VNL_TH_RACT_ACTION", int(ract_action))
    w.set_i("VNL_TH_METRICS_RAW", int(metrics_raw))
    for slot, v in zip(("RCOM", "RVEL", "RTRUNK", "RQUAT", "RACT", "RAPP"), weights):
        w.set_f("VNL_TH_W_" + slot, v)
    w.set_i("VNL_TH_CLIP_LEN", clip.position.shape[0] // nclips)
    w.set_i("VNL_TH_REF_LEN", ref_traj_length)
    w.set_i("VNL_TH_SUB_CLIP_LEN", sub_clip_length)
    w.set_i("VNL_TH_NTRACK", ntrack)
    w.set_i("VNL_TH_NJIDX", len(joint_idxs))
    w.set_i("VNL_TH_NAPP", len(app_idx))
    w.set_i("VNL_TH_NEE", len(end_eff_idx))
    w.set_i("VNL_TH_NFRAMES", n_frames)
    w.set_i("VNL_TH_OBS_SIZE", obs_size)
    w.set_i("VNL_TH_TRAJ_SIZE", traj_size)
    w.set_i("VNL_TH_COM_REF_IDX", min(max(int(com_idx), 0), ntrack - 1))
    w.set_i("VNL_TH_TORSO_BODY", torso_body)
    w.set_f("VNL_TH_HEALTHY_LO", healthy_z_range[0])
    w.set_f("VNL_TH_HEALTHY_HI", healthy_z_range[1])
    w.set_f("VNL_TH_TERM_THRESHOLD", termination_threshold)
    w.set_f("VNL_TH_BODY_ERR_MULT", body_error_multiplier)
    assert clip.body_positions.shape[1] == ntrack
    w.add("VNL_T_POSITION", clip.position, np.float32)
    w.add("VNL_T_QUATERNION", clip.quaternion, np.float32)
    w.add("VNL_T_JOINTS", clip.joints, np.float32)
    w.add("VNL_T_BODY_POSITIONS", clip.body_positions, np.float32)
    w.add("VNL_T_VELOCITY", clip.velocity, np.float32)
    w.add("VNL_T_ANGULAR_VELOCITY", clip.angular_velocity, np.float32)
    w.add("VNL_T_JOINTS_VELOCITY", clip.joints_velocity, np.float32)
    w.add("VNL_T_BODY_IDXS", body_idxs, np.int32)
    w.add("VNL_T_EE_IDX", end_eff_idx, np.int32)
    w.add("VNL_T_APP_IDX", app_idx, np.int32)
    w.add("VNL_T_APP_REF_IDX", np.clip(np.asarray(app_idx, dtype=np.int64), 0, ntrack - 1), np.int32)
    w.add("VNL_T_JOINT_COL", np.clip(np.asarray(joint_idxs, dtype=np.int64), 0, njoint_cols - 1), np.int32)
    com = getattr(clip, "center_of_mass", None)
    w.add("VNL_T_CENTER_OF_MASS", com if com is not None else np.zeros((clip.position.shape[0], 3)), np.float32)
    return w.finish()

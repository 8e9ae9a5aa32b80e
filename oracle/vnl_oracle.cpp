/*
 * vnl_oracle.cpp -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED for dynamics: the arithmetic of the reference path lives in third-party
 * packages that are neither vendored in /root/reference nor installable in this image
 * (requirements.txt:3-5, unpinned: mujoco / mujoco-mjx ~3.1.4-3.1.5, brax ~0.10.3).  This file
 * restates the published MJX algorithm (`mjx.forward` / `mjx.step`, dense-Jacobian formulation
 * that `opt.jacobian = 0` selects, envs/rodent.py:63) as reached from the reference call sites
 *   envs/rodent.py:148  pipeline_init  -> mjx.forward
 *   envs/rodent.py:181  pipeline_step  -> n_frames x mjx.step
 * and the task logic of envs/rodent.py:178-470 line by line.  What IS pinned by reference
 * fixtures: forward kinematics, body COM and the clip-velocity pipeline
 * (clips/transform_snips_groom.p, tests/test_mjcf_clip.py, tests/test_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path never does.
 *
 * Templated on the scalar type: T = float mirrors the reference's fp32 (JAX default, x64 off),
 * T = double is the fp64 twin used to show that fp32 differences are rounding.
 * State crosses the C ABI as double in both cases (fp32 values round-trip exactly).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/vnl_blob.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double kMinVal = 1e-15;  // mjMINVAL
constexpr double kMinImp = 1e-4;   // mjMINIMP
constexpr double kMaxImp = 0.9999; // mjMAXIMP

struct ModelView {
  const uint32_t* w;
  int nq, nv, nu, na, nbody, njnt, ngeom, npair, ncon, nlimit, nefc, solver, iterations, ls_iterations, eulerdamp;
  float timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia;
  const int *body_parentid, *body_rootid, *body_jntadr, *body_jntnum, *jnt_type, *jnt_qposadr, *jnt_dofadr, *jnt_bodyid,
      *jnt_limited, *dof_bodyid, *dof_parentid, *geom_bodyid, *act_dofadr, *act_ctrllimited, *act_forcelimited,
      *act_dyntype, *act_actadr, *pair_geom1, *pair_geom2, *pair_type, *limit_jnt;
  const float *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia, *body_invweight0, *jnt_pos,
      *jnt_axis, *jnt_stiffness, *jnt_range, *jnt_margin, *jnt_solref, *jnt_solimp, *dof_armature, *dof_damping,
      *dof_invweight0, *qpos0, *qpos_spring, *geom_pos, *geom_quat, *geom_size, *act_gain, *act_gear, *act_ctrlrange,
      *act_forcerange, *act_dynprm, *pair_friction, *pair_solref, *pair_solimp, *pair_includemargin;
  explicit ModelView(const uint32_t* b) : w(b) {
    nq = vnl_hdr_i(b, VNL_MH_NQ); nv = vnl_hdr_i(b, VNL_MH_NV); nu = vnl_hdr_i(b, VNL_MH_NU); na = vnl_hdr_i(b, VNL_MH_NA);
    nbody = vnl_hdr_i(b, VNL_MH_NBODY); njnt = vnl_hdr_i(b, VNL_MH_NJNT); ngeom = vnl_hdr_i(b, VNL_MH_NGEOM);
    npair = vnl_hdr_i(b, VNL_MH_NPAIR); ncon = vnl_hdr_i(b, VNL_MH_NCON); nlimit = vnl_hdr_i(b, VNL_MH_NLIMIT);
    nefc = vnl_hdr_i(b, VNL_MH_NEFC); solver = vnl_hdr_i(b, VNL_MH_SOLVER); iterations = vnl_hdr_i(b, VNL_MH_ITERATIONS);
    ls_iterations = vnl_hdr_i(b, VNL_MH_LS_ITERATIONS); eulerdamp = vnl_hdr_i(b, VNL_MH_EULERDAMP);
    timestep = vnl_hdr_f(b, VNL_MH_TIMESTEP);
    gravity[0] = vnl_hdr_f(b, VNL_MH_GRAVITY_X); gravity[1] = vnl_hdr_f(b, VNL_MH_GRAVITY_Y); gravity[2] = vnl_hdr_f(b, VNL_MH_GRAVITY_Z);
    tolerance = vnl_hdr_f(b, VNL_MH_TOLERANCE); ls_tolerance = vnl_hdr_f(b, VNL_MH_LS_TOLERANCE);
    impratio = vnl_hdr_f(b, VNL_MH_IMPRATIO); meaninertia = vnl_hdr_f(b, VNL_MH_MEANINERTIA);
#define FI(name, id) name = vnl_field_i(b, id)
#define FF(name, id) name = vnl_field_f(b, id)
    FI(body_parentid, VNL_F_BODY_PARENTID); FI(body_rootid, VNL_F_BODY_ROOTID); FI(body_jntadr, VNL_F_BODY_JNTADR);
    FI(body_jntnum, VNL_F_BODY_JNTNUM); FI(jnt_type, VNL_F_JNT_TYPE); FI(jnt_qposadr, VNL_F_JNT_QPOSADR);
    FI(jnt_dofadr, VNL_F_JNT_DOFADR); FI(jnt_bodyid, VNL_F_JNT_BODYID); FI(jnt_limited, VNL_F_JNT_LIMITED);
    FI(dof_bodyid, VNL_F_DOF_BODYID); FI(dof_parentid, VNL_F_DOF_PARENTID); FI(geom_bodyid, VNL_F_GEOM_BODYID);
    FI(act_dofadr, VNL_F_ACT_DOFADR); FI(act_ctrllimited, VNL_F_ACT_CTRLLIMITED); FI(act_forcelimited, VNL_F_ACT_FORCELIMITED);
    FI(act_dyntype, VNL_F_ACT_DYNTYPE); FI(act_actadr, VNL_F_ACT_ACTADR); FI(pair_geom1, VNL_F_PAIR_GEOM1);
    FI(pair_geom2, VNL_F_PAIR_GEOM2); FI(pair_type, VNL_F_PAIR_TYPE); FI(limit_jnt, VNL_F_LIMIT_JNT);
    FF(body_pos, VNL_F_BODY_POS); FF(body_quat, VNL_F_BODY_QUAT); FF(body_ipos, VNL_F_BODY_IPOS); FF(body_iquat, VNL_F_BODY_IQUAT);
    FF(body_mass, VNL_F_BODY_MASS); FF(body_inertia, VNL_F_BODY_INERTIA); FF(body_invweight0, VNL_F_BODY_INVWEIGHT0);
    FF(jnt_pos, VNL_F_JNT_POS); FF(jnt_axis, VNL_F_JNT_AXIS); FF(jnt_stiffness, VNL_F_JNT_STIFFNESS); FF(jnt_range, VNL_F_JNT_RANGE);
    FF(jnt_margin, VNL_F_JNT_MARGIN); FF(jnt_solref, VNL_F_JNT_SOLREF); FF(jnt_solimp, VNL_F_JNT_SOLIMP);
    FF(dof_armature, VNL_F_DOF_ARMATURE); FF(dof_damping, VNL_F_DOF_DAMPING); FF(dof_invweight0, VNL_F_DOF_INVWEIGHT0);
    FF(qpos0, VNL_F_QPOS0); FF(qpos_spring, VNL_F_QPOS_SPRING); FF(geom_pos, VNL_F_GEOM_POS); FF(geom_quat, VNL_F_GEOM_QUAT);
    FF(geom_size, VNL_F_GEOM_SIZE); FF(act_gain, VNL_F_ACT_GAIN); FF(act_gear, VNL_F_ACT_GEAR); FF(act_ctrlrange, VNL_F_ACT_CTRLRANGE);
    FF(act_forcerange, VNL_F_ACT_FORCERANGE); FF(act_dynprm, VNL_F_ACT_DYNPRM); FF(pair_friction, VNL_F_PAIR_FRICTION);
    FF(pair_solref, VNL_F_PAIR_SOLREF); FF(pair_solimp, VNL_F_PAIR_SOLIMP); FF(pair_includemargin, VNL_F_PAIR_INCLUDEMARGIN);
#undef FI
#undef FF
  }
};

// ---- mjx/_src/math.py ----------------------------------------------------------------
template <class T> inline void cross3(const T* a, const T* b, T* r) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class T> inline T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> inline void quat_mul(const T* u, const T* v, T* r) {
  T w = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  T x = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  T y = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  T z = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
template <class T> inline void rotate(const T* vec, const T* q, T* r) {  // math.rotate
  T s = q[0];
  const T* u = q + 1;
  T ud = dot3(u, vec), uu = dot3(u, u), c[3];
  cross3(u, vec, c);
  for (int i = 0; i < 3; ++i) r[i] = T(2) * (ud * u[i]) + (s * s - uu) * vec[i] + T(2) * s * c[i];
}
template <class T> inline void quat_to_mat(const T* q, T* m) {  // math.quat_to_mat, row-major
  T q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  T q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[1] = T(2) * (q12 - q03); m[2] = T(2) * (q13 + q02);
  m[3] = T(2) * (q12 + q03); m[4] = q00 - q11 + q22 - q33; m[5] = T(2) * (q23 - q01);
  m[6] = T(2) * (q13 - q02); m[7] = T(2) * (q23 + q01); m[8] = q00 - q11 - q22 + q33;
}
template <class T> inline void axis_angle_to_quat(const T* axis, T angle, T* q) {
  T s = std::sin(angle * T(0.5)), c = std::cos(angle * T(0.5));
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
template <class T> inline T normalize3(T* v) {
  T n = std::sqrt(dot3(v, v));
  if (n == T(0)) { return n; }
  for (int i = 0; i < 3; ++i) v[i] /= n;
  return n;
}
template <class T> inline void normalize4(T* q) {
  T n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n == T(0)) return;
  for (int i = 0; i < 4; ++i) q[i] /= n;
}
// cinert layout: [Ixx Iyy Izz Ixy Ixz Iyz, m*cx m*cy m*cz, m]; spatial vectors are [ang, lin]
template <class T> inline void inert_mul(const T* i, const T* v, T* r) {  // math.inert_mul
  T ang[3] = {i[0] * v[0] + i[3] * v[1] + i[4] * v[2], i[3] * v[0] + i[1] * v[1] + i[5] * v[2], i[4] * v[0] + i[5] * v[1] + i[2] * v[2]};
  T c1[3], c2[3];
  cross3(i + 6, v + 3, c1);
  cross3(i + 6, v, c2);
  for (int k = 0; k < 3; ++k) { r[k] = ang[k] + c1[k]; r[3 + k] = i[9] * v[3 + k] - c2[k]; }
}
template <class T> inline void motion_cross(const T* u, const T* v, T* r) {  // math.motion_cross
  T a[3], b[3], c[3];
  cross3(u, v, a); cross3(u + 3, v, b); cross3(u, v + 3, c);
  for (int k = 0; k < 3; ++k) { r[k] = a[k]; r[3 + k] = b[k] + c[k]; }
}
template <class T> inline void motion_cross_force(const T* v, const T* f, T* r) {  // math.motion_cross_force
  T a[3], b[3], c[3];
  cross3(v, f, a); cross3(v + 3, f + 3, b); cross3(v, f + 3, c);
  for (int k = 0; k < 3; ++k) { r[k] = a[k] + b[k]; r[3 + k] = c[k]; }
}

// ---- per-env mjx.Data subset ---------------------------------------------------------------
template <class T> struct Data {
  std::vector<T> qpos, qvel, act, ctrl, qacc_warmstart;
  std::vector<T> xpos, xquat, xmat, xipos, ximat, xanchor, xaxis, geom_xpos, geom_xmat;
  std::vector<T> subtree_com, cinert, cdof, crb, qM, qLD, cvel, cdof_dot;
  std::vector<T> qfrc_passive, qfrc_bias, qfrc_actuator, act_dot, qfrc_smooth, qacc_smooth;
  std::vector<T> con_dist, con_pos, con_frame, con_friction, con_solref, con_solimp, con_includemargin;
  std::vector<int> con_body;
  std::vector<T> efc_J, efc_D, efc_aref, efc_pos, efc_force, qacc, qfrc_constraint;
  int solver_niter = 0, ls_niter = 0, ncon_active = 0, nlimit_active = 0;
  T time = 0;
  explicit Data(const ModelView& m) {
    qpos.assign(m.nq, 0); qvel.assign(m.nv, 0); act.assign(m.na, 0); ctrl.assign(m.nu, 0); qacc_warmstart.assign(m.nv, 0);
    xpos.assign(m.nbody * 3, 0); xquat.assign(m.nbody * 4, 0); xmat.assign(m.nbody * 9, 0); xipos.assign(m.nbody * 3, 0);
    ximat.assign(m.nbody * 9, 0); xanchor.assign(m.njnt * 3, 0); xaxis.assign(m.njnt * 3, 0);
    geom_xpos.assign(m.ngeom * 3, 0); geom_xmat.assign(m.ngeom * 9, 0);
    subtree_com.assign(m.nbody * 3, 0); cinert.assign(m.nbody * 10, 0); cdof.assign(m.nv * 6, 0); crb.assign(m.nbody * 10, 0);
    qM.assign(m.nv * m.nv, 0); qLD.assign(m.nv * m.nv, 0); cvel.assign(m.nbody * 6, 0); cdof_dot.assign(m.nv * 6, 0);
    qfrc_passive.assign(m.nv, 0); qfrc_bias.assign(m.nv, 0); qfrc_actuator.assign(m.nv, 0); act_dot.assign(m.na, 0);
    qfrc_smooth.assign(m.nv, 0); qacc_smooth.assign(m.nv, 0);
    con_dist.assign(m.ncon, 0); con_pos.assign(m.ncon * 3, 0); con_frame.assign(m.ncon * 9, 0); con_friction.assign(m.ncon * 5, 0);
    con_solref.assign(m.ncon * 2, 0); con_solimp.assign(m.ncon * 5, 0); con_includemargin.assign(m.ncon, 0); con_body.assign(m.ncon, 0);
    efc_J.assign(m.nefc * m.nv, 0); efc_D.assign(m.nefc, 0); efc_aref.assign(m.nefc, 0); efc_pos.assign(m.nefc, 0);
    efc_force.assign(m.nefc, 0); qacc.assign(m.nv, 0); qfrc_constraint.assign(m.nv, 0);
  }
};

// ---- smooth.kinematics -------------------------------------------------------------------
template <class T> void kinematics(const ModelView& m, Data<T>& d) {
  T* xpos = d.xpos.data(); T* xquat = d.xquat.data();
  xpos[0] = xpos[1] = xpos[2] = 0; xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
  for (int b = 1; b < m.nbody; ++b) {
    int p = m.body_parentid[b];
    T bp[3] = {T(m.body_pos[3 * b]), T(m.body_pos[3 * b + 1]), T(m.body_pos[3 * b + 2])};
    T bq[4] = {T(m.body_quat[4 * b]), T(m.body_quat[4 * b + 1]), T(m.body_quat[4 * b + 2]), T(m.body_quat[4 * b + 3])};
    T pos[3], quat[4], r[3];
    rotate(bp, xquat + 4 * p, r);
    for (int k = 0; k < 3; ++k) pos[k] = xpos[3 * p + k] + r[k];
    quat_mul(xquat + 4 * p, bq, quat);
    for (int jj = 0; jj < m.body_jntnum[b]; ++jj) {
      int j = m.body_jntadr[b] + jj, qa = m.jnt_qposadr[j];
      T jpos[3] = {T(m.jnt_pos[3 * j]), T(m.jnt_pos[3 * j + 1]), T(m.jnt_pos[3 * j + 2])};
      T jax[3] = {T(m.jnt_axis[3 * j]), T(m.jnt_axis[3 * j + 1]), T(m.jnt_axis[3 * j + 2])};
      if (m.jnt_type[j] == 0) {  // free: anchor = qpos[:3], axis = z; pos/quat from qpos, quat normalised and written back
        for (int k = 0; k < 3; ++k) { d.xanchor[3 * j + k] = d.qpos[qa + k]; pos[k] = d.qpos[qa + k]; }
        d.xaxis[3 * j] = 0; d.xaxis[3 * j + 1] = 0; d.xaxis[3 * j + 2] = 1;
        for (int k = 0; k < 4; ++k) quat[k] = d.qpos[qa + 3 + k];
        normalize4(quat);
        for (int k = 0; k < 4; ++k) d.qpos[qa + 3 + k] = quat[k];
      } else {  // hinge
        T anchor[3], axis[3], qloc[4], q2[4];
        rotate(jpos, quat, anchor);
        for (int k = 0; k < 3; ++k) anchor[k] += pos[k];
        rotate(jax, quat, axis);
        for (int k = 0; k < 3; ++k) { d.xanchor[3 * j + k] = anchor[k]; d.xaxis[3 * j + k] = axis[k]; }
        axis_angle_to_quat(jax, d.qpos[qa] - T(m.qpos0[qa]), qloc);
        quat_mul(quat, qloc, q2);
        for (int k = 0; k < 4; ++k) quat[k] = q2[k];
        rotate(jpos, quat, r);  // correct for off-center rotation
        for (int k = 0; k < 3; ++k) pos[k] = anchor[k] - r[k];
      }
    }
    for (int k = 0; k < 3; ++k) xpos[3 * b + k] = pos[k];
    for (int k = 0; k < 4; ++k) xquat[4 * b + k] = quat[k];
  }
  for (int b = 0; b < m.nbody; ++b) {
    quat_to_mat(xquat + 4 * b, d.xmat.data() + 9 * b);
    T ip[3] = {T(m.body_ipos[3 * b]), T(m.body_ipos[3 * b + 1]), T(m.body_ipos[3 * b + 2])}, r[3], q[4];
    T iq[4] = {T(m.body_iquat[4 * b]), T(m.body_iquat[4 * b + 1]), T(m.body_iquat[4 * b + 2]), T(m.body_iquat[4 * b + 3])};
    rotate(ip, xquat + 4 * b, r);
    for (int k = 0; k < 3; ++k) d.xipos[3 * b + k] = xpos[3 * b + k] + r[k];
    quat_mul(xquat + 4 * b, iq, q);
    quat_to_mat(q, d.ximat.data() + 9 * b);
  }
  for (int g = 0; g < m.ngeom; ++g) {
    int b = m.geom_bodyid[g];
    T gp[3] = {T(m.geom_pos[3 * g]), T(m.geom_pos[3 * g + 1]), T(m.geom_pos[3 * g + 2])}, r[3], q[4];
    T gq[4] = {T(m.geom_quat[4 * g]), T(m.geom_quat[4 * g + 1]), T(m.geom_quat[4 * g + 2]), T(m.geom_quat[4 * g + 3])};
    rotate(gp, xquat + 4 * b, r);
    for (int k = 0; k < 3; ++k) d.geom_xpos[3 * g + k] = xpos[3 * b + k] + r[k];
    quat_mul(xquat + 4 * b, gq, q);
    quat_to_mat(q, d.geom_xmat.data() + 9 * g);
  }
}

// ---- smooth.com_pos ------------------------------------------------------------------------
template <class T> void com_pos(const ModelView& m, Data<T>& d) {
  std::vector<T> pos(m.nbody * 3), mass(m.nbody);
  for (int b = 0; b < m.nbody; ++b) {
    mass[b] = T(m.body_mass[b]);
    for (int k = 0; k < 3; ++k) pos[3 * b + k] = d.xipos[3 * b + k] * T(m.body_mass[b]);
  }
  for (int b = m.nbody - 1; b > 0; --b) {
    int p = m.body_parentid[b];
    mass[p] += mass[b];
    for (int k = 0; k < 3; ++k) pos[3 * p + k] += pos[3 * b + k];
  }
  for (int b = 0; b < m.nbody; ++b)
    for (int k = 0; k < 3; ++k)
      d.subtree_com[3 * b + k] = (mass[b] < T(kMinVal)) ? d.xipos[3 * b + k] : pos[3 * b + k] / std::max(mass[b], T(kMinVal));
  for (int b = 0; b < m.nbody; ++b) {  // inert_com: (ximat * inertia) @ ximat.T + h @ h.T * mass
    const T* R = d.ximat.data() + 9 * b;
    const T* rc = d.subtree_com.data() + 3 * m.body_rootid[b];
    T off[3] = {d.xipos[3 * b] - rc[0], d.xipos[3 * b + 1] - rc[1], d.xipos[3 * b + 2] - rc[2]};
    T I[3] = {T(m.body_inertia[3 * b]), T(m.body_inertia[3 * b + 1]), T(m.body_inertia[3 * b + 2])}, ms = T(m.body_mass[b]);
    T in[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) in[3 * r + c] = R[3 * r] * I[0] * R[3 * c] + R[3 * r + 1] * I[1] * R[3 * c + 1] + R[3 * r + 2] * I[2] * R[3 * c + 2];
    T oo = dot3(off, off);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) in[3 * r + c] += (((r == c) ? oo : T(0)) - off[r] * off[c]) * ms;
    T* ci = d.cinert.data() + 10 * b;
    ci[0] = in[0]; ci[1] = in[4]; ci[2] = in[8]; ci[3] = in[1]; ci[4] = in[2]; ci[5] = in[5];
    ci[6] = off[0] * ms; ci[7] = off[1] * ms; ci[8] = off[2] * ms; ci[9] = ms;
  }
  for (int j = 0; j < m.njnt; ++j) {  // cdof = [axis, cross(axis, root_com - anchor)]
    int b = m.jnt_bodyid[j], dadr = m.jnt_dofadr[j];
    const T* rc = d.subtree_com.data() + 3 * m.body_rootid[b];
    T off[3] = {rc[0] - d.xanchor[3 * j], rc[1] - d.xanchor[3 * j + 1], rc[2] - d.xanchor[3 * j + 2]};
    if (m.jnt_type[j] == 0) {
      for (int a = 0; a < 3; ++a) {
        T* c = d.cdof.data() + 6 * (dadr + a);
        for (int k = 0; k < 6; ++k) c[k] = 0;
        c[3 + a] = 1;
        T ax[3] = {d.xmat[9 * b + a], d.xmat[9 * b + 3 + a], d.xmat[9 * b + 6 + a]};  // column a of xmat
        T* cr = d.cdof.data() + 6 * (dadr + 3 + a);
        cr[0] = ax[0]; cr[1] = ax[1]; cr[2] = ax[2];
        cross3(ax, off, cr + 3);
      }
    } else {
      T* c = d.cdof.data() + 6 * dadr;
      const T* ax = d.xaxis.data() + 3 * j;
      c[0] = ax[0]; c[1] = ax[1]; c[2] = ax[2];
      cross3(ax, off, c + 3);
    }
  }
}

// ---- smooth.crb + dense qM ------------------------------------------------------------------
template <class T> void crb(const ModelView& m, Data<T>& d) {
  d.crb = d.cinert;
  for (int b = m.nbody - 1; b > 0; --b) {
    int p = m.body_parentid[b];
    for (int k = 0; k < 10; ++k) d.crb[10 * p + k] += d.crb[10 * b + k];
  }
  for (int k = 0; k < 10; ++k) d.crb[k] = 0;
  std::fill(d.qM.begin(), d.qM.end(), T(0));
  int nv = m.nv;
  for (int i = 0; i < nv; ++i) {
    T f[6];
    inert_mul(d.crb.data() + 10 * m.dof_bodyid[i], d.cdof.data() + 6 * i, f);
    for (int j = i; j >= 0; j = m.dof_parentid[j]) {
      const T* c = d.cdof.data() + 6 * j;
      T v = f[0] * c[0] + f[1] * c[1] + f[2] * c[2] + f[3] * c[3] + f[4] * c[4] + f[5] * c[5];
      if (j == i) v += T(m.dof_armature[i]);
      d.qM[i * nv + j] = v;
      d.qM[j * nv + i] = v;
    }
  }
}

// dense Cholesky (lower) of an nv x nv SPD matrix; `jax.scipy.linalg.cho_factor` in smooth.factor_m
template <class T> void cho_factor(int n, const T* A, T* L) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      T s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = (i == j) ? std::sqrt(s) : s / L[j * n + j];
    }
}
template <class T> void cho_solve(int n, const T* L, const T* b, T* x) {
  for (int i = 0; i < n; ++i) {
    T s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    T s = x[i];
    for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}
template <class T> void mul_m(const ModelView& m, const Data<T>& d, const T* v, T* out) {  // support.mul_m (dense)
  for (int i = 0; i < m.nv; ++i) {
    T s = 0;
    for (int j = 0; j < m.nv; ++j) s += d.qM[i * m.nv + j] * v[j];
    out[i] = s;
  }
}

// ---- smooth.com_vel ----------------------------------------------------------------------
template <class T> void com_vel(const ModelView& m, Data<T>& d) {
  for (int k = 0; k < 6; ++k) d.cvel[k] = 0;
  for (int b = 1; b < m.nbody; ++b) {
    T cvel[6];
    for (int k = 0; k < 6; ++k) cvel[k] = d.cvel[6 * m.body_parentid[b] + k];
    for (int jj = 0; jj < m.body_jntnum[b]; ++jj) {
      int j = m.body_jntadr[b] + jj, da = m.jnt_dofadr[j];
      if (m.jnt_type[j] == 0) {
        for (int a = 0; a < 3; ++a)
          for (int k = 0; k < 6; ++k) { cvel[k] += d.cdof[6 * (da + a) + k] * d.qvel[da + a]; d.cdof_dot[6 * (da + a) + k] = 0; }
        for (int a = 3; a < 6; ++a) motion_cross(cvel, d.cdof.data() + 6 * (da + a), d.cdof_dot.data() + 6 * (da + a));
        for (int a = 3; a < 6; ++a)
          for (int k = 0; k < 6; ++k) cvel[k] += d.cdof[6 * (da + a) + k] * d.qvel[da + a];
      } else {
        motion_cross(cvel, d.cdof.data() + 6 * da, d.cdof_dot.data() + 6 * da);
        for (int k = 0; k < 6; ++k) cvel[k] += d.cdof[6 * da + k] * d.qvel[da];
      }
    }
    for (int k = 0; k < 6; ++k) d.cvel[6 * b + k] = cvel[k];
  }
}

// ---- passive.passive (joint springs + dampers; no tendons / fluid on this path) --------------
template <class T> void passive(const ModelView& m, Data<T>& d) {
  for (int j = 0; j < m.njnt; ++j) {
    int da = m.jnt_dofadr[j], qa = m.jnt_qposadr[j];
    if (m.jnt_type[j] == 0) {
      // free-joint stiffness is 0 on every model of the reference; the quaternion spring term is omitted
      for (int a = 0; a < 3; ++a) d.qfrc_passive[da + a] = -T(m.jnt_stiffness[j]) * (d.qpos[qa + a] - T(m.qpos_spring[qa + a]));
      for (int a = 3; a < 6; ++a) d.qfrc_passive[da + a] = 0;
    } else {
      d.qfrc_passive[da] = -T(m.jnt_stiffness[j]) * (d.qpos[qa] - T(m.qpos_spring[qa]));
    }
  }
  for (int i = 0; i < m.nv; ++i) d.qfrc_passive[i] -= T(m.dof_damping[i]) * d.qvel[i];
}

// ---- smooth.rne ----------------------------------------------------------------------------
template <class T> void rne(const ModelView& m, Data<T>& d) {
  std::vector<T> cacc(m.nbody * 6), cfrc(m.nbody * 6);
  for (int k = 0; k < 3; ++k) { cacc[k] = 0; cacc[3 + k] = -T(m.gravity[k]); }
  for (int b = 1; b < m.nbody; ++b) {
    for (int k = 0; k < 6; ++k) cacc[6 * b + k] = cacc[6 * m.body_parentid[b] + k];
    for (int jj = 0; jj < m.body_jntnum[b]; ++jj) {
      int j = m.body_jntadr[b] + jj, da = m.jnt_dofadr[j], nd = (m.jnt_type[j] == 0) ? 6 : 1;
      for (int a = 0; a < nd; ++a)
        for (int k = 0; k < 6; ++k) cacc[6 * b + k] += d.cdof_dot[6 * (da + a) + k] * d.qvel[da + a];
    }
  }
  for (int b = 0; b < m.nbody; ++b) {
    T f1[6], iv[6], f2[6];
    inert_mul(d.cinert.data() + 10 * b, cacc.data() + 6 * b, f1);
    inert_mul(d.cinert.data() + 10 * b, d.cvel.data() + 6 * b, iv);
    motion_cross_force(d.cvel.data() + 6 * b, iv, f2);
    for (int k = 0; k < 6; ++k) cfrc[6 * b + k] = f1[k] + f2[k];
  }
  for (int b = m.nbody - 1; b > 0; --b)
    for (int k = 0; k < 6; ++k) cfrc[6 * m.body_parentid[b] + k] += cfrc[6 * b + k];
  for (int i = 0; i < m.nv; ++i) {
    const T* c = d.cdof.data() + 6 * i;
    const T* f = cfrc.data() + 6 * m.dof_bodyid[i];
    d.qfrc_bias[i] = c[0] * f[0] + c[1] * f[1] + c[2] * f[2] + c[3] * f[3] + c[4] * f[4] + c[5] * f[5];
  }
}

// ---- forward.fwd_actuation (joint transmissions, fixed gain, no bias, filter / none dynamics) ---
template <class T> void fwd_actuation(const ModelView& m, Data<T>& d) {
  std::fill(d.qfrc_actuator.begin(), d.qfrc_actuator.end(), T(0));
  for (int u = 0; u < m.nu; ++u) {
    T ctrl = d.ctrl[u];
    if (m.act_ctrllimited[u]) ctrl = std::min(std::max(ctrl, T(m.act_ctrlrange[2 * u])), T(m.act_ctrlrange[2 * u + 1]));
    T ctrl_act = ctrl;
    int aa = m.act_actadr[u];
    if (aa >= 0) {
      d.act_dot[aa] = (ctrl - d.act[aa]) / std::max(T(m.act_dynprm[u]), T(kMinVal));
      ctrl_act = d.act[aa];
    }
    T force = T(m.act_gain[u]) * ctrl_act;
    if (m.act_forcelimited[u]) force = std::min(std::max(force, T(m.act_forcerange[2 * u])), T(m.act_forcerange[2 * u + 1]));
    d.qfrc_actuator[m.act_dofadr[u]] += T(m.act_gear[u]) * force;
  }
}

// ---- collision_driver.collision: static plane-{sphere,capsule,ellipsoid} pairs ----------------
template <class T> void make_frame(const T* a, T* frame) {  // math.make_frame
  T n[3] = {a[0], a[1], a[2]};
  normalize3(n);
  T b[3] = {0, 0, 0};
  if (T(-0.5) < n[1] && n[1] < T(0.5)) b[1] = 1; else b[2] = 1;
  T nb = dot3(n, b);
  for (int k = 0; k < 3; ++k) b[k] -= n[k] * nb;
  normalize3(b);
  for (int k = 0; k < 3; ++k) { frame[k] = n[k]; frame[3 + k] = b[k]; }
  cross3(n, b, frame + 6);
}
template <class T> void collision(const ModelView& m, Data<T>& d) {
  int c = 0;
  for (int p = 0; p < m.npair; ++p) {
    int g1 = m.pair_geom1[p], g2 = m.pair_geom2[p];
    const T* pm = d.geom_xmat.data() + 9 * g1;
    const T* gm = d.geom_xmat.data() + 9 * g2;
    T n[3] = {pm[2], pm[5], pm[8]};
    const T* ppos = d.geom_xpos.data() + 3 * g1;
    const T* gpos = d.geom_xpos.data() + 3 * g2;
    const float* size = m.geom_size + 3 * g2;
    int ncp = 1;
    if (m.pair_type[p] == 2) {  // plane_sphere
      T diff[3] = {gpos[0] - ppos[0], gpos[1] - ppos[1], gpos[2] - ppos[2]};
      T dist = dot3(diff, n) - T(size[0]);
      d.con_dist[c] = dist;
      for (int k = 0; k < 3; ++k) d.con_pos[3 * c + k] = gpos[k] - n[k] * (T(size[0]) + T(0.5) * dist);
      make_frame(n, d.con_frame.data() + 9 * c);
    } else if (m.pair_type[p] == 3) {  // plane_capsule: two plane-sphere tests, frame aligned with the capsule axis
      ncp = 2;
      T axis[3] = {gm[2], gm[5], gm[8]};
      T na = dot3(n, axis), b[3] = {axis[0] - n[0] * na, axis[1] - n[1] * na, axis[2] - n[2] * na};
      T bn = normalize3(b);
      if (bn < T(0.5)) {
        b[0] = 0; b[1] = 0; b[2] = 0;
        if (T(-0.5) < n[1] && n[1] < T(0.5)) b[1] = 1; else b[2] = 1;
      }
      T frame[9] = {n[0], n[1], n[2], b[0], b[1], b[2], 0, 0, 0};
      cross3(n, b, frame + 6);
      for (int e = 0; e < 2; ++e) {
        T sgn = e == 0 ? T(1) : T(-1);
        T sp[3] = {gpos[0] + sgn * axis[0] * T(size[1]), gpos[1] + sgn * axis[1] * T(size[1]), gpos[2] + sgn * axis[2] * T(size[1])};
        T diff[3] = {sp[0] - ppos[0], sp[1] - ppos[1], sp[2] - ppos[2]};
        T dist = dot3(diff, n) - T(size[0]);
        d.con_dist[c + e] = dist;
        for (int k = 0; k < 3; ++k) d.con_pos[3 * (c + e) + k] = sp[k] - n[k] * (T(size[0]) + T(0.5) * dist);
        for (int k = 0; k < 9; ++k) d.con_frame[9 * (c + e) + k] = frame[k];
      }
    } else {  // plane_ellipsoid: support point in the -normal direction
      T s[3] = {T(size[0]), T(size[1]), T(size[2])};
      T ln[3] = {gm[0] * n[0] + gm[3] * n[1] + gm[6] * n[2], gm[1] * n[0] + gm[4] * n[1] + gm[7] * n[2], gm[2] * n[0] + gm[5] * n[1] + gm[8] * n[2]};
      T sup[3] = {ln[0] * s[0], ln[1] * s[1], ln[2] * s[2]};
      normalize3(sup);
      for (int k = 0; k < 3; ++k) sup[k] = -sup[k] * s[k];
      T pos[3];
      for (int r = 0; r < 3; ++r) pos[r] = gpos[r] + gm[3 * r] * sup[0] + gm[3 * r + 1] * sup[1] + gm[3 * r + 2] * sup[2];
      T diff[3] = {pos[0] - ppos[0], pos[1] - ppos[1], pos[2] - ppos[2]};
      T dist = dot3(n, diff);
      d.con_dist[c] = dist;
      for (int k = 0; k < 3; ++k) d.con_pos[3 * c + k] = pos[k] - n[k] * dist * T(0.5);
      make_frame(n, d.con_frame.data() + 9 * c);
    }
    for (int e = 0; e < ncp; ++e) {
      d.con_body[c + e] = m.geom_bodyid[g2];
      for (int k = 0; k < 5; ++k) { d.con_friction[5 * (c + e) + k] = T(m.pair_friction[5 * p + k]); d.con_solimp[5 * (c + e) + k] = T(m.pair_solimp[5 * p + k]); }
      d.con_solref[2 * (c + e)] = T(m.pair_solref[2 * p]); d.con_solref[2 * (c + e) + 1] = T(m.pair_solref[2 * p + 1]);
      d.con_includemargin[c + e] = T(m.pair_includemargin[p]);
    }
    c += ncp;
  }
}

// ---- constraint.make_constraint ------------------------------------------------------------
template <class T> void kbi(const ModelView& m, const T* solref, const T* solimp, T pos, T& k, T& b, T& imp) {  // constraint._kbi
  T timeconst = std::max(solref[0], T(2) * T(m.timestep)), dampratio = solref[1];  // refsafe
  T dmin = std::min(std::max(solimp[0], T(kMinImp)), T(kMaxImp)), dmax = std::min(std::max(solimp[1], T(kMinImp)), T(kMaxImp));
  T width = std::max(T(kMinVal), solimp[2]), mid = std::min(std::max(solimp[3], T(kMinImp)), T(kMaxImp)), power = std::max(T(1), solimp[4]);
  k = T(1) / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  b = T(2) / (dmax * timeconst);
  if (solref[0] <= 0) k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0) b = -solref[1] / dmax;
  T imp_x = std::abs(pos) / width;
  T imp_a = (T(1) / std::pow(mid, power - T(1))) * std::pow(imp_x, power);
  T imp_b = T(1) - (T(1) / std::pow(T(1) - mid, power - T(1))) * std::pow(T(1) - imp_x, power);
  T imp_y = imp_x < mid ? imp_a : imp_b;
  imp = dmin + imp_y * (dmax - dmin);
  imp = std::min(std::max(imp, dmin), dmax);
  if (imp_x > T(1)) imp = dmax;
}
template <class T> void make_constraint(const ModelView& m, Data<T>& d) {
  int nv = m.nv;
  std::fill(d.efc_J.begin(), d.efc_J.end(), T(0));
  d.ncon_active = d.nlimit_active = 0;
  std::vector<T> invweight(m.nefc), solref(2 * m.nefc), solimp(5 * m.nefc);
  // _instantiate_limit_slide_hinge
  for (int r = 0; r < m.nlimit; ++r) {
    int j = m.limit_jnt[r];
    T q = d.qpos[m.jnt_qposadr[j]];
    T dmin = q - T(m.jnt_range[2 * j]), dmax = T(m.jnt_range[2 * j + 1]) - q;
    T pos = std::min(dmin, dmax) - T(m.jnt_margin[j]);
    bool active = pos < 0;
    d.efc_J[r * nv + m.jnt_dofadr[j]] = T((dmin < dmax) ? 1 : -1) * T(active ? 1 : 0);
    d.efc_pos[r] = pos;
    invweight[r] = T(m.dof_invweight0[m.jnt_dofadr[j]]);
    solref[2 * r] = T(m.jnt_solref[2 * j]); solref[2 * r + 1] = T(m.jnt_solref[2 * j + 1]);
    for (int k = 0; k < 5; ++k) solimp[5 * r + k] = T(m.jnt_solimp[5 * j + k]);
    d.nlimit_active += active;
  }
  // _instantiate_contact (pyramidal, condim 3); everything is multiplied by `active`
  for (int c = 0; c < m.ncon; ++c) {
    T dist = d.con_dist[c] - d.con_includemargin[c];
    bool active = dist < 0;
    d.ncon_active += active;
    int body = d.con_body[c];
    const T* rc = d.subtree_com.data() + 3 * m.body_rootid[body];
    T off[3] = {d.con_pos[3 * c] - rc[0], d.con_pos[3 * c + 1] - rc[1], d.con_pos[3 * c + 2] - rc[2]};
    std::vector<T> jacp(3 * nv, T(0));  // support.jac: ancestor dofs of `body`; body1 is the world (plane) -> zero
    for (int b = body; b > 0; b = m.body_parentid[b])
      for (int jj = 0; jj < m.body_jntnum[b]; ++jj) {
        int j = m.body_jntadr[b] + jj, nd = m.jnt_type[j] == 0 ? 6 : 1;
        for (int a = 0; a < nd; ++a) {
          int i = m.jnt_dofadr[j] + a;
          T cr[3];
          cross3(d.cdof.data() + 6 * i, off, cr);
          for (int k = 0; k < 3; ++k) jacp[k * nv + i] = d.cdof[6 * i + 3 + k] + cr[k];
        }
      }
    const T* fr = d.con_frame.data() + 9 * c;
    T t = T(m.body_invweight0[2 * body]);  // + body_invweight0[world] = 0
    for (int dir = 0; dir < 2; ++dir)
      for (int s = 0; s < 2; ++s) {
        int r = m.nlimit + 4 * c + 2 * dir + s;
        T f = d.con_friction[5 * c + dir] * (s == 0 ? T(1) : T(-1));
        for (int i = 0; i < nv; ++i) {
          T dn = fr[0] * jacp[i] + fr[1] * jacp[nv + i] + fr[2] * jacp[2 * nv + i];
          T dt = fr[3 * (dir + 1)] * jacp[i] + fr[3 * (dir + 1) + 1] * jacp[nv + i] + fr[3 * (dir + 1) + 2] * jacp[2 * nv + i];
          d.efc_J[r * nv + i] = active ? dn + dt * f : T(0);
        }
        invweight[r] = active ? (t + f * f * t) * T(2) * f * f / T(m.impratio) : T(0);
        d.efc_pos[r] = active ? dist : T(0);
        solref[2 * r] = active ? d.con_solref[2 * c] : T(0); solref[2 * r + 1] = active ? d.con_solref[2 * c + 1] : T(0);
        for (int k = 0; k < 5; ++k) solimp[5 * r + k] = active ? d.con_solimp[5 * c + k] : T(0);
      }
  }
  for (int r = 0; r < m.nefc; ++r) {
    T k, b, imp;
    kbi(m, solref.data() + 2 * r, solimp.data() + 5 * r, d.efc_pos[r], k, b, imp);
    T R = std::max(invweight[r] * (T(1) - imp) / imp, T(kMinVal));
    T vel = 0;
    for (int i = 0; i < nv; ++i) vel += d.efc_J[r * nv + i] * d.qvel[i];
    d.efc_aref[r] = -b * vel - k * imp * d.efc_pos[r];
    d.efc_D[r] = T(1) / R;
  }
}

// ---- solver.solve --------------------------------------------------------------------------
template <class T> struct Ctx {
  std::vector<T> qacc, qfrc_constraint, Jaref, efc_force, Ma, grad, Mgrad, search;
  std::vector<char> active;
  T gauss = 0, cost = std::numeric_limits<T>::infinity(), prev_cost = 0;
  int niter = 0;
};
template <class T> void update_constraint(const ModelView& m, const Data<T>& d, Ctx<T>& c) {
  int nv = m.nv;
  T s = 0;
  for (int r = 0; r < m.nefc; ++r) {
    c.active[r] = c.Jaref[r] < 0;
    c.efc_force[r] = d.efc_D[r] * -c.Jaref[r] * T(c.active[r]);
    s += d.efc_D[r] * c.Jaref[r] * c.Jaref[r] * T(c.active[r]);
  }
  for (int i = 0; i < nv; ++i) {
    T q = 0;
    for (int r = 0; r < m.nefc; ++r) q += d.efc_J[r * nv + i] * c.efc_force[r];
    c.qfrc_constraint[i] = q;
  }
  T g = 0;
  for (int i = 0; i < nv; ++i) g += (c.Ma[i] - d.qfrc_smooth[i]) * (c.qacc[i] - d.qacc_smooth[i]);
  c.gauss = T(0.5) * g;
  c.prev_cost = c.cost;
  c.cost = T(0.5) * s + c.gauss;
}
template <class T> void update_gradient(const ModelView& m, const Data<T>& d, Ctx<T>& c) {
  int nv = m.nv;
  for (int i = 0; i < nv; ++i) c.grad[i] = c.Ma[i] - d.qfrc_smooth[i] - c.qfrc_constraint[i];
  if (m.solver == 1) {
    cho_solve(nv, d.qLD.data(), c.grad.data(), c.Mgrad.data());
  } else {  // Newton: H = qM + J^T diag(D * active) J
    std::vector<T> h(d.qM), L(nv * nv, T(0));
    for (int r = 0; r < m.nefc; ++r)
      if (c.active[r])
        for (int i = 0; i < nv; ++i) {
          T ji = d.efc_J[r * nv + i] * d.efc_D[r];
          if (ji != 0)
            for (int j = 0; j < nv; ++j) h[i * nv + j] += ji * d.efc_J[r * nv + j];
        }
    cho_factor(nv, h.data(), L.data());
    cho_solve(nv, L.data(), c.grad.data(), c.Mgrad.data());
  }
}
template <class T> void ctx_create(const ModelView& m, const Data<T>& d, const T* qacc, bool grad, Ctx<T>& c) {
  int nv = m.nv;
  c.qacc.assign(qacc, qacc + nv);
  c.qfrc_constraint.assign(nv, 0); c.Jaref.assign(m.nefc, 0); c.efc_force.assign(m.nefc, 0); c.Ma.assign(nv, 0);
  c.grad.assign(nv, 0); c.Mgrad.assign(nv, 0); c.search.assign(nv, 0); c.active.assign(m.nefc, 0);
  c.gauss = 0; c.cost = std::numeric_limits<T>::infinity(); c.prev_cost = 0; c.niter = 0;
  for (int r = 0; r < m.nefc; ++r) {
    T s = 0;
    for (int i = 0; i < nv; ++i) s += d.efc_J[r * nv + i] * qacc[i];
    c.Jaref[r] = s - d.efc_aref[r];
  }
  mul_m(m, d, qacc, c.Ma.data());
  update_constraint(m, d, c);
  if (grad) {
    update_gradient(m, d, c);
    for (int i = 0; i < nv; ++i) c.search[i] = -c.Mgrad[i];
  }
}
template <class T> struct LSPoint { T alpha, cost, d0, d1; };
template <class T> int linesearch(const ModelView& m, const Data<T>& d, Ctx<T>& c) {  // solver._linesearch
  int nv = m.nv, ne = m.nefc;
  T sn = 0;
  for (int i = 0; i < nv; ++i) sn += c.search[i] * c.search[i];
  T smag = std::sqrt(sn) * T(m.meaninertia) * T(std::max(1, nv));
  T gtol = T(m.tolerance) * T(m.ls_tolerance) * smag;
  std::vector<T> mv(nv), jv(ne), quad(3 * ne);
  mul_m(m, d, c.search.data(), mv.data());
  for (int r = 0; r < ne; ++r) {
    T s = 0;
    for (int i = 0; i < nv; ++i) s += d.efc_J[r * nv + i] * c.search[i];
    jv[r] = s;
  }
  T qg[3] = {c.gauss, 0, 0}, a1 = 0, a2 = 0, a3 = 0;
  for (int i = 0; i < nv; ++i) { a1 += c.search[i] * c.Ma[i]; a2 += c.search[i] * d.qfrc_smooth[i]; a3 += c.search[i] * mv[i]; }
  qg[1] = a1 - a2; qg[2] = T(0.5) * a3;
  for (int r = 0; r < ne; ++r) {
    quad[3 * r] = T(0.5) * c.Jaref[r] * c.Jaref[r] * d.efc_D[r];
    quad[3 * r + 1] = jv[r] * c.Jaref[r] * d.efc_D[r];
    quad[3 * r + 2] = T(0.5) * jv[r] * jv[r] * d.efc_D[r];
  }
  auto point = [&](T alpha) {
    T q0 = 0, q1 = 0, q2 = 0;
    for (int r = 0; r < ne; ++r)
      if (c.Jaref[r] + alpha * jv[r] < 0) { q0 += quad[3 * r]; q1 += quad[3 * r + 1]; q2 += quad[3 * r + 2]; }
    q0 += qg[0]; q1 += qg[1]; q2 += qg[2];
    LSPoint<T> p;
    p.alpha = alpha;
    p.cost = alpha * alpha * q2 + alpha * q1 + q0;
    p.d0 = T(2) * alpha * q2 + q1;
    p.d1 = T(2) * q2 + (q2 == 0 ? T(kMinVal) : T(0));
    return p;
  };
  LSPoint<T> p0 = point(T(0));
  LSPoint<T> lo = point(p0.alpha - p0.d0 / p0.d1), hi;
  bool lesser = lo.d0 < p0.d0;
  hi = lesser ? p0 : lo;
  lo = lesser ? lo : p0;
  bool swap = true;
  int it = 0;
  while (true) {
    bool done = it >= m.ls_iterations;
    done |= !swap;
    done |= (lo.d0 < 0) && (lo.d0 > -gtol);
    done |= (hi.d0 > 0) && (hi.d0 < gtol);
    if (done) break;
    LSPoint<T> lo_next = point(lo.alpha - lo.d0 / lo.d1), hi_next = point(hi.alpha - hi.d0 / hi.d1);
    LSPoint<T> mid = point(T(0.5) * (lo.alpha + hi.alpha));
    bool s1 = (lo.d0 > 0) || (lo.d0 < lo_next.d0);
    if (s1) lo = lo_next;
    bool s2 = (mid.d0 < 0) && (lo.d0 < mid.d0);
    if (s2) lo = mid;
    bool s3 = (hi_next.d0 < 0) && (lo.d0 < hi_next.d0);
    if (s3) lo = hi_next;
    bool s4 = (hi.d0 < 0) || (hi.d0 > hi_next.d0);
    if (s4) hi = hi_next;
    bool s5 = (mid.d0 > 0) && (hi.d0 > mid.d0);
    if (s5) hi = mid;
    bool s6 = (lo_next.d0 > 0) && (hi.d0 > lo_next.d0);
    if (s6) hi = lo_next;
    swap = s1 || s2 || s3 || s4 || s5 || s6;
    ++it;
  }
  bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  T alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  // `x + improved * v * alpha` in the reference: a NaN alpha propagates even when not improved
  T ia = improved ? alpha : T(0) * alpha;
  for (int i = 0; i < nv; ++i) { c.qacc[i] += c.search[i] * ia; c.Ma[i] += mv[i] * ia; }
  for (int r = 0; r < ne; ++r) c.Jaref[r] += jv[r] * ia;
  return it;
}
template <class T> void solve(const ModelView& m, Data<T>& d) {
  int nv = m.nv;
  Ctx<T> warm, smth, c;
  ctx_create(m, d, d.qacc_warmstart.data(), false, warm);
  ctx_create(m, d, d.qacc_smooth.data(), false, smth);
  const T* q0 = (warm.cost < smth.cost) ? d.qacc_warmstart.data() : d.qacc_smooth.data();
  ctx_create(m, d, q0, true, c);
  T scale = T(m.meaninertia) * T(std::max(1, nv));
  d.ls_niter = 0;
  auto body = [&]() {
    d.ls_niter += linesearch(m, d, c);
    std::vector<T> pg(c.grad), pMg(c.Mgrad);
    update_constraint(m, d, c);
    update_gradient(m, d, c);
    if (m.solver == 2) {
      for (int i = 0; i < nv; ++i) c.search[i] = -c.Mgrad[i];
    } else {  // Polak-Ribiere
      T num = 0, den = 0;
      for (int i = 0; i < nv; ++i) { num += c.grad[i] * (c.Mgrad[i] - pMg[i]); den += pg[i] * pMg[i]; }
      T beta = std::max(T(0), num / std::max(T(kMinVal), den));
      for (int i = 0; i < nv; ++i) c.search[i] = -c.Mgrad[i] + beta * c.search[i];
    }
    c.niter++;
  };
  if (m.iterations == 1) {
    body();
  } else {
    while (true) {
      T improvement = (c.prev_cost - c.cost) / scale;
      T gn = 0;
      for (int i = 0; i < nv; ++i) gn += c.grad[i] * c.grad[i];
      T gradient = std::sqrt(gn) / scale;
      bool done = c.niter >= m.iterations;
      done |= improvement < T(m.tolerance);
      done |= gradient < T(m.tolerance);
      if (done) break;
      body();
    }
  }
  d.qacc = c.qacc;
  d.qacc_warmstart = c.qacc;
  d.qfrc_constraint = c.qfrc_constraint;
  d.efc_force = c.efc_force;
  d.solver_niter = c.niter;
}

// ---- forward.forward / forward.euler / forward.step -------------------------------------------
template <class T> void forward(const ModelView& m, Data<T>& d) {
  kinematics(m, d);
  com_pos(m, d);
  crb(m, d);
  cho_factor(m.nv, d.qM.data(), d.qLD.data());
  collision(m, d);
  make_constraint(m, d);  // fwd_position ends here; efc_aref uses qvel, which fwd_velocity does not change
  com_vel(m, d);
  passive(m, d);
  rne(m, d);
  fwd_actuation(m, d);
  for (int i = 0; i < m.nv; ++i) d.qfrc_smooth[i] = d.qfrc_passive[i] - d.qfrc_bias[i] + d.qfrc_actuator[i];
  cho_solve(m.nv, d.qLD.data(), d.qfrc_smooth.data(), d.qacc_smooth.data());
  if (m.nefc == 0) { d.qacc = d.qacc_smooth; return; }
  solve(m, d);
}
template <class T> void euler(const ModelView& m, Data<T>& d) {
  int nv = m.nv;
  T dt = T(m.timestep);
  std::vector<T> qacc(d.qacc);
  if (m.eulerdamp) {
    std::vector<T> h(d.qM), L(nv * nv, T(0)), f(nv);
    for (int i = 0; i < nv; ++i) { h[i * nv + i] += dt * T(m.dof_damping[i]); f[i] = d.qfrc_smooth[i] + d.qfrc_constraint[i]; }
    cho_factor(nv, h.data(), L.data());
    cho_solve(nv, L.data(), f.data(), qacc.data());
  }
  for (int a = 0; a < m.na; ++a) d.act[a] += d.act_dot[a] * dt;
  for (int i = 0; i < nv; ++i) d.qvel[i] += qacc[i] * dt;
  for (int j = 0; j < m.njnt; ++j) {  // _integrate_pos with the NEW qvel (semi-implicit)
    int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
    if (m.jnt_type[j] == 0) {
      for (int k = 0; k < 3; ++k) d.qpos[qa + k] += d.qvel[da + k] * dt;
      T v[3] = {d.qvel[da + 3], d.qvel[da + 4], d.qvel[da + 5]};
      T norm = normalize3(v);  // math.quat_integrate
      T qr[4], q2[4];
      axis_angle_to_quat(v, dt * norm, qr);
      quat_mul(d.qpos.data() + qa + 3, qr, q2);
      normalize4(q2);
      for (int k = 0; k < 4; ++k) d.qpos[qa + 3 + k] = q2[k];
    } else {
      d.qpos[qa] += d.qvel[da] * dt;
    }
  }
  d.time += dt;
}
template <class T> void step(const ModelView& m, Data<T>& d) { forward(m, d); euler(m, d); }

// ---- task logic: envs/rodent.py ----------------------------------------------------------------
struct TaskView {
  const uint32_t* w;
  int T_, ref_len, sub_clip_len, ntrack, njidx, napp, nee, nframes, obs_size, traj_size, com_ref_idx, torso;
  int reward_old_state, term_mean, use_subclip, obs_qfrc, com_from_field, rot_body, traj_old_frame, ract_action, metrics_raw;
  float healthy_lo, healthy_hi, term_threshold, body_err_mult, done_rtrunk;
  double w_rcom, w_rvel, w_rtrunk, w_rquat, w_ract, w_rapp;
  const float *position, *quaternion, *joints, *body_positions, *velocity, *angular_velocity, *joints_velocity, *center_of_mass;
  const int *body_idxs, *ee_idx, *app_idx, *app_ref_idx, *joint_col;
  explicit TaskView(const uint32_t* b) : w(b) {
    T_ = vnl_hdr_i(b, VNL_TH_CLIP_LEN); ref_len = vnl_hdr_i(b, VNL_TH_REF_LEN); sub_clip_len = vnl_hdr_i(b, VNL_TH_SUB_CLIP_LEN);
    ntrack = vnl_hdr_i(b, VNL_TH_NTRACK); njidx = vnl_hdr_i(b, VNL_TH_NJIDX); napp = vnl_hdr_i(b, VNL_TH_NAPP); nee = vnl_hdr_i(b, VNL_TH_NEE);
    nframes = vnl_hdr_i(b, VNL_TH_NFRAMES); obs_size = vnl_hdr_i(b, VNL_TH_OBS_SIZE); traj_size = vnl_hdr_i(b, VNL_TH_TRAJ_SIZE);
    com_ref_idx = vnl_hdr_i(b, VNL_TH_COM_REF_IDX); torso = vnl_hdr_i(b, VNL_TH_TORSO_BODY);
    healthy_lo = vnl_hdr_f(b, VNL_TH_HEALTHY_LO); healthy_hi = vnl_hdr_f(b, VNL_TH_HEALTHY_HI);
    term_threshold = vnl_hdr_f(b, VNL_TH_TERM_THRESHOLD); body_err_mult = vnl_hdr_f(b, VNL_TH_BODY_ERR_MULT);
    reward_old_state = vnl_hdr_i(b, VNL_TH_REWARD_OLD_STATE); term_mean = vnl_hdr_i(b, VNL_TH_TERM_MEAN);
    use_subclip = vnl_hdr_i(b, VNL_TH_USE_SUBCLIP); obs_qfrc = vnl_hdr_i(b, VNL_TH_OBS_QFRC);
    com_from_field = vnl_hdr_i(b, VNL_TH_COM_FROM_FIELD); done_rtrunk = vnl_hdr_f(b, VNL_TH_DONE_RTRUNK);
    rot_body = vnl_hdr_i(b, VNL_TH_ROT_BODY); traj_old_frame = vnl_hdr_i(b, VNL_TH_TRAJ_OLD_FRAME);
    ract_action = vnl_hdr_i(b, VNL_TH_RACT_ACTION); metrics_raw = vnl_hdr_i(b, VNL_TH_METRICS_RAW);
    // the reward weights are short decimal literals in the reference (0.01, 0.20, 1e-4 ...); the blob holds them as
    // fp32, so the fp64 build snaps them back to the literal (7 significant digits) instead of inheriting fp32 rounding
    auto lit = [](float f) { if (f == 0.0f) return 0.0; double e = std::pow(10.0, 6 - std::floor(std::log10(std::fabs((double)f)))); return std::round((double)f * e) / e; };
    w_rcom = lit(vnl_hdr_f(b, VNL_TH_W_RCOM)); w_rvel = lit(vnl_hdr_f(b, VNL_TH_W_RVEL)); w_rtrunk = lit(vnl_hdr_f(b, VNL_TH_W_RTRUNK));
    w_rquat = lit(vnl_hdr_f(b, VNL_TH_W_RQUAT)); w_ract = lit(vnl_hdr_f(b, VNL_TH_W_RACT)); w_rapp = lit(vnl_hdr_f(b, VNL_TH_W_RAPP));
    center_of_mass = vnl_field_f(b, VNL_T_CENTER_OF_MASS);
    position = vnl_field_f(b, VNL_T_POSITION); quaternion = vnl_field_f(b, VNL_T_QUATERNION); joints = vnl_field_f(b, VNL_T_JOINTS);
    body_positions = vnl_field_f(b, VNL_T_BODY_POSITIONS); velocity = vnl_field_f(b, VNL_T_VELOCITY);
    angular_velocity = vnl_field_f(b, VNL_T_ANGULAR_VELOCITY); joints_velocity = vnl_field_f(b, VNL_T_JOINTS_VELOCITY);
    body_idxs = vnl_field_i(b, VNL_T_BODY_IDXS); ee_idx = vnl_field_i(b, VNL_T_EE_IDX); app_idx = vnl_field_i(b, VNL_T_APP_IDX);
    app_ref_idx = vnl_field_i(b, VNL_T_APP_REF_IDX); joint_col = vnl_field_i(b, VNL_T_JOINT_COL);
  }
};
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// _calculate_termination (rodent.py:241-264) on (qpos, xpos) of the given state and frame
template <class T> T termination(const ModelView& m, const TaskView& t, const T* qpos, const T* xpos, int frame) {
  int f = clampi(frame, 0, t.T_ - 1), nj = m.nq - 7;
  T ej = 0;
  for (int j = 0; j < nj; ++j) ej += std::abs(T(t.joints[f * nj + j]) - qpos[7 + j]);
  T col[3] = {0, 0, 0};  // jp.linalg.norm(matrix, ord=1) = max column abs-sum (quirk Q9)
  for (int b = 0; b < t.ntrack; ++b)
    for (int k = 0; k < 3; ++k) col[k] += std::abs(T(t.body_positions[(f * t.ntrack + b) * 3 + k]) - xpos[3 * t.body_idxs[b] + k]);
  T eb = std::max(col[0], std::max(col[1], col[2]));
  if (t.term_mean) {  // humanoid.py:256-258: jp.mean(jp.abs(.)) over the joints and over all body coordinates
    ej = ej / T(nj);
    eb = (col[0] + col[1] + col[2]) / T(3 * t.ntrack);
  }
  T error = T(0.5) * T(t.body_err_mult) * eb + T(0.5) * ej;
  return T(1) - error / T(t.term_threshold);
}
// _get_obs (rodent.py:318-344)
template <class T> void get_obs(const ModelView& m, const TaskView& t, const Data<T>& d, T* obs) {
  int o = 0;
  for (int i = 0; i < m.nq; ++i) obs[o++] = d.qpos[i];
  for (int i = 0; i < m.nv; ++i) obs[o++] = d.qvel[i];
  if (!t.obs_qfrc) return;  // humanoid.py:359-366: [qpos, qvel] only
  for (int i = 0; i < m.nv; ++i) obs[o++] = d.qfrc_actuator[i];
  for (int e = 0; e < t.nee; ++e)
    for (int k = 0; k < 3; ++k) obs[o++] = d.xpos[3 * t.ee_idx[e] + k];
}
// _get_traj (rodent.py:346-448); window start = clamp(cur_frame + 1, 0, T - ref_len) (dynamic_slice semantics)
template <class T> void get_traj(const ModelView& m, const TaskView& t, const Data<T>& d, int cur_frame, T* traj) {
  int s = clampi(cur_frame + 1, 0, t.T_ - t.ref_len), nj = m.nq - 7, o = 0;
  const T* R = d.xmat.data() + 9 * t.rot_body;  // rodent.py:385 xmat[1]; ant.py:333 xmat[0] (world: identity)
  for (int w = 0; w < t.ref_len; ++w)  // get_reference_appendages_pos: filtered[:, app_idx] clamped (Q5)
    for (int a = 0; a < t.napp; ++a)
      for (int k = 0; k < 3; ++k) traj[o++] = T(t.body_positions[((s + w) * t.ntrack + t.app_ref_idx[a]) * 3 + k]);
  for (int w = 0; w < t.ref_len; ++w)  // bodies, local frame: (ref - xpos) @ xmat[torso]
    for (int b = 0; b < t.ntrack; ++b) {
      T df[3];
      for (int k = 0; k < 3; ++k) df[k] = T(t.body_positions[((s + w) * t.ntrack + b) * 3 + k]) - d.xpos[3 * t.body_idxs[b] + k];
      for (int c = 0; c < 3; ++c) traj[o++] = df[0] * R[c] + df[1] * R[3 + c] + df[2] * R[6 + c];
    }
  for (int w = 0; w < t.ref_len; ++w)  // bodies, global frame
    for (int b = 0; b < t.ntrack; ++b)
      for (int k = 0; k < 3; ++k) traj[o++] = T(t.body_positions[((s + w) * t.ntrack + b) * 3 + k]) - d.xpos[3 * t.body_idxs[b] + k];
  for (int w = 0; w < t.ref_len; ++w) {  // root, local frame
    T df[3];
    for (int k = 0; k < 3; ++k) df[k] = T(t.position[(s + w) * 3 + k]) - d.qpos[k];
    for (int c = 0; c < 3; ++c) traj[o++] = df[0] * R[c] + df[1] * R[3 + c] + df[2] * R[6 + c];
  }
  for (int w = 0; w < t.ref_len; ++w)  // joints: (ref.joints - qpos[7:])[:, joint_idxs] clamped (Q6)
    for (int j = 0; j < t.njidx; ++j) traj[o++] = T(t.joints[(s + w) * nj + t.joint_col[j]]) - d.qpos[7 + t.joint_col[j]];
}
template <class T> bool data_has_nan(const Data<T>& d) {
  auto chk = [](const std::vector<T>& v) { for (T x : v) if (std::isnan(x)) return true; return false; };
  return chk(d.qpos) || chk(d.qvel) || chk(d.act) || chk(d.qacc) || chk(d.qacc_warmstart) || chk(d.xpos) || chk(d.xquat) ||
         chk(d.subtree_com) || chk(d.qfrc_actuator) || chk(d.qfrc_bias) || chk(d.qfrc_constraint) || chk(d.cvel) || chk(d.efc_force);
}

struct StateIO {  // host mirrors of VnlState with double payloads
  double *qpos, *qvel, *act, *qacc_warmstart, *xpos, *xquat, *subtree_com, *qfrc_actuator;
  int32_t *cur_frame, *sub_clip_frame;
};
struct OutIO { double *obs, *traj, *reward, *done, *metrics; int32_t* stats; };

template <class T> void load_state(const ModelView& m, const StateIO& s, int e, Data<T>& d) {
  for (int i = 0; i < m.nq; ++i) d.qpos[i] = T(s.qpos[e * m.nq + i]);
  for (int i = 0; i < m.nv; ++i) d.qvel[i] = T(s.qvel[e * m.nv + i]);
  for (int i = 0; i < m.na; ++i) d.act[i] = s.act ? T(s.act[e * m.na + i]) : T(0);
  for (int i = 0; i < m.nv; ++i) d.qacc_warmstart[i] = s.qacc_warmstart ? T(s.qacc_warmstart[e * m.nv + i]) : T(0);
}
template <class T> void store_state(const ModelView& m, const StateIO& s, int e, const Data<T>& d, int torso) {
  for (int i = 0; i < m.nq; ++i) s.qpos[e * m.nq + i] = d.qpos[i];
  for (int i = 0; i < m.nv; ++i) s.qvel[e * m.nv + i] = d.qvel[i];
  for (int i = 0; i < m.na; ++i) s.act[e * m.na + i] = d.act[i];
  for (int i = 0; i < m.nv; ++i) s.qacc_warmstart[e * m.nv + i] = d.qacc_warmstart[i];
  for (int i = 0; i < m.nbody * 3; ++i) s.xpos[e * m.nbody * 3 + i] = d.xpos[i];
  for (int i = 0; i < m.nbody * 4; ++i) s.xquat[e * m.nbody * 4 + i] = d.xquat[i];
  for (int k = 0; k < 3; ++k) s.subtree_com[e * 3 + k] = d.subtree_com[3 * torso + k];
  for (int i = 0; i < m.nv; ++i) s.qfrc_actuator[e * m.nv + i] = d.qfrc_actuator[i];
}

// RodentTracking.step (rodent.py:178-239) for env e
template <class T> void env_step(const ModelView& m, const TaskView& t, int e, const StateIO& in, const double* action,
                                 const StateIO& out, const OutIO& o) {
  Data<T> d(m);
  load_state(m, in, e, d);
  std::vector<T> qpos_old(d.qpos), xpos_old(m.nbody * 3), qvel_old(d.qvel), qfrc_old(m.nv);
  for (int i = 0; i < m.nbody * 3; ++i) xpos_old[i] = T(in.xpos[e * m.nbody * 3 + i]);
  for (int i = 0; i < m.nv; ++i) qfrc_old[i] = in.qfrc_actuator ? T(in.qfrc_actuator[e * m.nv + i]) : T(0);
  T com_old[3] = {in.subtree_com ? T(in.subtree_com[e * 3]) : T(0), in.subtree_com ? T(in.subtree_com[e * 3 + 1]) : T(0),
                  in.subtree_com ? T(in.subtree_com[e * 3 + 2]) : T(0)};
  for (int u = 0; u < m.nu; ++u) d.ctrl[u] = T(action[e * m.nu + u]);
  int stats[4] = {0, 0, 0, 0};
  for (int f = 0; f < t.nframes; ++f) {  // pipeline_step
    step(m, d);
    stats[0] += d.solver_niter; stats[1] += d.ls_niter; stats[2] += d.ncon_active; stats[3] += d.nlimit_active;
  }
  int frame_old = in.cur_frame[e];
  int cur_frame = frame_old + 1, sub_clip_frame = in.sub_clip_frame[e] + 1;
  std::vector<T> obs(t.obs_size), traj(t.traj_size);
  get_obs(m, t, d, obs.data());
  get_traj(m, t, d, t.traj_old_frame ? frame_old : cur_frame, traj.data());  // ant.py:182: window from the OLD info
  // _calculate_reward (rodent.py:266-316): every reference lookup uses the OLD cur_frame (Q3)
  int f = clampi(frame_old, 0, t.T_ - 1);
  // humanoid.py:275 evaluates every term on the PRE-step state (`data_c = state.pipeline_state`), the rodent on the new one
  const bool old = t.reward_old_state != 0;
  const T* r_qvel = old ? qvel_old.data() : d.qvel.data();
  const T* r_qpos = old ? qpos_old.data() : d.qpos.data();
  const T* r_qfrc = old ? qfrc_old.data() : d.qfrc_actuator.data();
  const T* r_com = old ? com_old : d.subtree_com.data() + 3 * t.torso;
  T s = 0;
  for (int k = 0; k < 3; ++k) {
    T cref = t.com_from_field ? T(t.center_of_mass[f * 3 + k]) : T(t.body_positions[(f * t.ntrack + t.com_ref_idx) * 3 + k]);
    T df = r_com[k] - cref;
    s += df * df;
  }
  T rcom = std::exp(T(-100) * std::sqrt(s));
  s = 0;
  for (int i = 0; i < m.nv; ++i) {
    T ref = i < 3 ? T(t.velocity[f * 3 + i]) : (i < 6 ? T(t.angular_velocity[f * 3 + i - 3]) : T(t.joints_velocity[f * (m.nv - 6) + i - 6]));
    T df = r_qvel[i] - ref;
    s += df * df;
  }
  T rvel = std::exp(T(-0.1) * std::sqrt(s));
  T rtrunk = termination(m, t, qpos_old.data(), xpos_old.data(), frame_old);  // OLD state, OLD frame (Q2)
  T qc[4] = {r_qpos[3], r_qpos[4], r_qpos[5], r_qpos[6]};
  T qr[4] = {T(t.quaternion[4 * f]), T(t.quaternion[4 * f + 1]), T(t.quaternion[4 * f + 2]), T(t.quaternion[4 * f + 3])};
  normalize4(qc); normalize4(qr);  // _bounded_quat_dist (rodent.py:450-470)
  T dq = qc[0] * qr[0] + qc[1] * qr[1] + qc[2] * qr[2] + qc[3] * qr[3];
  T dist = std::min(T(1), T(2) * dq * dq - T(1));
  T rquat = std::exp(T(-2) * std::abs(T(0.5) * std::acos(dist)));
  s = 0;
  for (int i = 0; i < m.nv; ++i) s += r_qfrc[i] * r_qfrc[i];
  T ract = T(-0.015) * (s / T(m.nv));
  if (t.ract_action) {  // ant.py:251
    s = 0;
    for (int u = 0; u < m.nu; ++u) s += T(action[e * m.nu + u]) * T(action[e * m.nu + u]);
    ract = T(0.01) * T(-0.015) * s / T(m.nu);
  }
  s = 0;
  for (int a = 0; a < t.napp; ++a)
    for (int k = 0; k < 3; ++k) { T df = d.xpos[3 * t.app_idx[a] + k] - T(t.body_positions[(f * t.ntrack + t.app_ref_idx[a]) * 3 + k]); s += df * df; }
  T rapp = t.napp > 0 ? std::exp(T(-400) * std::sqrt(s)) : T(0);  // no appendage term in humanoid.py:200-205
  T healthy = r_qpos[2] < T(t.healthy_lo) ? T(0) : T(1);
  if (r_qpos[2] > T(t.healthy_hi)) healthy = 0;
  T done = rtrunk < T(t.done_rtrunk) ? T(1) : T(0);  // rodent.py:213 (scaled rtrunk < 0) / humanoid.py:199 (rtrunk < 0.5 before scaling)
  const T raw[6] = {rcom, rvel, rtrunk, rquat, ract, rapp};
  rcom *= T(t.w_rcom); rvel *= T(t.w_rvel); rapp *= T(t.w_rapp); rtrunk *= T(t.w_rtrunk); rquat *= T(t.w_rquat); ract *= T(t.w_ract);  // rodent.py:193-199 / ant.py:186-192
  T total = rcom + rvel + rtrunk + rquat + ract + rapp;
  T sub_healthy = (!t.use_subclip || sub_clip_frame < t.sub_clip_len) ? T(1) : T(0);
  done = std::max(T(1) - healthy, done);
  done = std::max(T(1) - sub_healthy, done);
  T reward = std::isnan(total) ? T(0) : total;  // jp.nan_to_num (inf -> large finite, as jnp does)
  if (std::isinf(reward)) reward = reward > 0 ? std::numeric_limits<T>::max() : std::numeric_limits<T>::lowest();
  if (data_has_nan(d)) done = 1;
  store_state(m, out, e, d, t.torso);
  out.cur_frame[e] = cur_frame; out.sub_clip_frame[e] = sub_clip_frame;
  for (int i = 0; i < t.obs_size; ++i) {
    T v = obs[i];
    if (std::isnan(v)) v = 0;
    if (std::isinf(v)) v = v > 0 ? std::numeric_limits<T>::max() : std::numeric_limits<T>::lowest();
    o.obs[e * t.obs_size + i] = v;
  }
  for (int i = 0; i < t.traj_size; ++i) o.traj[e * t.traj_size + i] = traj[i];
  o.reward[e] = reward; o.done[e] = done;
  double* mt = o.metrics + 7 * e;
  mt[0] = rcom; mt[1] = rvel; mt[2] = rtrunk; mt[3] = rquat; mt[4] = ract; mt[5] = rapp; mt[6] = rtrunk;
  if (t.metrics_raw) { for (int k = 0; k < 6; ++k) mt[k] = raw[k]; mt[6] = raw[2]; }  // ant.py:203-210
  if (o.stats) for (int k = 0; k < 4; ++k) o.stats[4 * e + k] = stats[k];
}

// RodentTracking.reset tail (rodent.py:148-176)
template <class T> void env_reset(const ModelView& m, const TaskView& t, int e, const StateIO& in, const StateIO& out, const OutIO& o) {
  Data<T> d(m);
  for (int i = 0; i < m.nq; ++i) d.qpos[i] = T(in.qpos[e * m.nq + i]);
  for (int i = 0; i < m.nv; ++i) d.qvel[i] = T(in.qvel[e * m.nv + i]);
  forward(m, d);
  int start = in.cur_frame[e];
  std::vector<T> obs(t.obs_size), traj(t.traj_size);
  get_traj(m, t, d, start, traj.data());
  get_obs(m, t, d, obs.data());
  T term = termination(m, t, d.qpos.data(), d.xpos.data(), start);
  store_state(m, out, e, d, t.torso);
  out.cur_frame[e] = start; out.sub_clip_frame[e] = 0;
  for (int i = 0; i < t.obs_size; ++i) o.obs[e * t.obs_size + i] = obs[i];
  for (int i = 0; i < t.traj_size; ++i) o.traj[e * t.traj_size + i] = traj[i];
  o.reward[e] = 0; o.done[e] = 0;
  for (int k = 0; k < 6; ++k) o.metrics[7 * e + k] = 0;
  o.metrics[7 * e + 6] = term;
  if (o.stats) { o.stats[4 * e] = d.solver_niter; o.stats[4 * e + 1] = d.ls_niter; o.stats[4 * e + 2] = d.ncon_active; o.stats[4 * e + 3] = d.nlimit_active; }
}

template <class T> void forward_dump(const ModelView& m, int e, const StateIO& in, const double* ctrl, double* dump, size_t stride) {
  Data<T> d(m);
  load_state(m, in, e, d);
  for (int u = 0; u < m.nu; ++u) d.ctrl[u] = ctrl ? T(ctrl[e * m.nu + u]) : T(0);
  forward(m, d);
  double* p = dump + stride * e;
  auto put = [&](const std::vector<T>& v) { for (T x : v) *p++ = double(x); };
  put(d.xpos); put(d.xquat); put(d.xmat); put(d.xipos); put(d.ximat); put(d.xanchor); put(d.xaxis); put(d.subtree_com);
  put(d.cinert); put(d.cdof); put(d.crb); put(d.qM); put(d.cvel); put(d.cdof_dot); put(d.qfrc_passive); put(d.qfrc_bias);
  put(d.qfrc_actuator); put(d.act_dot); put(d.qfrc_smooth); put(d.qacc_smooth); put(d.con_dist); put(d.con_pos); put(d.con_frame);
  put(d.efc_pos); put(d.efc_D); put(d.efc_aref); put(d.efc_J); put(d.qacc); put(d.qfrc_constraint); put(d.efc_force);
  *p++ = d.solver_niter; *p++ = d.ls_niter; *p++ = d.ncon_active; *p++ = d.nlimit_active;
}
size_t dump_size(const ModelView& m) {
  size_t nb = m.nbody, nv = m.nv;
  return nb * 3 + nb * 4 + nb * 9 + nb * 3 + nb * 9 + m.njnt * 3 + m.njnt * 3 + nb * 3 + nb * 10 + nv * 6 + nb * 10 + nv * nv + nb * 6 +
         nv * 6 + nv * 5 + m.na + m.ncon * 13 + m.nefc * 3 + (size_t)m.nefc * nv + nv * 2 + m.nefc + 4;
}

template <class T> void pipeline_steps(const ModelView& m, int e, int nsteps, const StateIO& in, const double* ctrl, const StateIO& out, int32_t* stats) {
  Data<T> d(m);
  load_state(m, in, e, d);
  for (int u = 0; u < m.nu; ++u) d.ctrl[u] = ctrl ? T(ctrl[e * m.nu + u]) : T(0);
  int st[4] = {0, 0, 0, 0};
  for (int f = 0; f < nsteps; ++f) {
    step(m, d);
    st[0] += d.solver_niter; st[1] += d.ls_niter; st[2] += d.ncon_active; st[3] += d.nlimit_active;
  }
  store_state(m, out, e, d, 1);
  if (stats) for (int k = 0; k < 4; ++k) stats[4 * e + k] = st[k];
}

}  // namespace

extern "C" {

int vnl_oracle_step(const uint32_t* model, const uint32_t* task, int precision, int B, const StateIO* in, const double* action,
                    const StateIO* out, const OutIO* o, int nthreads) {
  if (model[0] != VNL_MAGIC_MODEL || task[0] != VNL_MAGIC_TASK) return -1;
  ModelView m(model);
  TaskView t(task);
  (void)nthreads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
  for (int e = 0; e < B; ++e) {
    if (precision == 64) env_step<double>(m, t, e, *in, action, *out, *o);
    else env_step<float>(m, t, e, *in, action, *out, *o);
  }
  return 0;
}

int vnl_oracle_reset(const uint32_t* model, const uint32_t* task, int precision, int B, const StateIO* in, const StateIO* out,
                     const OutIO* o, int nthreads) {
  if (model[0] != VNL_MAGIC_MODEL || task[0] != VNL_MAGIC_TASK) return -1;
  ModelView m(model);
  TaskView t(task);
  (void)nthreads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
  for (int e = 0; e < B; ++e) {
    if (precision == 64) env_reset<double>(m, t, e, *in, *out, *o);
    else env_reset<float>(m, t, e, *in, *out, *o);
  }
  return 0;
}

int vnl_oracle_pipeline_step(const uint32_t* model, int precision, int B, int nsteps, const StateIO* in, const double* ctrl,
                             const StateIO* out, int32_t* stats, int nthreads) {
  if (model[0] != VNL_MAGIC_MODEL) return -1;
  ModelView m(model);
  (void)nthreads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
  for (int e = 0; e < B; ++e) {
    if (precision == 64) pipeline_steps<double>(m, e, nsteps, *in, ctrl, *out, stats);
    else pipeline_steps<float>(m, e, nsteps, *in, ctrl, *out, stats);
  }
  return 0;
}

size_t vnl_oracle_dump_size(const uint32_t* model) { return dump_size(ModelView(model)); }

int vnl_oracle_forward_dump(const uint32_t* model, int precision, int B, const StateIO* in, const double* ctrl, double* dump) {
  if (model[0] != VNL_MAGIC_MODEL) return -1;
  ModelView m(model);
  size_t stride = dump_size(m);
  for (int e = 0; e < B; ++e) {
    if (precision == 64) forward_dump<double>(m, e, *in, ctrl, dump, stride);
    else forward_dump<float>(m, e, *in, ctrl, dump, stride);
  }
  return 0;
}

int vnl_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"

// vnl_kernels.cu -- fused MJX-style physics + imitation reward/obs/termination step for sm_100a.
//
// One CTA per environment.  The whole env step -- n_frames x (forward dynamics, constraint
// solve, semi-implicit Euler) and the clip-indexed reward / observation / trajectory window /
// termination -- runs in ONE launch with the per-env working set resident in shared memory;
// HBM is touched only for the state in / state + observations out (about 8 KB per env step).
//
// Formulation (differs from the dense one XLA executes for the reference, same mathematics):
//   * joint-space inertia kept tree-sparse (MuJoCo qM layout), factorised as L^T D L without
//     fill-in; the unit-triangular factor is inverted once per factorisation (the inverse has
//     the same ancestor sparsity), so every later M^-1 x is two parallel sparse mat-vecs
//     instead of two serial triangular sweeps;
//   * the constraint Jacobian is never materialised: limit rows are one-hot, contact rows are
//     frame . (v_lin + w x r) of the contact's body, so J x and J^T f are sums over the
//     ancestor chain of a handful of bodies; only ACTIVE rows (pos < 0) are kept -- inactive
//     rows contribute exactly zero to every solver quantity in the reference formulation;
//   * subtree sums (composite inertia, RNE backward pass) use the DFS-preorder body numbering:
//     a subtree is a contiguous id range, so they are flat reductions, not level-by-level scans.
//
// Reference semantics restated: mjx.forward / mjx.step as reached from envs/rodent.py:148,181
// and RodentTracking.step / reset (envs/rodent.py:119-470) -- see oracle/vnl_oracle.cpp, the
// CPU restatement these kernels are tested against.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/vnl_blob.h"
#include "vnl_device.cuh"
#include "vnl_kernels.h"

namespace vnl {

#define PROF(c, i) do { if ((c).prof) { __syncthreads(); if (threadIdx.x == 0) { long long t_ = clock64(); (c).prof[i] += t_ - (c).t0; (c).t0 = t_; } } } while (0)

__host__ __device__ inline int align4(int x) { return (x + 3) & ~3; }

// Shared-memory layout (float offsets), identical on host and device.
__host__ __device__ inline void make_layout(const Dims& d, Lay& L) {
  int o = 0;
#define A(name, n) L.name = o; o += align4(n)
  A(qpos, d.nq); A(qvel, d.nv); A(act, d.na); A(ctrl, d.nu); A(warm, d.nv);
  A(xpos, d.nbody * 3); A(xquat, d.nbody * 4); A(xipos, d.nbody * 3); A(xanchor, d.njnt * 3); A(xaxis, d.njnt * 3);
  A(rcom, d.nbody * 3);
  A(cinert, d.nbody * 10); A(crb, d.nbody * 10); A(cdof, d.nv * 6); A(cdofdot, d.nv * 6); A(cvel, d.nbody * 6);
  A(cacc, d.nbody * 6); A(cfrc, d.nbody * 6);
  A(M, d.nM); A(Lf, d.nM); A(K, d.nM);
  A(qfrc_smooth, d.nv); A(qacc_smooth, d.nv); A(qfrc_act, d.nv); A(act_dot, d.na);
  A(lim_dof, d.nlimit); A(lim_sign, d.nlimit); A(limrow_of_dof, d.nv);
  A(cbody, d.ncon); A(crel, d.ncon * 3); A(cframe, d.ncon * 9); A(cmu, d.ncon); A(cwrench, d.ncon * 6);
  A(efcD, d.nefc); A(aref, d.nefc); A(Jaref, d.nefc); A(Jv, d.nefc);
  A(qacc, d.nv); A(Ma, d.nv); A(grad, d.nv); A(Mgrad, d.nv); A(search, d.nv); A(Mv, d.nv); A(qfrc_con, d.nv);
  A(tmpv, d.nv);
  A(red, 2 * kMaxWarps * 12);
  A(ints, 16);
  A(mcol8, (d.nM + 3) / 4); A(mrow8, (d.nM + 3) / 4); A(madr16, (d.nv + 2) / 2); A(dadr16, (d.nv + 2) / 2);
  A(dent16, (d.nM - d.nv + 1) / 2); A(drow8, (d.nM - d.nv + 3) / 4); A(dls8, (d.maxdepth + 2 + 3) / 4); A(dld8, (d.nv + 3) / 4);
#undef A
  L.total = o;
}

// CTA-uniform context.  It lives at the start of dynamic shared memory so that the big phases can be
// real (non-inlined) functions: one copy of each in the instruction stream instead of one per call site.
struct __align__(16) Ctx {
  Dims d;
  Lay L;
  float* s;            // shared memory base of the float arrays
  const uint32_t* mb;  // model blob (global)
  uint32_t foff[VNL_F_MODEL_COUNT];
  const uint8_t *mcol, *mrow, *drow, *dls, *dld;
  const uint16_t *madr, *dadr, *dent;
  long long* prof;
  long long t0;
  int nt, nw;
  __device__ __forceinline__ const int* fi(int f) const { return (const int*)(mb + foff[f]); }
  __device__ __forceinline__ const float* ff(int f) const { return (const float*)(mb + foff[f]); }
};
constexpr int kCtxFloats = (int)((sizeof(Ctx) + 15) / 16 * 4);
#define TID ((int)threadIdx.x)
#define LANE ((int)(threadIdx.x & 31))
#define WARP ((int)(threadIdx.x >> 5))

// Sum N per-thread values over the CTA; every thread returns the same totals (fixed order ->
// bit-reproducible).  One __syncthreads per call (scratch is double-buffered).
template <int N>
__device__ __forceinline__ void block_sum(Ctx& c, float (&v)[N]) {
#pragma unroll
  for (int n = 0; n < N; ++n) v[n] = warp_sum(v[n]);
  float* buf = c.s + c.L.red;
  if (LANE == 0) {
#pragma unroll
    for (int n = 0; n < N; ++n) buf[WARP * N + n] = v[n];
  }
  __syncthreads();
  const int nw = c.nw;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    float t = 0.0f;
#pragma unroll 1
    for (int w = 0; w < nw; ++w) t += buf[w * N + n];
    v[n] = t;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// sparse-inertia helpers
// ---------------------------------------------------------------------------------------------
// out = M x   (tree-sparse M, rows hold [diag, ancestors...])
__device__ __noinline__ void mul_m(Ctx& c, const float* x, float* out) {
  const uint16_t* madr = c.madr; const uint8_t* mcol = c.mcol;
  const uint16_t* dadr = c.dadr; const uint16_t* dent = c.dent; const uint8_t* drow = c.drow;
  const float* M = c.s + c.L.M;
  for (int i = TID; i < c.d.nv; i += c.nt) {
    float a0 = 0.0f, a1 = 0.0f;
    int a = madr[i];
    const int ae = madr[i + 1];
    for (; a + 1 < ae; a += 2) { a0 += M[a] * x[mcol[a]]; a1 += M[a + 1] * x[mcol[a + 1]]; }
    if (a < ae) a0 += M[a] * x[mcol[a]];
    int k = dadr[i];
    const int ke = dadr[i + 1];
    for (; k + 1 < ke; k += 2) { a0 += M[dent[k]] * x[drow[k]]; a1 += M[dent[k + 1]] * x[drow[k + 1]]; }
    if (k < ke) a0 += M[dent[k]] * x[drow[k]];
    out[i] = a0 + a1;
  }
}

// L^T D L factorisation of `src` (+ dt * damping on the diagonal when `damp`), then K = L^-1.
// Leaves: K off-diagonals in L.K, 1 / D in the diagonal slots of L.K.
__device__ __noinline__ void factor(Ctx& c, const float* src, bool damp) {
  const int nv = c.d.nv, nM = c.d.nM;
  const uint16_t* madr = c.madr; const uint8_t* mcol = c.mcol; const uint8_t* mrow = c.mrow;
  float* Lf = c.s + c.L.Lf;
  float* K = c.s + c.L.K;
  const float* damping = c.ff(VNL_F_DOF_DAMPING);
  for (int e = TID; e < nM; e += c.nt) {
    float v = src[e];
    if (damp && mcol[e] == mrow[e]) v += c.d.timestep * damping[mrow[e]];
    Lf[e] = v;
  }
  __syncthreads();
  // eliminate dofs from the leaves: for ancestors a >= 1 of k and entries cidx >= 0 of that ancestor's row
  //   L[anc_a(k)][cidx] -= L[k][a] * L[k][a + cidx] / L[k][0]     (rows stay un-normalised until the end)
  for (int k = nv - 1; k > 0; --k) {
    const int base = madr[k], dk = madr[k + 1] - base - 1;  // dk = number of ancestors
    if (dk > 0) {
      const float inv = 1.0f / Lf[base];
      for (int a = 1 + WARP; a <= dk; a += c.nw) {  // targets are distinct for distinct (a, cidx)
        const int tb = madr[mcol[base + a]];
        const float t = Lf[base + a] * inv;
        for (int cidx = LANE; cidx <= dk - a; cidx += 32) Lf[tb + cidx] -= t * Lf[base + a + cidx];
      }
    }
    __syncthreads();
  }
  // normalise rows: Lhat = L / diag; keep 1 / D in the diagonal slot of K
  for (int e = TID; e < nM; e += c.nt) {
    const int b0 = madr[mrow[e]];
    if (e == b0) K[e] = 1.0f / Lf[e];
    else Lf[e] = Lf[e] / Lf[b0];
  }
  __syncthreads();
  // K = Lhat^-1 by levels of dof depth (Lhat K = I): for a dof i with ancestors anc_1..anc_d,
  //   K[i][c] = -( Lhat[i][c] + sum_{a=1}^{c-1} Lhat[i][a] * K[anc_a(i)][c - a] ),  K[.][0] = 1 implicit
  const uint8_t* dls = c.dls; const uint8_t* dld = c.dld;
  for (int dl = 1; dl <= c.d.maxdepth; ++dl) {
    const int s0 = dls[dl], n = (dls[dl + 1] - s0) * dl;
    for (int it = TID; it < n; it += c.nt) {
      const int i = dld[s0 + it / dl], cc = 1 + it % dl, base = madr[i];
      float a0 = Lf[base + cc], a1 = 0.0f;
      int a = 1;
      for (; a + 1 < cc; a += 2) {
        a0 += Lf[base + a] * K[madr[mcol[base + a]] + cc - a];
        a1 += Lf[base + a + 1] * K[madr[mcol[base + a + 1]] + cc - a - 1];
      }
      if (a < cc) a0 += Lf[base + a] * K[madr[mcol[base + a]] + cc - a];
      K[base + cc] = -(a0 + a1);
    }
    __syncthreads();
  }
}

// x <- M^-1 x   via  K (D^-1 (K^T x));  `tmp` is nv scratch.
__device__ __noinline__ void solve_m(Ctx& c, const float* x, float* out, float* tmp) {
  const uint16_t* madr = c.madr; const uint8_t* mcol = c.mcol;
  const uint16_t* dadr = c.dadr; const uint16_t* dent = c.dent; const uint8_t* drow = c.drow;
  const float* K = c.s + c.L.K;
  for (int j = TID; j < c.d.nv; j += c.nt) {
    float a0 = x[j], a1 = 0.0f;
    int k = dadr[j];
    const int ke = dadr[j + 1];
    for (; k + 1 < ke; k += 2) { a0 += K[dent[k]] * x[drow[k]]; a1 += K[dent[k + 1]] * x[drow[k + 1]]; }
    if (k < ke) a0 += K[dent[k]] * x[drow[k]];
    tmp[j] = (a0 + a1) * K[madr[j]];
  }
  __syncthreads();
  for (int i = TID; i < c.d.nv; i += c.nt) {
    float a0 = tmp[i], a1 = 0.0f;
    int a = madr[i] + 1;
    const int ae = madr[i + 1];
    for (; a + 1 < ae; a += 2) { a0 += K[a] * tmp[mcol[a]]; a1 += K[a + 1] * tmp[mcol[a + 1]]; }
    if (a < ae) a0 += K[a] * tmp[mcol[a]];
    out[i] = a0 + a1;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// constraint Jacobian products on the compact active set
// ---------------------------------------------------------------------------------------------
// out[row] = (J x)[row]
__device__ __noinline__ void jmul(Ctx& c, const float* x, float* out) {
  const int* ints = (const int*)(c.s + c.L.ints);
  const int nl = ints[0], nc = ints[1];
  const uint16_t* madr = c.madr; const uint8_t* mcol = c.mcol;
  const int* lastdof = c.fi(VNL_F_BODY_LASTDOF);
  const float* cdof = c.s + c.L.cdof;
  const int* cbody = (const int*)(c.s + c.L.cbody);
  const int* lim_dof = (const int*)(c.s + c.L.lim_dof);
  const float* lim_sign = c.s + c.L.lim_sign;
  for (int k = WARP; k < nc; k += c.nw) {
    const int dl = lastdof[cbody[k]];
    float sacc[6] = {0, 0, 0, 0, 0, 0};
    if (dl >= 0) {
      for (int a = madr[dl] + LANE; a < madr[dl + 1]; a += 32) {
        const int j = mcol[a];
        const float xj = x[j];
#pragma unroll
        for (int q = 0; q < 6; ++q) sacc[q] += cdof[j * 6 + q] * xj;
      }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) sacc[q] = warp_sum(sacc[q]);
    if (LANE < 4) {
      const float* fr = c.s + c.L.cframe + 9 * k;
      V3 vel = v3(sacc[3], sacc[4], sacc[5]) + cross(v3(sacc[0], sacc[1], sacc[2]), ld3(c.s + c.L.crel + 3 * k));
      const float un = dot(ld3(fr), vel);
      const float ut = dot(ld3(fr + 3 * (1 + (LANE >> 1))), vel);
      const float mu = c.s[c.L.cmu + k];
      out[nl + 4 * k + LANE] = un + ut * ((LANE & 1) ? -mu : mu);
    }
  }
  for (int r = TID; r < nl; r += c.nt) out[r] = lim_sign[r] * x[lim_dof[r]];
}

// qfrc_con = J^T f with f[row] = -D Jaref [Jaref < 0]; returns nothing, needs a sync after.
__device__ __noinline__ void jtmul_force(Ctx& c, float* qfrc) {
  const int* ints = (const int*)(c.s + c.L.ints);
  const int nl = ints[0], nc = ints[1];
  const float* D = c.s + c.L.efcD;
  const float* Jaref = c.s + c.L.Jaref;
  for (int k = TID; k < nc; k += c.nt) {
    float f[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float ja = Jaref[nl + 4 * k + q];
      f[q] = (ja < 0.0f) ? -D[nl + 4 * k + q] * ja : 0.0f;
    }
    const float mu = c.s[c.L.cmu + k];
    const float fn = f[0] + f[1] + f[2] + f[3], f1 = mu * (f[0] - f[1]), f2 = mu * (f[2] - f[3]);
    const float* fr = c.s + c.L.cframe + 9 * k;
    V3 F = ld3(fr) * fn + ld3(fr + 3) * f1 + ld3(fr + 6) * f2;
    V3 tq = cross(ld3(c.s + c.L.crel + 3 * k), F);
    st3(c.s + c.L.cwrench + 6 * k, tq);
    st3(c.s + c.L.cwrench + 6 * k + 3, F);
  }
  __syncthreads();
  const int* dof_body = c.fi(VNL_F_DOF_BODYID);
  const int* sub_end = c.fi(VNL_F_BODY_SUBTREE_END);
  const int* cbody = (const int*)(c.s + c.L.cbody);
  const int* limrow = (const int*)(c.s + c.L.limrow_of_dof);
  const float* lim_sign = c.s + c.L.lim_sign;
  for (int i = TID; i < c.d.nv; i += c.nt) {
    const int b = dof_body[i], be = sub_end[b];
    const float* cd = c.s + c.L.cdof + 6 * i;
    float acc = 0.0f;
    for (int k = 0; k < nc; ++k) {
      const int cb = cbody[k];
      if (cb >= b && cb < be) acc += dot6(cd, c.s + c.L.cwrench + 6 * k);
    }
    const int r = limrow[i];
    if (r >= 0) {
      const float ja = Jaref[r];
      if (ja < 0.0f) acc += lim_sign[r] * (-D[r] * ja);
    }
    qfrc[i] = acc;
  }
}

// constraint._kbi
__device__ __forceinline__ void kbi(const Dims& d, float sr0, float sr1, const float* solimp, float pos, float& k, float& b, float& imp) {
  const float timeconst = fmaxf(sr0, 2.0f * d.timestep), dampratio = sr1;
  const float dmin = fminf(fmaxf(solimp[0], VNL_MINIMP), VNL_MAXIMP), dmax = fminf(fmaxf(solimp[1], VNL_MINIMP), VNL_MAXIMP);
  const float width = fmaxf(VNL_MINVAL, solimp[2]), mid = fminf(fmaxf(solimp[3], VNL_MINIMP), VNL_MAXIMP), power = fmaxf(1.0f, solimp[4]);
  k = 1.0f / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  b = 2.0f / (dmax * timeconst);
  if (sr0 <= 0.0f) k = -sr0 / (dmax * dmax);
  if (sr1 <= 0.0f) b = -sr1 / dmax;
  const float imp_x = fabsf(pos) / width;
  const float imp_a = (1.0f / powf(mid, power - 1.0f)) * powf(imp_x, power);
  const float imp_b = 1.0f - (1.0f / powf(1.0f - mid, power - 1.0f)) * powf(1.0f - imp_x, power);
  const float imp_y = imp_x < mid ? imp_a : imp_b;
  imp = dmin + imp_y * (dmax - dmin);
  imp = fminf(fmaxf(imp, dmin), dmax);
  if (imp_x > 1.0f) imp = dmax;
}

// ---------------------------------------------------------------------------------------------
// mjx.forward for the env held in shared memory
// ---------------------------------------------------------------------------------------------
struct LSP { float alpha, cost, d0, d1; };

__device__ __noinline__ void forward(Ctx& c, int* stats, float* dump) {
  const Dims& d = c.d;
  const Lay& L = c.L;
  float* s = c.s;
  const int tid = TID, nt = c.nt;
  int* ints = (int*)(s + L.ints);

  // ---- smooth.kinematics: level-synchronous walk down the body tree --------------------------
  {
    const int* lstart = c.fi(VNL_F_LEVEL_START);
    const int* lbody = c.fi(VNL_F_LEVEL_BODY);
    const int* parent = c.fi(VNL_F_BODY_PARENTID);
    const int* jntadr = c.fi(VNL_F_BODY_JNTADR);
    const int* jntnum = c.fi(VNL_F_BODY_JNTNUM);
    const int* jtype = c.fi(VNL_F_JNT_TYPE);
    const int* jqadr = c.fi(VNL_F_JNT_QPOSADR);
    const float* bpos = c.ff(VNL_F_BODY_POS);
    const float* bquat = c.ff(VNL_F_BODY_QUAT);
    const float* jpos = c.ff(VNL_F_JNT_POS);
    const float* jaxis = c.ff(VNL_F_JNT_AXIS);
    const float* qpos0 = c.ff(VNL_F_QPOS0);
    if (tid == 0) {
      s[L.xpos] = s[L.xpos + 1] = s[L.xpos + 2] = 0.0f;
      s[L.xquat] = 1.0f; s[L.xquat + 1] = s[L.xquat + 2] = s[L.xquat + 3] = 0.0f;
    }
    __syncthreads();
    for (int lv = 0; lv < d.nlevel; ++lv) {
      for (int k = lstart[lv] + tid; k < lstart[lv + 1]; k += nt) {
        const int b = lbody[k], p = parent[b];
        Q4 pq = ld4(s + L.xquat + 4 * p);
        V3 pos = ld3(s + L.xpos + 3 * p) + rotate(ld3(bpos + 3 * b), pq);
        Q4 quat = quat_mul(pq, ld4(bquat + 4 * b));
        for (int jj = 0; jj < jntnum[b]; ++jj) {
          const int j = jntadr[b] + jj, qa = jqadr[j];
          if (jtype[j] == 0) {
            pos = ld3(s + L.qpos + qa);
            st3(s + L.xanchor + 3 * j, pos);
            st3(s + L.xaxis + 3 * j, v3(0.0f, 0.0f, 1.0f));
            quat = quat_normalize(ld4(s + L.qpos + qa + 3));
            st4(s + L.qpos + qa + 3, quat);
          } else {
            V3 jp = ld3(jpos + 3 * j), ja = ld3(jaxis + 3 * j);
            V3 anchor = rotate(jp, quat) + pos;
            st3(s + L.xanchor + 3 * j, anchor);
            st3(s + L.xaxis + 3 * j, rotate(ja, quat));
            quat = quat_mul(quat, axis_angle_quat(ja, s[L.qpos + qa] - qpos0[qa]));
            pos = anchor - rotate(jp, quat);
          }
        }
        st3(s + L.xpos + 3 * b, pos);
        st4(s + L.xquat + 4 * b, quat);
      }
      __syncthreads();
    }
  }
  PROF(c, 0);
  // ---- smooth.com_pos: xipos, root COM, cinert, cdof ------------------------------------------
  {
    const float* ipos = c.ff(VNL_F_BODY_IPOS);
    for (int b = tid; b < d.nbody; b += nt)
      st3(s + L.xipos + 3 * b, ld3(s + L.xpos + 3 * b) + rotate(ld3(ipos + 3 * b), ld4(s + L.xquat + 4 * b)));
    __syncthreads();
    const int* rootid = c.fi(VNL_F_BODY_ROOTID);
    const int* sub_end = c.fi(VNL_F_BODY_SUBTREE_END);
    const float* mass = c.ff(VNL_F_BODY_MASS);
    // one warp per kinematic tree: subtree COM of the root = mass-weighted mean over its id range
    for (int b = 1 + WARP; b < d.nbody; b += c.nw) {
      if (rootid[b] != b) continue;
      float acc[4] = {0, 0, 0, 0};
      for (int q = b + LANE; q < sub_end[b]; q += 32) {
        const float mq = mass[q];
        acc[0] += s[L.xipos + 3 * q] * mq; acc[1] += s[L.xipos + 3 * q + 1] * mq; acc[2] += s[L.xipos + 3 * q + 2] * mq; acc[3] += mq;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = warp_sum(acc[q]);
      if (LANE < 3) {
        const float a = LANE == 0 ? acc[0] : (LANE == 1 ? acc[1] : acc[2]);
        s[L.rcom + 3 * b + LANE] = (acc[3] < VNL_MINVAL) ? s[L.xipos + 3 * b + LANE] : a / fmaxf(acc[3], VNL_MINVAL);
      }
    }
    __syncthreads();
    const float* iquat = c.ff(VNL_F_BODY_IQUAT);
    const float* inertia = c.ff(VNL_F_BODY_INERTIA);
    for (int b = tid; b < d.nbody; b += nt) {
      float* ci = s + L.cinert + 10 * b;
      if (b == 0) {
        for (int q = 0; q < 10; ++q) ci[q] = 0.0f;
        continue;
      }
      float R[9];
      quat_to_mat(quat_mul(ld4(s + L.xquat + 4 * b), ld4(iquat + 4 * b)), R);
      V3 off = ld3(s + L.xipos + 3 * b) - ld3(s + L.rcom + 3 * rootid[b]);
      const float I0 = inertia[3 * b], I1 = inertia[3 * b + 1], I2 = inertia[3 * b + 2], ms = mass[b];
      const float oo = dot(off, off);
      const float o[3] = {off.x, off.y, off.z};
      // (ximat * inertia) @ ximat.T + h @ h.T * mass, h = cross(off, -I): entries [00 11 22 01 02 12]
      const int rr[6] = {0, 1, 2, 0, 0, 1}, cc[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const int r = rr[q], cl = cc[q];
        float v = R[3 * r] * I0 * R[3 * cl] + R[3 * r + 1] * I1 * R[3 * cl + 1] + R[3 * r + 2] * I2 * R[3 * cl + 2];
        v += (((r == cl) ? oo : 0.0f) - o[r] * o[cl]) * ms;
        ci[q] = v;
      }
      ci[6] = off.x * ms; ci[7] = off.y * ms; ci[8] = off.z * ms; ci[9] = ms;
    }
    const int* jtype = c.fi(VNL_F_JNT_TYPE);
    const int* jbody = c.fi(VNL_F_JNT_BODYID);
    const int* jdof = c.fi(VNL_F_JNT_DOFADR);
    for (int j = tid; j < d.njnt; j += nt) {
      const int b = jbody[j], da = jdof[j];
      V3 off = ld3(s + L.rcom + 3 * rootid[b]) - ld3(s + L.xanchor + 3 * j);
      if (jtype[j] == 0) {
        float R[9];
        quat_to_mat(ld4(s + L.xquat + 4 * b), R);
        for (int a = 0; a < 3; ++a) {
          float* ct = s + L.cdof + 6 * (da + a);
          for (int q = 0; q < 6; ++q) ct[q] = 0.0f;
          ct[3 + a] = 1.0f;
          V3 ax = v3(R[a], R[3 + a], R[6 + a]);
          st3(s + L.cdof + 6 * (da + 3 + a), ax);
          st3(s + L.cdof + 6 * (da + 3 + a) + 3, cross(ax, off));
        }
      } else {
        V3 ax = ld3(s + L.xaxis + 3 * j);
        st3(s + L.cdof + 6 * da, ax);
        st3(s + L.cdof + 6 * da + 3, cross(ax, off));
      }
    }
    __syncthreads();
  }
  PROF(c, 1);
  // ---- composite inertia (flat subtree sums), cvel chains ----------------------------------------
  const uint16_t* madr = c.madr; const uint8_t* mcol = c.mcol; const uint8_t* mrow = c.mrow;
  const int* lastdof = c.fi(VNL_F_BODY_LASTDOF);
  const int* sub_end = c.fi(VNL_F_BODY_SUBTREE_END);
  const int* dof_body = c.fi(VNL_F_DOF_BODYID);
  const int* dof_jnt = c.fi(VNL_F_DOF_JNTID);
  const int* jtype = c.fi(VNL_F_JNT_TYPE);
  const int* jdofadr = c.fi(VNL_F_JNT_DOFADR);
  {
    for (int it = tid; it < d.nbody * 10; it += nt) {
      const int b = it / 10, q = it - 10 * b;
      float acc = 0.0f;
      if (b > 0)
        for (int k = b; k < sub_end[b]; ++k) acc += s[L.cinert + 10 * k + q];
      s[L.crb + it] = acc;
    }
    // velocity of the chain BEFORE each dof (smooth.com_vel): ancestors only; the three rotational
    // dofs of a free joint all see the velocity after its translational dofs.
    for (int it = tid; it < d.nv * 6; it += nt) {
      const int i = it / 6, q = it - 6 * i;
      const int j = dof_jnt[i];
      const bool freej = jtype[j] == 0;
      const int fa = jdofadr[j];
      float acc = 0.0f;
      for (int a = madr[i + 1] - 1; a > madr[i]; --a) {  // root-most ancestor first
        const int k = mcol[a];
        if (freej && k >= fa + 3) continue;
        acc += s[L.cdof + 6 * k + q] * s[L.qvel + k];
      }
      s[L.cdofdot + it] = acc;
    }
    for (int it = tid; it < d.nbody * 6; it += nt) {
      const int b = it / 6, q = it - 6 * b;
      const int dl = lastdof[b];
      float acc = 0.0f;
      if (dl >= 0)
        for (int a = madr[dl + 1] - 1; a >= madr[dl]; --a) { const int k = mcol[a]; acc += s[L.cdof + 6 * k + q] * s[L.qvel + k]; }
      s[L.cvel + it] = acc;
    }
    __syncthreads();
    for (int i = tid; i < d.nv; i += nt) {  // cdof_dot = motion_cross(cvel_before, cdof); 0 for free translations
      float* cd = s + L.cdofdot + 6 * i;
      const int j = dof_jnt[i];
      if (jtype[j] == 0 && i < jdofadr[j] + 3) {
        for (int q = 0; q < 6; ++q) cd[q] = 0.0f;
      } else {
        float v[6], r[6];
        for (int q = 0; q < 6; ++q) v[q] = cd[q];
        motion_cross(v, s + L.cdof + 6 * i, r);
        for (int q = 0; q < 6; ++q) cd[q] = r[q];
      }
    }
    __syncthreads();
  }
  PROF(c, 2);
  // ---- smooth.rne: cacc chains, local cfrc, subtree sums, qfrc_bias (kept in tmpv) -------------------
  {
    for (int it = tid; it < d.nbody * 6; it += nt) {
      const int b = it / 6, q = it - 6 * b;
      const int dl = lastdof[b];
      float acc = (q == 3) ? -d.gx : ((q == 4) ? -d.gy : ((q == 5) ? -d.gz : 0.0f));
      if (dl >= 0)
        for (int a = madr[dl + 1] - 1; a >= madr[dl]; --a) { const int k = mcol[a]; acc += s[L.cdofdot + 6 * k + q] * s[L.qvel + k]; }
      s[L.cacc + it] = acc;
    }
    __syncthreads();
    for (int b = tid; b < d.nbody; b += nt) {
      float f1[6], iv[6], f2[6];
      inert_mul(s + L.cinert + 10 * b, s + L.cacc + 6 * b, f1);
      inert_mul(s + L.cinert + 10 * b, s + L.cvel + 6 * b, iv);
      motion_cross_force(s + L.cvel + 6 * b, iv, f2);
      for (int q = 0; q < 6; ++q) s[L.cfrc + 6 * b + q] = f1[q] + f2[q];
    }
    __syncthreads();
    for (int it = tid; it < d.nbody * 6; it += nt) {  // cacc <- subtree-summed cfrc
      const int b = it / 6, q = it - 6 * b;
      float acc = 0.0f;
      for (int k = b; k < sub_end[b]; ++k) acc += s[L.cfrc + 6 * k + q];
      s[L.cacc + it] = acc;
    }
    __syncthreads();
  }
  PROF(c, 3);
  // ---- qfrc_smooth = passive - bias + actuator; act_dot ---------------------------------------------
  {
    const int* jqadr = c.fi(VNL_F_JNT_QPOSADR);
    const float* stiff = c.ff(VNL_F_JNT_STIFFNESS);
    const float* qspring = c.ff(VNL_F_QPOS_SPRING);
    const float* damping = c.ff(VNL_F_DOF_DAMPING);
    for (int i = tid; i < d.nv; i += nt) s[L.qfrc_act + i] = 0.0f;
    __syncthreads();
    const int* adof = c.fi(VNL_F_ACT_DOFADR);
    const int* aadr = c.fi(VNL_F_ACT_ACTADR);
    const int* flim = c.fi(VNL_F_ACT_FORCELIMITED);
    const float* gain = c.ff(VNL_F_ACT_GAIN);
    const float* gear = c.ff(VNL_F_ACT_GEAR);
    const float* frange = c.ff(VNL_F_ACT_FORCERANGE);
    const float* dynprm = c.ff(VNL_F_ACT_DYNPRM);
    // one thread per dof gathers its actuators in actuator order (deterministic, no atomics)
    for (int i = tid; i < d.nv; i += nt) {
      float acc = 0.0f;
      for (int u = 0; u < d.nu; ++u) {
        if (adof[u] != i) continue;
        const float ctrl = s[L.ctrl + u];
        float ca = ctrl;
        const int aa = aadr[u];
        if (aa >= 0) {
          s[L.act_dot + aa] = (ctrl - s[L.act + aa]) / fmaxf(dynprm[u], VNL_MINVAL);
          ca = s[L.act + aa];
        }
        float force = gain[u] * ca;
        if (flim[u]) force = fminf(fmaxf(force, frange[2 * u]), frange[2 * u + 1]);
        acc += gear[u] * force;
      }
      s[L.qfrc_act + i] = acc;
      // passive
      const int j = dof_jnt[i];
      float pas;
      if (jtype[j] == 0) {
        const int a = i - jdofadr[j];
        pas = (a < 3) ? -stiff[j] * (s[L.qpos + jqadr[j] + a] - qspring[jqadr[j] + a]) : 0.0f;
      } else {
        pas = -stiff[j] * (s[L.qpos + jqadr[j]] - qspring[jqadr[j]]);
      }
      pas -= damping[i] * s[L.qvel + i];
      const float bias = dot6(s + L.cdof + 6 * i, s + L.cacc + 6 * dof_body[i]);
      s[L.qfrc_smooth + i] = pas - bias + acc;
      if (dump) { dump[c.d.dump_passive + i] = pas; dump[c.d.dump_passive + d.nv + i] = bias; }
    }
    __syncthreads();
  }
  PROF(c, 4);
  // ---- joint-space inertia (tree sparse): M[i][a] = cdof[anc_a(i)] . (crb[body_i] cdof_i) -------
  {
    const float* armature = c.ff(VNL_F_DOF_ARMATURE);
    float* fd = s + L.cdofdot;  // cdof_dot is dead from here on: reuse as crb * cdof
    for (int i = tid; i < d.nv; i += nt) inert_mul(s + L.crb + 10 * dof_body[i], s + L.cdof + 6 * i, fd + 6 * i);
    __syncthreads();
    for (int e = tid; e < d.nM; e += nt) {
      const int i = mrow[e], j = mcol[e];
      float v = dot6(fd + 6 * i, s + L.cdof + 6 * j);
      if (i == j) v += armature[i];
      s[L.M + e] = v;
    }
    __syncthreads();
  }
  PROF(c, 5);
  factor(c, s + L.M, false);
  PROF(c, 6);
  solve_m(c, s + L.qfrc_smooth, s + L.qacc_smooth, s + L.tmpv);
  PROF(c, 7);

  // ---- collision + constraint rows, compacted to the active set -----------------------------------
  {
    if (tid == 0) { ints[0] = 0; ints[1] = 0; }
    for (int i = tid; i < d.nv; i += nt) ((int*)(s + L.limrow_of_dof))[i] = -1;
    __syncthreads();
    if (WARP == 0) {
      // joint limits (constraint._instantiate_limit_slide_hinge)
      const int* ljnt = c.fi(VNL_F_LIMIT_JNT);
      const int* jqadr = c.fi(VNL_F_JNT_QPOSADR);
      const float* range = c.ff(VNL_F_JNT_RANGE);
      const float* margin = c.ff(VNL_F_JNT_MARGIN);
      const float* solref = c.ff(VNL_F_JNT_SOLREF);
      const float* solimp = c.ff(VNL_F_JNT_SOLIMP);
      const float* invw = c.ff(VNL_F_DOF_INVWEIGHT0);
      int base = 0;
      for (int r0 = 0; r0 < d.nlimit; r0 += 32) {
        const int r = r0 + LANE;
        bool active = false;
        float pos = 0.0f, sign = 0.0f;
        int j = 0;
        if (r < d.nlimit) {
          j = ljnt[r];
          const float q = s[L.qpos + jqadr[j]];
          const float dmin = q - range[2 * j], dmax = range[2 * j + 1] - q;
          pos = fminf(dmin, dmax) - margin[j];
          active = pos < 0.0f;
          sign = (dmin < dmax) ? 1.0f : -1.0f;
        }
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (active) {
          const int slot = base + __popc(m & ((1u << LANE) - 1u));
          const int dof = jdofadr[j];
          ((int*)(s + L.lim_dof))[slot] = dof;
          s[L.lim_sign + slot] = sign;
          ((int*)(s + L.limrow_of_dof))[dof] = slot;
          float k, b, imp;
          kbi(d, solref[2 * j], solref[2 * j + 1], solimp + 5 * j, pos, k, b, imp);
          const float R = fmaxf(invw[dof] * (1.0f - imp) / imp, VNL_MINVAL);
          s[L.efcD + slot] = 1.0f / R;
          s[L.aref + slot] = -b * (sign * s[L.qvel + dof]) - k * imp * pos;
          if (dump) { dump[d.dump_efc + r] = pos; dump[d.dump_efc + d.nefc + r] = 1.0f / R; dump[d.dump_efc + 2 * d.nefc + r] = s[L.aref + slot]; }
        }
        base += __popc(m);
      }
      if (LANE == 0) ints[0] = base;
    }
    __syncthreads();
    if (WARP == 0) {
      // contacts (collision_driver + constraint._instantiate_contact, pyramidal condim 3)
      const int nl = ints[0];
      const int* cpair = c.fi(VNL_F_CON_PAIR);
      const float* csign = c.ff(VNL_F_CON_SIGN);
      const int* ptype = c.fi(VNL_F_PAIR_TYPE);
      const int* g1 = c.fi(VNL_F_PAIR_GEOM1);
      const int* g2 = c.fi(VNL_F_PAIR_GEOM2);
      const int* gbody = c.fi(VNL_F_GEOM_BODYID);
      const int* rootid = c.fi(VNL_F_BODY_ROOTID);
      const float* gpos = c.ff(VNL_F_GEOM_POS);
      const float* gquat = c.ff(VNL_F_GEOM_QUAT);
      const float* gsize = c.ff(VNL_F_GEOM_SIZE);
      const float* pfric = c.ff(VNL_F_PAIR_FRICTION);
      const float* psolref = c.ff(VNL_F_PAIR_SOLREF);
      const float* psolimp = c.ff(VNL_F_PAIR_SOLIMP);
      const float* pmargin = c.ff(VNL_F_PAIR_INCLUDEMARGIN);
      const float* binvw = c.ff(VNL_F_BODY_INVWEIGHT0);
      int base = 0;
      for (int c0 = 0; c0 < d.ncon; c0 += 32) {
        const int ci = c0 + LANE;
        bool active = false;
        float dist = 0.0f;
        V3 cp = v3(0, 0, 0), n = v3(0, 0, 1), fb = v3(0, 1, 0);
        int p = 0, body = 0;
        if (ci < d.ncon) {
          p = cpair[ci];
          const int ga = g1[p], gb = g2[p];
          body = gbody[gb];
          // plane is attached to the world body: world frame = local frame
          float Pm[9], Gm[9];
          quat_to_mat(ld4(gquat + 4 * ga), Pm);
          n = v3(Pm[2], Pm[5], Pm[8]);
          const V3 ppos = ld3(gpos + 3 * ga);
          const Q4 bq = ld4(s + L.xquat + 4 * body);
          const V3 gp = ld3(s + L.xpos + 3 * body) + rotate(ld3(gpos + 3 * gb), bq);
          quat_to_mat(quat_mul(bq, ld4(gquat + 4 * gb)), Gm);
          const float* sz = gsize + 3 * gb;
          const int ty = ptype[p];
          if (ty == 2) {  // plane_sphere
            dist = dot(gp - ppos, n) - sz[0];
            cp = gp - n * (sz[0] + 0.5f * dist);
          } else if (ty == 3) {  // plane_capsule end
            V3 axis = v3(Gm[2], Gm[5], Gm[8]);
            V3 b = axis - n * dot(n, axis);
            const float bn = normalize3(b);
            if (bn < 0.5f) b = (-0.5f < n.y && n.y < 0.5f) ? v3(0, 1, 0) : v3(0, 0, 1);
            fb = b;
            V3 sp = gp + axis * (csign[ci] * sz[1]);
            dist = dot(sp - ppos, n) - sz[0];
            cp = sp - n * (sz[0] + 0.5f * dist);
          } else {  // plane_ellipsoid
            V3 ln = v3(Gm[0] * n.x + Gm[3] * n.y + Gm[6] * n.z, Gm[1] * n.x + Gm[4] * n.y + Gm[7] * n.z, Gm[2] * n.x + Gm[5] * n.y + Gm[8] * n.z);
            V3 sup = v3(ln.x * sz[0], ln.y * sz[1], ln.z * sz[2]);
            normalize3(sup);
            sup = v3(-sup.x * sz[0], -sup.y * sz[1], -sup.z * sz[2]);
            V3 pt = gp + v3(Gm[0] * sup.x + Gm[1] * sup.y + Gm[2] * sup.z, Gm[3] * sup.x + Gm[4] * sup.y + Gm[5] * sup.z, Gm[6] * sup.x + Gm[7] * sup.y + Gm[8] * sup.z);
            dist = dot(n, pt - ppos);
            cp = pt - n * (dist * 0.5f);
          }
          if (ty != 3) {  // math.make_frame
            V3 nn = n;
            normalize3(nn);
            n = nn;
            V3 b = (-0.5f < n.y && n.y < 0.5f) ? v3(0, 1, 0) : v3(0, 0, 1);
            b = b - n * dot(n, b);
            normalize3(b);
            fb = b;
          }
          if (dump) {
            dump[d.dump_con + ci] = dist;
            st3(dump + d.dump_con + d.ncon + 3 * ci, cp);
            st3(dump + d.dump_con + 4 * d.ncon + 9 * ci, n); st3(dump + d.dump_con + 4 * d.ncon + 9 * ci + 3, fb);
            st3(dump + d.dump_con + 4 * d.ncon + 9 * ci + 6, cross(n, fb));
          }
          dist -= pmargin[p];
          active = dist < 0.0f;
        }
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (active) {
          const int k = base + __popc(m & ((1u << LANE) - 1u));
          ((int*)(s + L.cbody))[k] = body;
          const V3 rel = cp - ld3(s + L.rcom + 3 * rootid[body]);
          st3(s + L.crel + 3 * k, rel);
          const V3 t2 = cross(n, fb);
          st3(s + L.cframe + 9 * k, n); st3(s + L.cframe + 9 * k + 3, fb); st3(s + L.cframe + 9 * k + 6, t2);
          const float mu = pfric[5 * p];
          s[L.cmu + k] = mu;
          float kk, bb, imp;
          kbi(d, psolref[2 * p], psolref[2 * p + 1], psolimp + 5 * p, dist, kk, bb, imp);
          const float t = binvw[2 * body];
          const float invweight = (t + mu * mu * t) * 2.0f * mu * mu / d.impratio;
          const float R = fmaxf(invweight * (1.0f - imp) / imp, VNL_MINVAL);
          // J qvel for the four pyramid rows = frame . point velocity (cvel is the body's spatial velocity)
          const float* cv = s + L.cvel + 6 * body;
          const V3 vel = ld3(cv + 3) + cross(ld3(cv), rel);
          const float un = dot(n, vel), u1 = dot(fb, vel), u2 = dot(t2, vel);
          const float ref = -kk * imp * dist;
          const int r = nl + 4 * k;
          s[L.efcD + r] = s[L.efcD + r + 1] = s[L.efcD + r + 2] = s[L.efcD + r + 3] = 1.0f / R;
          s[L.aref + r] = -bb * (un + mu * u1) + ref;
          s[L.aref + r + 1] = -bb * (un - mu * u1) + ref;
          s[L.aref + r + 2] = -bb * (un + mu * u2) + ref;
          s[L.aref + r + 3] = -bb * (un - mu * u2) + ref;
          if (dump) {
            const int rr = d.nlimit + 4 * ci;
            for (int q = 0; q < 4; ++q) {
              dump[d.dump_efc + rr + q] = dist; dump[d.dump_efc + d.nefc + rr + q] = 1.0f / R;
              dump[d.dump_efc + 2 * d.nefc + rr + q] = s[L.aref + r + q];
            }
          }
        }
        base += __popc(m);
      }
      if (LANE == 0) ints[1] = base;
    }
    __syncthreads();
  }
  PROF(c, 8);
  const int nl = ints[0], nc = ints[1], nrow = nl + 4 * nc;
  if (stats && tid == 0) { stats[2] += nc; stats[3] += nl; }

  // ---- solver.solve (CG with the MJX line search) ----------------------------------------------------
  float* qacc = s + L.qacc;
  float* Ma = s + L.Ma;
  float* Jaref = s + L.Jaref;
  float* Jv = s + L.Jv;
  float* efcD = s + L.efcD;
  float* grad = s + L.grad;
  float* Mgrad = s + L.Mgrad;
  float* search = s + L.search;
  float* Mv = s + L.Mv;
  float* qfrc_con = s + L.qfrc_con;
  const float* qfrc_smooth = s + L.qfrc_smooth;
  const float* qacc_smooth = s + L.qacc_smooth;
  int niter = 0, lsiter = 0;
  {
    // cost of a candidate: 0.5 sum D Jaref^2 [Jaref<0] + 0.5 (Ma - qfrc_smooth).(qacc - qacc_smooth)
    float cost_w, cost_s;
    {
      mul_m(c, s + L.warm, Ma);
      jmul(c, s + L.warm, Jaref);
      __syncthreads();
      float v[2] = {0.0f, 0.0f};
      for (int r = tid; r < nrow; r += nt) { const float ja = Jaref[r] - s[L.aref + r]; if (ja < 0.0f) v[0] += efcD[r] * ja * ja; }
      for (int i = tid; i < d.nv; i += nt) v[1] += (Ma[i] - qfrc_smooth[i]) * (s[L.warm + i] - qacc_smooth[i]);
      block_sum<2>(c, v);
      cost_w = 0.5f * v[0] + 0.5f * v[1];
      __syncthreads();
      mul_m(c, qacc_smooth, Ma);
      jmul(c, qacc_smooth, Jaref);
      __syncthreads();
      float w[2] = {0.0f, 0.0f};
      for (int r = tid; r < nrow; r += nt) { const float ja = Jaref[r] - s[L.aref + r]; if (ja < 0.0f) w[0] += efcD[r] * ja * ja; }
      for (int i = tid; i < d.nv; i += nt) w[1] += (Ma[i] - qfrc_smooth[i]) * (qacc_smooth[i] - qacc_smooth[i]);
      block_sum<2>(c, w);
      cost_s = 0.5f * w[0] + 0.5f * w[1];
      __syncthreads();
    }
    PROF(c, 9);
    const bool use_warm = cost_w < cost_s;
    if (use_warm) {
      for (int i = tid; i < d.nv; i += nt) qacc[i] = s[L.warm + i];
      __syncthreads();
      mul_m(c, qacc, Ma);
      jmul(c, qacc, Jaref);
    } else {
      for (int i = tid; i < d.nv; i += nt) qacc[i] = qacc_smooth[i];  // Ma, Jaref already hold the smooth candidate
    }
    __syncthreads();
    for (int r = tid; r < nrow; r += nt) Jaref[r] -= s[L.aref + r];
    __syncthreads();
    const float scale = d.meaninertia * (float)max(1, d.nv);
    float cost = INFINITY, prev_cost = 0.0f, gauss = 0.0f, gradnorm = 0.0f;
    // _update_constraint + _update_gradient
    auto update = [&]() {
      jtmul_force(c, qfrc_con);
      float v[2] = {0.0f, 0.0f};
      for (int r = tid; r < nrow; r += nt) { const float ja = Jaref[r]; if (ja < 0.0f) v[0] += efcD[r] * ja * ja; }
      for (int i = tid; i < d.nv; i += nt) v[1] += (Ma[i] - qfrc_smooth[i]) * (qacc[i] - qacc_smooth[i]);
      block_sum<2>(c, v);  // also orders qfrc_con writes before the reads below
      gauss = 0.5f * v[1];
      prev_cost = cost;
      cost = 0.5f * v[0] + gauss;
      float g[1] = {0.0f};
      for (int i = tid; i < d.nv; i += nt) { const float gi = Ma[i] - qfrc_smooth[i] - qfrc_con[i]; grad[i] = gi; g[0] += gi * gi; }
      block_sum<1>(c, g);
      gradnorm = sqrtf(g[0]);
      solve_m(c, grad, Mgrad, s + L.tmpv);
    };
    update();
    PROF(c, 10);
    for (int i = tid; i < d.nv; i += nt) search[i] = -Mgrad[i];
    __syncthreads();
    while (true) {
      const float improvement = (prev_cost - cost) / scale;
      const float gradient = gradnorm / scale;
      bool done = niter >= d.iterations;
      if (d.iterations != 1) { done |= improvement < d.tolerance; done |= gradient < d.tolerance; }
      if (done) break;
      // ---- _linesearch ----
      mul_m(c, search, Mv);
      jmul(c, search, Jv);
      __syncthreads();
      float qg[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int i = tid; i < d.nv; i += nt) {
        const float si = search[i];
        qg[0] += si * si; qg[1] += si * Ma[i]; qg[2] += si * qfrc_smooth[i]; qg[3] += si * Mv[i];
      }
      block_sum<4>(c, qg);
      const float gtol = d.tolerance * d.ls_tolerance * (sqrtf(qg[0]) * scale);
      const float g0 = gauss, g1 = qg[1] - qg[2], g2 = 0.5f * qg[3];
      auto points = [&](int n, const float* alpha, LSP* out) {
        float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int r = tid; r < nrow; r += nt) {
          const float x = Jaref[r], jv = Jv[r], D = efcD[r];
          const float q0 = 0.5f * x * x * D, q1 = jv * x * D, q2 = 0.5f * jv * jv * D;
#pragma unroll
          for (int k = 0; k < 3; ++k)
            if (k < n && x + alpha[k] * jv < 0.0f) { acc[3 * k] += q0; acc[3 * k + 1] += q1; acc[3 * k + 2] += q2; }
        }
        block_sum<9>(c, acc);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (k < n) {
            const float q0 = acc[3 * k] + g0, q1 = acc[3 * k + 1] + g1, q2 = acc[3 * k + 2] + g2, a = alpha[k];
            out[k].alpha = a;
            out[k].cost = a * a * q2 + a * q1 + q0;
            out[k].d0 = 2.0f * a * q2 + q1;
            out[k].d1 = 2.0f * q2 + (q2 == 0.0f ? VNL_MINVAL : 0.0f);
          }
      };
      LSP p0, lo, hi, tmp3[3];
      float al[3] = {0.0f, 0.0f, 0.0f};
      points(1, al, tmp3);
      p0 = tmp3[0];
      al[0] = p0.alpha - p0.d0 / p0.d1;
      points(1, al, tmp3);
      lo = tmp3[0];
      const bool lesser = lo.d0 < p0.d0;
      hi = lesser ? p0 : lo;
      lo = lesser ? lo : p0;
      bool swap = true;
      int it = 0;
      while (true) {
        bool ldone = it >= d.ls_iterations;
        ldone |= !swap;
        ldone |= (lo.d0 < 0.0f) && (lo.d0 > -gtol);
        ldone |= (hi.d0 > 0.0f) && (hi.d0 < gtol);
        if (ldone) break;
        al[0] = lo.alpha - lo.d0 / lo.d1;
        al[1] = hi.alpha - hi.d0 / hi.d1;
        al[2] = 0.5f * (lo.alpha + hi.alpha);
        points(3, al, tmp3);
        const LSP lo_next = tmp3[0], hi_next = tmp3[1], mid = tmp3[2];
        const bool s1 = (lo.d0 > 0.0f) || (lo.d0 < lo_next.d0);
        if (s1) lo = lo_next;
        const bool s2 = (mid.d0 < 0.0f) && (lo.d0 < mid.d0);
        if (s2) lo = mid;
        const bool s3 = (hi_next.d0 < 0.0f) && (lo.d0 < hi_next.d0);
        if (s3) lo = hi_next;
        const bool s4 = (hi.d0 < 0.0f) || (hi.d0 > hi_next.d0);
        if (s4) hi = hi_next;
        const bool s5 = (mid.d0 > 0.0f) && (hi.d0 > mid.d0);
        if (s5) hi = mid;
        const bool s6 = (lo_next.d0 > 0.0f) && (hi.d0 > lo_next.d0);
        if (s6) hi = lo_next;
        swap = s1 || s2 || s3 || s4 || s5 || s6;
        ++it;
      }
      lsiter += it;
      PROF(c, 11);
      const bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
      const float alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
      const float ia = improved ? alpha : 0.0f * alpha;
      for (int i = tid; i < d.nv; i += nt) { qacc[i] += search[i] * ia; Ma[i] += Mv[i] * ia; }
      for (int r = tid; r < nrow; r += nt) Jaref[r] += Jv[r] * ia;
      // previous grad . Mgrad before they are overwritten
      float pg[1] = {0.0f};
      for (int i = tid; i < d.nv; i += nt) { pg[0] += grad[i] * Mgrad[i]; Mv[i] = Mgrad[i]; }  // Mv <- previous Mgrad
      block_sum<1>(c, pg);
      update();
      if (d.solver == 2) {
        for (int i = tid; i < d.nv; i += nt) search[i] = -Mgrad[i];
      } else {  // Polak-Ribiere
        float nb[1] = {0.0f};
        for (int i = tid; i < d.nv; i += nt) nb[0] += grad[i] * (Mgrad[i] - Mv[i]);
        block_sum<1>(c, nb);
        const float beta = fmaxf(0.0f, nb[0] / fmaxf(VNL_MINVAL, pg[0]));
        for (int i = tid; i < d.nv; i += nt) search[i] = -Mgrad[i] + beta * search[i];
      }
      __syncthreads();
      PROF(c, 12);
      ++niter;
    }
  }
  if (stats && tid == 0) { stats[0] += niter; stats[1] += lsiter; }
  // qacc_warmstart <- qacc
  for (int i = tid; i < d.nv; i += nt) s[L.warm + i] = qacc[i];
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// forward.euler (implicit joint damping when enabled) + _advance
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void euler(Ctx& c) {
  const Dims& d = c.d;
  const Lay& L = c.L;
  float* s = c.s;
  const int tid = TID, nt = c.nt;
  const float dt = d.timestep;
  float* qacc = s + L.qacc;
  PROF(c, 13);
  if (d.eulerdamp) {
    factor(c, s + L.M, true);
    PROF(c, 14);
    for (int i = tid; i < d.nv; i += nt) s[L.grad + i] = s[L.qfrc_smooth + i] + s[L.qfrc_con + i];
    __syncthreads();
    solve_m(c, s + L.grad, s + L.Mgrad, s + L.tmpv);
    qacc = s + L.Mgrad;
  }
  for (int a = tid; a < d.na; a += nt) s[L.act + a] += s[L.act_dot + a] * dt;
  for (int i = tid; i < d.nv; i += nt) s[L.qvel + i] += qacc[i] * dt;
  __syncthreads();
  const int* jtype = c.fi(VNL_F_JNT_TYPE);
  const int* jqadr = c.fi(VNL_F_JNT_QPOSADR);
  const int* jdofadr = c.fi(VNL_F_JNT_DOFADR);
  for (int j = tid; j < d.njnt; j += nt) {
    const int qa = jqadr[j], da = jdofadr[j];
    if (jtype[j] == 0) {
      for (int k = 0; k < 3; ++k) s[L.qpos + qa + k] += s[L.qvel + da + k] * dt;
      V3 w = ld3(s + L.qvel + da + 3);
      const float norm = normalize3(w);
      Q4 q = quat_normalize(quat_mul(ld4(s + L.qpos + qa + 3), axis_angle_quat(w, dt * norm)));
      st4(s + L.qpos + qa + 3, q);
    } else {
      s[L.qpos + qa] += s[L.qvel + da] * dt;
    }
  }
  __syncthreads();
  PROF(c, 15);
}

__device__ __forceinline__ float nan_to_num(float v) {
  if (isnan(v)) return 0.0f;
  if (isinf(v)) return v > 0.0f ? 3.402823466e+38f : -3.402823466e+38f;
  return v;
}

// ---------------------------------------------------------------------------------------------
// kernel: MODE 0 = env step, 1 = env reset tail, 2 = physics only, 3 = forward stage dump
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads, 4) vnl_env_kernel(Params p) {
  extern __shared__ __align__(16) float smem[];
  const int e = blockIdx.x;
  if (e >= p.B) return;
  Ctx& c = *reinterpret_cast<Ctx*>(smem);
  float* s = smem + kCtxFloats;
  const int tid = TID, nt = blockDim.x;
  if (tid == 0) {
    c.s = s;
    c.mb = p.model;
    c.nt = blockDim.x; c.nw = blockDim.x >> 5;
    c.prof = (p.prof && blockIdx.x == p.prof_block) ? p.prof : nullptr;
    c.t0 = clock64();
    c.d = p.dims;
    make_layout(c.d, c.L);
    const Lay& L0 = c.L;
    c.mcol = (const uint8_t*)(s + L0.mcol8); c.mrow = (const uint8_t*)(s + L0.mrow8); c.drow = (const uint8_t*)(s + L0.drow8);
    c.dls = (const uint8_t*)(s + L0.dls8); c.dld = (const uint8_t*)(s + L0.dld8);
    c.madr = (const uint16_t*)(s + L0.madr16); c.dadr = (const uint16_t*)(s + L0.dadr16); c.dent = (const uint16_t*)(s + L0.dent16);
  }
  for (int f = tid; f < VNL_F_MODEL_COUNT; f += nt) c.foff[f] = p.model[VNL_TABLE_OFF + 2 * f];
  __syncthreads();
  const Dims& d = c.d;
  const Lay& L = c.L;
  {
    uint8_t* mcol8 = (uint8_t*)(s + L.mcol8); uint8_t* mrow8 = (uint8_t*)(s + L.mrow8); uint8_t* drow8 = (uint8_t*)(s + L.drow8);
    uint8_t* dls8 = (uint8_t*)(s + L.dls8); uint8_t* dld8 = (uint8_t*)(s + L.dld8);
    uint16_t* madr16 = (uint16_t*)(s + L.madr16); uint16_t* dadr16 = (uint16_t*)(s + L.dadr16); uint16_t* dent16 = (uint16_t*)(s + L.dent16);
    const int* g_mcol = vnl_field_i(p.model, VNL_F_M_COL); const int* g_mrow = vnl_field_i(p.model, VNL_F_M_ROW);
    const int* g_madr = vnl_field_i(p.model, VNL_F_DOF_MADR); const int* g_dadr = vnl_field_i(p.model, VNL_F_DESC_ADR);
    const int* g_dent = vnl_field_i(p.model, VNL_F_DESC_ENTRY);
    const int* g_dls = vnl_field_i(p.model, VNL_F_DOFLEVEL_START); const int* g_dld = vnl_field_i(p.model, VNL_F_DOFLEVEL_DOF);
    for (int i = tid; i < d.nM; i += nt) { mcol8[i] = (uint8_t)g_mcol[i]; mrow8[i] = (uint8_t)g_mrow[i]; }
    for (int i = tid; i <= d.nv; i += nt) { madr16[i] = (uint16_t)g_madr[i]; dadr16[i] = (uint16_t)g_dadr[i]; }
    for (int i = tid; i < d.nM - d.nv; i += nt) { const int en = g_dent[i]; dent16[i] = (uint16_t)en; drow8[i] = (uint8_t)g_mrow[en]; }
    for (int i = tid; i < d.maxdepth + 2; i += nt) dls8[i] = (uint8_t)g_dls[i];
    for (int i = tid; i < d.nv; i += nt) dld8[i] = (uint8_t)g_dld[i];
  }
  int* ints = (int*)(s + L.ints);
  if (tid < 16) ints[tid] = 0;

  // ---- load state ------------------------------------------------------------------------
  for (int i = tid; i < d.nq; i += nt) s[L.qpos + i] = p.in.qpos[(size_t)e * d.nq + i];
  for (int i = tid; i < d.nv; i += nt) s[L.qvel + i] = p.in.qvel[(size_t)e * d.nv + i];
  if (MODE == 1) {
    for (int i = tid; i < d.na; i += nt) s[L.act + i] = 0.0f;
    for (int i = tid; i < d.nv; i += nt) s[L.warm + i] = 0.0f;
    for (int i = tid; i < d.nu; i += nt) s[L.ctrl + i] = 0.0f;
  } else {
    for (int i = tid; i < d.na; i += nt) s[L.act + i] = p.in.act ? p.in.act[(size_t)e * d.na + i] : 0.0f;
    for (int i = tid; i < d.nv; i += nt) s[L.warm + i] = p.in.qacc_warmstart ? p.in.qacc_warmstart[(size_t)e * d.nv + i] : 0.0f;
  }
  __syncthreads();
  if (MODE != 1) {
    const int* climited = c.fi(VNL_F_ACT_CTRLLIMITED);
    const float* crange = c.ff(VNL_F_ACT_CTRLRANGE);
    for (int u = tid; u < d.nu; u += nt) {
      float v = p.ctrl ? p.ctrl[(size_t)e * d.nu + u] : 0.0f;
      if (climited[u]) v = fminf(fmaxf(v, crange[2 * u]), crange[2 * u + 1]);
      s[L.ctrl + u] = v;
    }
  }
  const uint32_t* tb = p.task;
  // termination error of the PREVIOUS state and frame (envs/rodent.py:241-264, quirks Q2 / Q9):
  // depends only on inputs, so evaluate it before the physics overwrites them.
  float rtrunk = 0.0f;
  int frame_old = 0;
  if (MODE == 0) {
    frame_old = p.in.cur_frame[e];
    const int T = vnl_hdr_i(tb, VNL_TH_CLIP_LEN), ntrack = vnl_hdr_i(tb, VNL_TH_NTRACK), nj = d.nq - 7;
    const int f = min(max(frame_old, 0), T - 1);
    const float* rj = vnl_field_f(tb, VNL_T_JOINTS) + (size_t)f * nj;
    const float* rb = vnl_field_f(tb, VNL_T_BODY_POSITIONS) + (size_t)f * ntrack * 3;
    const int* bidx = vnl_field_i(tb, VNL_T_BODY_IDXS);
    const float* xold = p.in.xpos + (size_t)e * d.nbody * 3;
    float v[4] = {0, 0, 0, 0};
    for (int j = tid; j < nj; j += nt) v[0] += fabsf(rj[j] - s[L.qpos + 7 + j]);
    for (int b = tid; b < ntrack; b += nt)
      for (int k = 0; k < 3; ++k) v[1 + k] += fabsf(rb[3 * b + k] - xold[3 * bidx[b] + k]);
    block_sum<4>(c, v);
    const float eb = fmaxf(v[1], fmaxf(v[2], v[3]));
    const float err = 0.5f * vnl_hdr_f(tb, VNL_TH_BODY_ERR_MULT) * eb + 0.5f * v[0];
    rtrunk = 1.0f - err / vnl_hdr_f(tb, VNL_TH_TERM_THRESHOLD);
  }
  __syncthreads();

  int* stats = ints + 4;  // [4..7]
  float* dump = (MODE == 3) ? p.dump + (size_t)e * d.dump_total : nullptr;
  const int nsteps = (MODE == 1 || MODE == 3) ? 1 : p.nsteps;
  for (int st = 0; st < nsteps; ++st) {
    forward(c, stats, dump);
    if (MODE == 1 || MODE == 3) break;
    euler(c);
  }

  if (MODE == 3) {
    // stage dump (layout = oracle.dump_layout); arrays this formulation never materialises stay NaN
    auto put = [&](int off, const float* src, int n) { for (int i = tid; i < n; i += nt) dump[off + i] = src[i]; };
    put(d.dump_xpos, s + L.xpos, d.nbody * 3);
    put(d.dump_xpos + d.nbody * 3, s + L.xquat, d.nbody * 4);
    put(d.dump_xipos, s + L.xipos, d.nbody * 3);
    put(d.dump_xanchor, s + L.xanchor, d.njnt * 3);
    put(d.dump_xanchor + d.njnt * 3, s + L.xaxis, d.njnt * 3);
    put(d.dump_cinert, s + L.cinert, d.nbody * 10);
    put(d.dump_cinert + d.nbody * 10, s + L.cdof, d.nv * 6);
    put(d.dump_cinert + d.nbody * 10 + d.nv * 6, s + L.crb, d.nbody * 10);
    const int* rootid = c.fi(VNL_F_BODY_ROOTID);
    for (int b = tid; b < d.nbody; b += nt)
      if (b > 0 && rootid[b] == b) st3(dump + d.dump_subtree_com + 3 * b, ld3(s + L.rcom + 3 * b));
    const uint8_t* mrow = c.mrow; const uint8_t* mcol = c.mcol;
    for (int i = tid; i < d.nv * d.nv; i += nt) dump[d.dump_qM + i] = 0.0f;
    __syncthreads();
    for (int q = tid; q < d.nM; q += nt) {
      dump[d.dump_qM + mrow[q] * d.nv + mcol[q]] = s[L.M + q];
      dump[d.dump_qM + mcol[q] * d.nv + mrow[q]] = s[L.M + q];
    }
    put(d.dump_cvel, s + L.cvel, d.nbody * 6);
    put(d.dump_passive + 2 * d.nv, s + L.qfrc_act, d.nv);
    put(d.dump_passive + 3 * d.nv, s + L.act_dot, d.na);
    put(d.dump_passive + 3 * d.nv + d.na, s + L.qfrc_smooth, d.nv);
    put(d.dump_passive + 4 * d.nv + d.na, s + L.qacc_smooth, d.nv);
    put(d.dump_qacc, s + L.qacc, d.nv);
    put(d.dump_qacc + d.nv, s + L.qfrc_con, d.nv);
    if (tid < 4) dump[d.dump_total - 4 + tid] = (float)stats[tid];
    return;
  }

  // ---- write state -------------------------------------------------------------------------
  const VnlState& o = p.out;
  for (int i = tid; i < d.nq; i += nt) o.qpos[(size_t)e * d.nq + i] = s[L.qpos + i];
  for (int i = tid; i < d.nv; i += nt) o.qvel[(size_t)e * d.nv + i] = s[L.qvel + i];
  for (int i = tid; i < d.na; i += nt) o.act[(size_t)e * d.na + i] = s[L.act + i];
  for (int i = tid; i < d.nv; i += nt) o.qacc_warmstart[(size_t)e * d.nv + i] = s[L.warm + i];
  for (int i = tid; i < d.nbody * 3; i += nt) o.xpos[(size_t)e * d.nbody * 3 + i] = s[L.xpos + i];
  for (int i = tid; i < d.nbody * 4; i += nt) o.xquat[(size_t)e * d.nbody * 4 + i] = s[L.xquat + i];
  for (int i = tid; i < d.nv; i += nt) o.qfrc_actuator[(size_t)e * d.nv + i] = s[L.qfrc_act + i];
  const int torso = (MODE == 2) ? 1 : vnl_hdr_i(tb, VNL_TH_TORSO_BODY);
  if (tid < 3) o.subtree_com[(size_t)e * 3 + tid] = s[L.rcom + 3 * torso + tid];
  if (MODE == 2) {
    if (p.stats && tid < 4) p.stats[4 * e + tid] = stats[tid];
    return;
  }

  // ---- task outputs: obs, traj, reward, done (envs/rodent.py:178-239 / 149-176) -----------------------
  const int T = vnl_hdr_i(tb, VNL_TH_CLIP_LEN), ref_len = vnl_hdr_i(tb, VNL_TH_REF_LEN), ntrack = vnl_hdr_i(tb, VNL_TH_NTRACK);
  const int njidx = vnl_hdr_i(tb, VNL_TH_NJIDX), napp = vnl_hdr_i(tb, VNL_TH_NAPP), nee = vnl_hdr_i(tb, VNL_TH_NEE);
  const int obs_size = vnl_hdr_i(tb, VNL_TH_OBS_SIZE), traj_size = vnl_hdr_i(tb, VNL_TH_TRAJ_SIZE), nj = d.nq - 7;
  const int* bidx = vnl_field_i(tb, VNL_T_BODY_IDXS);
  const int* eeidx = vnl_field_i(tb, VNL_T_EE_IDX);
  const int* appidx = vnl_field_i(tb, VNL_T_APP_IDX);
  const int* apprefidx = vnl_field_i(tb, VNL_T_APP_REF_IDX);
  const int* jcol = vnl_field_i(tb, VNL_T_JOINT_COL);
  const float* rbody = vnl_field_f(tb, VNL_T_BODY_POSITIONS);
  const float* rpos = vnl_field_f(tb, VNL_T_POSITION);
  const float* rjoints = vnl_field_f(tb, VNL_T_JOINTS);
  int cur_frame, sub_clip_frame;
  if (MODE == 0) { cur_frame = frame_old + 1; sub_clip_frame = p.in.sub_clip_frame[e] + 1; }
  else { cur_frame = p.in.cur_frame[e]; sub_clip_frame = 0; }
  if (tid == 0) { o.cur_frame[e] = cur_frame; o.sub_clip_frame[e] = sub_clip_frame; }
  {
    float* obs = p.outputs.obs + (size_t)e * obs_size;
    for (int i = tid; i < obs_size; i += nt) {
      float v;
      if (i < d.nq) v = s[L.qpos + i];
      else if (i < d.nq + d.nv) v = s[L.qvel + i - d.nq];
      else if (i < d.nq + 2 * d.nv) v = s[L.qfrc_act + i - d.nq - d.nv];
      else { const int k = i - d.nq - 2 * d.nv; v = s[L.xpos + 3 * eeidx[k / 3] + k % 3]; }
      obs[i] = (MODE == 0) ? nan_to_num(v) : v;
    }
    float R[9];
    quat_to_mat(ld4(s + L.xquat + 4 * torso), R);
    const int ws = min(max(cur_frame + 1, 0), T - ref_len);
    float* traj = p.outputs.traj + (size_t)e * traj_size;
    const int n_app = ref_len * napp * 3, n_bod = ref_len * ntrack * 3, n_root = ref_len * 3;
    for (int i = tid; i < traj_size; i += nt) {
      float v;
      if (i < n_app) {
        const int w = i / (napp * 3), r = i - w * napp * 3;
        v = rbody[((size_t)(ws + w) * ntrack + apprefidx[r / 3]) * 3 + r % 3];
      } else if (i < n_app + n_bod) {
        const int q = i - n_app, w = q / (ntrack * 3), r = q - w * ntrack * 3, b = r / 3, cl = r % 3;
        const float* rb = rbody + ((size_t)(ws + w) * ntrack + b) * 3;
        const float* xp = s + L.xpos + 3 * bidx[b];
        v = (rb[0] - xp[0]) * R[cl] + (rb[1] - xp[1]) * R[3 + cl] + (rb[2] - xp[2]) * R[6 + cl];
      } else if (i < n_app + 2 * n_bod) {
        const int q = i - n_app - n_bod, w = q / (ntrack * 3), r = q - w * ntrack * 3, b = r / 3, k = r % 3;
        v = rbody[((size_t)(ws + w) * ntrack + b) * 3 + k] - s[L.xpos + 3 * bidx[b] + k];
      } else if (i < n_app + 2 * n_bod + n_root) {
        const int q = i - n_app - 2 * n_bod, w = q / 3, cl = q % 3;
        const float* rp = rpos + (size_t)(ws + w) * 3;
        v = (rp[0] - s[L.qpos]) * R[cl] + (rp[1] - s[L.qpos + 1]) * R[3 + cl] + (rp[2] - s[L.qpos + 2]) * R[6 + cl];
      } else {
        const int q = i - n_app - 2 * n_bod - n_root, w = q / njidx, j = jcol[q - w * njidx];
        v = rjoints[(size_t)(ws + w) * nj + j] - s[L.qpos + 7 + j];
      }
      traj[i] = v;
    }
  }
  if (MODE == 1) {
    // info["termination_error"] of the fresh state (rodent.py:169)
    const int f = min(max(cur_frame, 0), T - 1);
    float v[4] = {0, 0, 0, 0};
    for (int j = tid; j < nj; j += nt) v[0] += fabsf(rjoints[(size_t)f * nj + j] - s[L.qpos + 7 + j]);
    for (int b = tid; b < ntrack; b += nt)
      for (int k = 0; k < 3; ++k) v[1 + k] += fabsf(rbody[((size_t)f * ntrack + b) * 3 + k] - s[L.xpos + 3 * bidx[b] + k]);
    block_sum<4>(c, v);
    const float eb = fmaxf(v[1], fmaxf(v[2], v[3]));
    const float err = 0.5f * vnl_hdr_f(tb, VNL_TH_BODY_ERR_MULT) * eb + 0.5f * v[0];
    if (tid == 0) {
      p.outputs.reward[e] = 0.0f; p.outputs.done[e] = 0.0f;
      float* mt = p.outputs.metrics + 7 * (size_t)e;
      for (int k = 0; k < 6; ++k) mt[k] = 0.0f;
      mt[6] = 1.0f - err / vnl_hdr_f(tb, VNL_TH_TERM_THRESHOLD);
    }
    if (p.outputs.stats && tid < 4) p.outputs.stats[4 * e + tid] = stats[tid];
    return;
  }
  // _calculate_reward (rodent.py:266-316): every reference lookup uses the OLD cur_frame
  {
    const int f = min(max(frame_old, 0), T - 1);
    const float* rvel = vnl_field_f(tb, VNL_T_VELOCITY) + (size_t)f * 3;
    const float* rang = vnl_field_f(tb, VNL_T_ANGULAR_VELOCITY) + (size_t)f * 3;
    const float* rjv = vnl_field_f(tb, VNL_T_JOINTS_VELOCITY) + (size_t)f * (d.nv - 6);
    float v[4] = {0, 0, 0, 0};  // |qvel - ref|^2, sum qfrc_actuator^2, |app - ref|^2, nan count
    for (int i = tid; i < d.nv; i += nt) {
      const float ref = i < 3 ? rvel[i] : (i < 6 ? rang[i - 3] : rjv[i - 6]);
      const float df = s[L.qvel + i] - ref;
      v[0] += df * df;
      const float qa = s[L.qfrc_act + i];
      v[1] += qa * qa;
      if (isnan(s[L.qvel + i]) || isnan(s[L.warm + i]) || isnan(qa)) v[3] += 1.0f;
    }
    for (int i = tid; i < napp * 3; i += nt) {
      const int a = i / 3, k = i % 3;
      const float df = s[L.xpos + 3 * appidx[a] + k] - rbody[((size_t)f * ntrack + apprefidx[a]) * 3 + k];
      v[2] += df * df;
    }
    for (int i = tid; i < d.nq; i += nt) if (isnan(s[L.qpos + i])) v[3] += 1.0f;
    for (int i = tid; i < d.nbody * 3; i += nt) if (isnan(s[L.xpos + i])) v[3] += 1.0f;
    for (int i = tid; i < d.na; i += nt) if (isnan(s[L.act + i])) v[3] += 1.0f;
    block_sum<4>(c, v);
    if (tid == 0) {
      const int cri = vnl_hdr_i(tb, VNL_TH_COM_REF_IDX);
      V3 dc = ld3(s + L.rcom + 3 * torso) - ld3(rbody + ((size_t)f * ntrack + cri) * 3);
      float rcom = expf(-100.0f * sqrtf(dot(dc, dc)));
      float rvl = expf(-0.1f * sqrtf(v[0]));
      Q4 qc = quat_normalize(ld4(s + L.qpos + 3));
      Q4 qr = quat_normalize(ld4(vnl_field_f(tb, VNL_T_QUATERNION) + (size_t)f * 4));
      const float dq = qc.w * qr.w + qc.x * qr.x + qc.y * qr.y + qc.z * qr.z;
      const float dist = fminf(1.0f, 2.0f * dq * dq - 1.0f);
      float rquat = expf(-2.0f * fabsf(0.5f * acosf(dist)));
      float ract = -0.015f * (v[1] / (float)d.nv);
      float rapp = expf(-400.0f * sqrtf(v[2]));
      const float z = s[L.qpos + 2];
      float healthy = z < vnl_hdr_f(tb, VNL_TH_HEALTHY_LO) ? 0.0f : 1.0f;
      if (z > vnl_hdr_f(tb, VNL_TH_HEALTHY_HI)) healthy = 0.0f;
      rcom *= 0.01f; rvl *= 0.01f; rapp *= 0.01f;
      float rtr = rtrunk * 0.01f;
      rquat *= 0.01f; ract *= 0.0001f;
      const float total = rcom + rvl + rtr + rquat + ract + rapp;
      const float sub_healthy = sub_clip_frame < vnl_hdr_i(tb, VNL_TH_SUB_CLIP_LEN) ? 1.0f : 0.0f;
      float done = rtr < 0.0f ? 1.0f : 0.0f;
      done = fmaxf(1.0f - healthy, done);
      done = fmaxf(1.0f - sub_healthy, done);
      if (v[3] > 0.0f) done = 1.0f;
      p.outputs.reward[e] = nan_to_num(total);
      p.outputs.done[e] = done;
      float* mt = p.outputs.metrics + 7 * (size_t)e;
      mt[0] = rcom; mt[1] = rvl; mt[2] = rtr; mt[3] = rquat; mt[4] = ract; mt[5] = rapp; mt[6] = rtr;
      ints[8] = done > 0.0f;
    }
    if (p.outputs.stats && tid < 4) p.outputs.stats[4 * e + tid] = stats[tid];
    // brax AutoResetWrapper.step fused in: where done, the pipeline-state leaves and obs are replaced by the cached
    // first ones; info (frames, traj), reward, done and metrics are kept.
    if (p.first.qpos) {
      __syncthreads();
      if (ints[8]) {
        const VnlState& f = p.first;
        for (int i = tid; i < d.nq; i += nt) o.qpos[(size_t)e * d.nq + i] = f.qpos[(size_t)e * d.nq + i];
        for (int i = tid; i < d.nv; i += nt) o.qvel[(size_t)e * d.nv + i] = f.qvel[(size_t)e * d.nv + i];
        for (int i = tid; i < d.na; i += nt) o.act[(size_t)e * d.na + i] = f.act[(size_t)e * d.na + i];
        for (int i = tid; i < d.nv; i += nt) o.qacc_warmstart[(size_t)e * d.nv + i] = f.qacc_warmstart[(size_t)e * d.nv + i];
        for (int i = tid; i < d.nbody * 3; i += nt) o.xpos[(size_t)e * d.nbody * 3 + i] = f.xpos[(size_t)e * d.nbody * 3 + i];
        for (int i = tid; i < d.nbody * 4; i += nt) o.xquat[(size_t)e * d.nbody * 4 + i] = f.xquat[(size_t)e * d.nbody * 4 + i];
        for (int i = tid; i < d.nv; i += nt) o.qfrc_actuator[(size_t)e * d.nv + i] = f.qfrc_actuator[(size_t)e * d.nv + i];
        if (tid < 3) o.subtree_com[(size_t)e * 3 + tid] = f.subtree_com[(size_t)e * 3 + tid];
        if (p.first_obs)
          for (int i = tid; i < obs_size; i += nt) p.outputs.obs[(size_t)e * obs_size + i] = p.first_obs[(size_t)e * obs_size + i];
      }
    }
  }
}

template __global__ void vnl_env_kernel<0>(Params);
template __global__ void vnl_env_kernel<1>(Params);
template __global__ void vnl_env_kernel<2>(Params);
template __global__ void vnl_env_kernel<3>(Params);

int smem_bytes(const Dims& d) {
  Lay L;
  make_layout(d, L);
  return (L.total + kCtxFloats) * (int)sizeof(float);
}

cudaError_t launch(int mode, const Params& p, cudaStream_t stream) {
  const int bytes = smem_bytes(p.dims);
  void (*k)(Params) = mode == 0 ? vnl_env_kernel<0> : mode == 1 ? vnl_env_kernel<1> : mode == 2 ? vnl_env_kernel<2> : vnl_env_kernel<3>;
  cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (err != cudaSuccess) return err;
  // leave the rest of the unified 228 KB array to L1: the model tables are re-read from it every substep
  int ctas = 4;
  while (ctas > 1 && ctas * (bytes + 1024) > 227 * 1024) --ctas;
  int pct = (ctas * (bytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
  k<<<p.B, kThreads, bytes, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace vnl

// vnl_kernels.h -- shared declarations between the kernels and the C ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vnl_b200.h"

namespace vnl {

constexpr int kThreads = 128;  // one CTA per env
constexpr int kMaxWarps = 8;

struct Dims {
  int nq, nv, nu, na, nbody, njnt, ngeom, npair, ncon, nlimit, nefc, nM, nlevel, maxdepth;
  int solver, iterations, ls_iterations, eulerdamp;
  float timestep, gx, gy, gz, tolerance, ls_tolerance, impratio, meaninertia;
  // stage-dump offsets (layout of oracle.dump_layout)
  int dump_xpos, dump_xipos, dump_xanchor, dump_subtree_com, dump_cinert, dump_qM, dump_cvel, dump_passive, dump_con,
      dump_efc, dump_qacc, dump_total;
};

struct Lay {
  int qpos, qvel, act, ctrl, warm, xpos, xquat, xipos, xanchor, xaxis, rcom, cinert, crb, cdof, cdofdot, cvel, cacc, cfrc, M,
      Lf, K, qfrc_smooth, qacc_smooth, qfrc_act, act_dot, lim_dof, lim_sign, limrow_of_dof, cbody, crel, cframe, cmu, cwrench,
      efcD, aref, Jaref, Jv, qacc, Ma, grad, Mgrad, search, Mv, qfrc_con, tmpv, red, ints, mcol8, mrow8, madr16, dadr16,
      dent16, drow8, dls8, dld8, total;
};

struct Params {
  const uint32_t* model;
  const uint32_t* task;
  Dims dims;
  int B, nsteps;
  VnlState in, out;
  VnlState first;         // optional cached first state (AutoReset), first.qpos == nullptr when unused
  const float* first_obs;
  const float* ctrl;
  VnlOutputs outputs;
  int32_t* stats;
  float* dump;
  long long* prof;  // optional [32] per-phase clock64 accumulators of one CTA (developer hook)
  int prof_block;
};

int smem_bytes(const Dims& d);
cudaError_t launch(int mode, const Params& p, cudaStream_t stream);

}  // namespace vnl

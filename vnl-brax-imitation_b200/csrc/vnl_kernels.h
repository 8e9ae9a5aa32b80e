// vnl_kernels.h -- shared declarations between the kernels and the C ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vnl_b200.h"

namespace vnl {


struct Dims {
  int nq, nv, nu, na, nbody, njnt, ngeom, npair, ncon, nlimit, nefc, nM, nlevel, maxdepth, nroot;
  int solver, iterations, ls_iterations, eulerdamp;
  float timestep, gx, gy, gz, tolerance, ls_tolerance, impratio, meaninertia;
  int ktab_words;  // size of VNL_F_KTAB
  int ndslot;      // partial-sum slots of the descendant mat-vec program (VNL_KS_NDSLOT)
  int TA, TD;      // steps of the mat-vec lane programs (VNL_MH_TA / VNL_MH_TD)
  int naslot;      // partial-sum slots of the ancestor mat-vec program (VNL_MH_NASLOT)
  int env_warps;   // warps cooperating on one env (VNL_MH_ENV_WARPS): lane count of the mat-vec programs / 32
  // stage-dump offsets (layout of oracle.dump_layout)
  int dump_xpos, dump_xipos, dump_xanchor, dump_subtree_com, dump_cinert, dump_qM, dump_cvel, dump_passive, dump_con,
      dump_efc, dump_qacc, dump_total;
};

// Per-env shared-memory layout (float offsets); see make_layout for the regions and who lives in them when.
struct Lay {
  int qpos, qvel, act, ctrl, warm, cdof, Mdiag, Kdiag, rcom, qfrc_smooth, qacc_smooth, act_dot, ints;  // whole substep
  int Mn, H, jr;                                                          // Newton solver only (dense Hessian)
  int xpos, xquat, cvel;                                                  // region R0, kinematics .. constraint rows
  int Jaref, qacc, Ma, grad, Mgrad, search, Mv, qfrc_con;                 // region R0, solver .. integrator
  int xipos, xanchor, xaxis, cacc, t16;                                   // region R1, kinematic passes
  int part, tmpv, lim_dof, limrow_of_dof, cbody, crel, cframe, cmu, efcD, Jv;  // region R1, solves / constraint rows
  int K;                                                                  // region R1, factorisation workspace
  int total;
};

struct Params {
  const uint32_t* model;
  const uint32_t* task;
  Dims dims;
  int B, nsteps;
  VnlState in, out;
  VnlState first;         // optional cached first state (AutoReset), first.qpos == nullptr when unused
  const float* first_obs;
  VnlEpisode episode;     // optional EpisodeWrapper bookkeeping, episode.steps_out == nullptr when unused
  const float* ctrl;
  VnlOutputs outputs;
  int32_t* stats;
  float* dump;
  long long* prof;  // optional [32] per-phase clock64 accumulators of one env (developer hook)
  int prof_env;
  float* work;      // inertia workspace (vnl_set_workspace), work_stride floats per resident env
  int work_stride;
  int lsgroups;     // lockstep groups per CTA (1 = the whole CTA)
  int lockstep;     // 0 = warps free-run, 1 = CTA barrier at every substep start, 2 = also before the integrator, 3 = at every phase boundary (default)
};

struct LaunchInfo { int smem_bytes, warps_per_cta, ctas; };  // warps_per_cta = env groups per CTA

// floats of workspace per resident env: the A-order and the D-order copy of the off-diagonal entries of the inertia M,
// then the same two copies of its inverse factor K
inline int work_stride(const Dims& d) { return (2 * (d.TA + d.TD) * 32 * d.env_warps + 31) & ~31; }

// vnl_kernels.cu is compiled once per env-group width (-DVNL_EW=1, 2) into its own namespace.
namespace ew1 { LaunchInfo launch_info(const Dims& d, int B); cudaError_t launch(int mode, const Params& p, cudaStream_t stream); }
namespace ew2 { LaunchInfo launch_info(const Dims& d, int B); cudaError_t launch(int mode, const Params& p, cudaStream_t stream); }

inline LaunchInfo any_launch_info(const Dims& d, int B) { return d.env_warps == 2 ? ew2::launch_info(d, B) : ew1::launch_info(d, B); }
inline cudaError_t any_launch(int mode, const Params& p, cudaStream_t stream) {
  if (p.dims.env_warps == 2) return ew2::launch(mode, p, stream);
  if (p.dims.env_warps == 1) return ew1::launch(mode, p, stream);
  return cudaErrorInvalidValue;
}

}  // namespace vnl

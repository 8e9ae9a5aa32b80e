// vnl_capi.cu -- extern "C" boundary of include/vnl_b200.h.
//
// The model dimensions needed for the launch geometry live in the blob header.  The blobs
// passed to vnl_step & co. are DEVICE buffers, so the (tiny) header is cached host-side per
// device pointer the first time a blob is registered with vnl_register_blob(); no call on
// the step path copies, allocates or synchronises.
#include <cuda_runtime.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/vnl_blob.h"
#include "vnl_kernels.h"

namespace {

struct Header { uint32_t w[VNL_DATA_OFF]; };  // scalar header + field table
std::mutex g_mu;
std::unordered_map<const void*, Header> g_headers;  // device blob pointer -> host copy of its scalar header
struct Work { float* p; size_t bytes; };
std::unordered_map<const void*, Work> g_work;        // device model blob pointer -> bound inertia workspace

bool lookup(const void* dev, Header& h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_headers.find(dev);
  if (it == g_headers.end()) return false;
  h = it->second;
  return true;
}

void fill_dims(const uint32_t* w, vnl::Dims& d) {
  d.nq = vnl_hdr_i(w, VNL_MH_NQ); d.nv = vnl_hdr_i(w, VNL_MH_NV); d.nu = vnl_hdr_i(w, VNL_MH_NU); d.na = vnl_hdr_i(w, VNL_MH_NA);
  d.nbody = vnl_hdr_i(w, VNL_MH_NBODY); d.njnt = vnl_hdr_i(w, VNL_MH_NJNT); d.ngeom = vnl_hdr_i(w, VNL_MH_NGEOM);
  d.npair = vnl_hdr_i(w, VNL_MH_NPAIR); d.ncon = vnl_hdr_i(w, VNL_MH_NCON); d.nlimit = vnl_hdr_i(w, VNL_MH_NLIMIT);
  d.nefc = vnl_hdr_i(w, VNL_MH_NEFC); d.nM = vnl_hdr_i(w, VNL_MH_NM); d.nlevel = vnl_hdr_i(w, VNL_MH_NLEVEL);
  d.maxdepth = vnl_hdr_i(w, VNL_MH_MAXDEPTH); d.nroot = vnl_hdr_i(w, VNL_MH_NROOT); d.ndslot = vnl_hdr_i(w, VNL_MH_NDSLOT); d.env_warps = vnl_hdr_i(w, VNL_MH_ENV_WARPS); d.naslot = vnl_hdr_i(w, VNL_MH_NASLOT); d.TA = vnl_hdr_i(w, VNL_MH_TA); d.TD = vnl_hdr_i(w, VNL_MH_TD); d.ktab_words = (int)w[VNL_TABLE_OFF + 2 * VNL_F_KTAB + 1];
  d.solver = vnl_hdr_i(w, VNL_MH_SOLVER); d.iterations = vnl_hdr_i(w, VNL_MH_ITERATIONS);
  d.ls_iterations = vnl_hdr_i(w, VNL_MH_LS_ITERATIONS); d.eulerdamp = vnl_hdr_i(w, VNL_MH_EULERDAMP);
  d.timestep = vnl_hdr_f(w, VNL_MH_TIMESTEP); d.gx = vnl_hdr_f(w, VNL_MH_GRAVITY_X); d.gy = vnl_hdr_f(w, VNL_MH_GRAVITY_Y);
  d.gz = vnl_hdr_f(w, VNL_MH_GRAVITY_Z); d.tolerance = vnl_hdr_f(w, VNL_MH_TOLERANCE); d.ls_tolerance = vnl_hdr_f(w, VNL_MH_LS_TOLERANCE);
  d.impratio = vnl_hdr_f(w, VNL_MH_IMPRATIO); d.meaninertia = vnl_hdr_f(w, VNL_MH_MEANINERTIA);
  const int nb = d.nbody, nv = d.nv;
  d.dump_xpos = 0;
  d.dump_xipos = nb * 3 + nb * 4 + nb * 9;
  d.dump_xanchor = d.dump_xipos + nb * 3 + nb * 9;
  d.dump_subtree_com = d.dump_xanchor + d.njnt * 6;
  d.dump_cinert = d.dump_subtree_com + nb * 3;
  d.dump_qM = d.dump_cinert + nb * 10 + nv * 6 + nb * 10;
  d.dump_cvel = d.dump_qM + nv * nv;
  d.dump_passive = d.dump_cvel + nb * 6 + nv * 6;
  d.dump_con = d.dump_passive + 5 * nv + d.na;
  d.dump_efc = d.dump_con + 13 * d.ncon;
  d.dump_qacc = d.dump_efc + 3 * d.nefc + d.nefc * nv;
  d.dump_total = d.dump_qacc + 2 * nv + d.nefc + 4;
  vnl::any_decide_stream(d);
}

int check_blob(const void* host, size_t nbytes, uint32_t magic, int nfields) {
  if (!host || nbytes < sizeof(uint32_t) * VNL_DATA_OFF) return -1;
  const uint32_t* w = (const uint32_t*)host;
  if (w[0] != magic) return -2;
  if (w[1] != VNL_BLOB_VERSION) return -3;
  if ((size_t)w[2] * 4 != nbytes) return -4;
  if ((int)w[3] != nfields) return -5;
  for (int f = 0; f < nfields; ++f) {
    const uint64_t off = w[VNL_TABLE_OFF + 2 * f], n = w[VNL_TABLE_OFF + 2 * f + 1];
    if (off < VNL_DATA_OFF || off + n > w[2]) return -6;
  }
  return 0;
}

int prepare(const void* model, const void* task, bool need_task, vnl::Params& p) {
  Header hm;
  if (!model || !lookup(model, hm)) return -10;
  if (hm.w[0] != VNL_MAGIC_MODEL) return -11;
  fill_dims(hm.w, p.dims);
  if (p.dims.stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_work.find(model);
    if (it == g_work.end()) return -20;
    const vnl::LaunchInfo li = vnl::any_launch_info(p.dims, 1 << 20);
    p.work_stride = vnl::work_stride(p.dims);
    if ((size_t)li.ctas * li.warps_per_cta * p.work_stride * sizeof(float) > it->second.bytes) return -21;
    p.work = it->second.p;
  }
  p.model = (const uint32_t*)model;
  p.task = (const uint32_t*)task;
  if (need_task) {
    Header ht;
    if (!task || !lookup(task, ht)) return -12;
    if (ht.w[0] != VNL_MAGIC_TASK) return -13;
  }
  return 0;
}

}  // namespace

__global__ void vnl_ffma_probe_kernel(int iters, float* out) {
  float a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = 1.0f + 1e-3f * (float)(threadIdx.x + k);
  const float m = 0.9999f, c = 1e-4f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fmaf(a[k], m, c);
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keeps the chains alive; practically never true
}

extern "C" {

const char* vnl_version(void) { return "vnl_b200 0.1 (sm_100a)"; }

int vnl_check_model(const void* model_host, size_t nbytes) { return check_blob(model_host, nbytes, VNL_MAGIC_MODEL, VNL_F_MODEL_COUNT); }
int vnl_check_task(const void* task_host, size_t nbytes) { return check_blob(task_host, nbytes, VNL_MAGIC_TASK, VNL_TASK_COUNT); }

int vnl_register_blob(const void* blob_dev, const void* blob_host, size_t nbytes) {
  const uint32_t* w = (const uint32_t*)blob_host;
  if (!blob_dev || !w) return -1;
  const int rc = (w[0] == VNL_MAGIC_MODEL) ? vnl_check_model(blob_host, nbytes) : vnl_check_task(blob_host, nbytes);
  if (rc) return rc;
  Header h;
  memcpy(h.w, w, sizeof(h.w));
  std::lock_guard<std::mutex> lk(g_mu);
  g_headers[blob_dev] = h;
  return 0;
}

int vnl_unregister_blob(const void* blob_dev) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_work.erase(blob_dev);
  return g_headers.erase(blob_dev) ? 0 : -1;
}

size_t vnl_workspace_bytes(const void* model_host) {
  vnl::Dims d;
  fill_dims((const uint32_t*)model_host, d);
  const vnl::LaunchInfo li = vnl::any_launch_info(d, 1 << 20);
  return (size_t)li.ctas * li.warps_per_cta * vnl::work_stride(d) * sizeof(float);
}

int vnl_set_workspace(const void* model_dev, void* workspace_dev, size_t nbytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!model_dev || g_headers.find(model_dev) == g_headers.end()) return -1;
  if (!workspace_dev) { g_work.erase(model_dev); return 0; }
  g_work[model_dev] = Work{(float*)workspace_dev, nbytes};
  return 0;
}

int vnl_step_smem_bytes(const void* model_host) {
  vnl::Dims d;
  fill_dims((const uint32_t*)model_host, d);
  return vnl::any_launch_info(d, 1 << 20).smem_bytes;
}

int vnl_envs_per_cta(const void* model_host) {
  vnl::Dims d;
  fill_dims((const uint32_t*)model_host, d);
  return vnl::any_launch_info(d, 1 << 20).warps_per_cta;
}

int vnl_resident_envs(const void* model_host) {
  vnl::Dims d;
  fill_dims((const uint32_t*)model_host, d);
  const vnl::LaunchInfo li = vnl::any_launch_info(d, 1 << 20);
  return li.ctas * li.warps_per_cta;
}

int vnl_debug_layout(const void* model_host, const char** names, int32_t* offsets, int32_t* sizes, int cap) {
  vnl::Dims d;
  fill_dims((const uint32_t*)model_host, d);
  vnl::LayoutEntry e[64];
  const int n = vnl::ew1s1::layout_table(d, e, 64);
  for (int i = 0; i < n && i < cap; ++i) { names[i] = e[i].name; offsets[i] = e[i].offset; sizes[i] = e[i].size; }
  return n;
}

size_t vnl_dump_size(const void* model_host) {
  vnl::Dims d;
  fill_dims((const uint32_t*)model_host, d);
  return (size_t)d.dump_total;
}

int vnl_step(const void* model, const void* task, int B, const VnlState* in, const float* action, const VnlState* out,
             const VnlOutputs* outputs, void* stream) {
  if (B <= 0 || !in || !out || !outputs || !action) return -1;
  vnl::Params p;
  memset(&p, 0, sizeof(p));
  int rc = prepare(model, task, true, p);
  if (rc) return rc;
  Header ht;
  lookup(task, ht);
  p.B = B; p.nsteps = vnl_hdr_i(ht.w, VNL_TH_NFRAMES); p.in = *in; p.out = *out; p.ctrl = action; p.outputs = *outputs;
  return (int)vnl::any_launch(0, p, (cudaStream_t)stream);
}

int vnl_step_autoreset(const void* model, const void* task, int B, const VnlState* in, const float* action, const VnlState* out,
                       const VnlOutputs* outputs, const VnlState* first, const float* first_obs, void* stream) {
  if (B <= 0 || !in || !out || !outputs || !action || !first || !first->qpos) return -1;
  vnl::Params p;
  memset(&p, 0, sizeof(p));
  int rc = prepare(model, task, true, p);
  if (rc) return rc;
  Header ht;
  lookup(task, ht);
  p.B = B; p.nsteps = vnl_hdr_i(ht.w, VNL_TH_NFRAMES); p.in = *in; p.out = *out; p.ctrl = action; p.outputs = *outputs;
  p.first = *first; p.first_obs = first_obs;
  return (int)vnl::any_launch(0, p, (cudaStream_t)stream);
}

int vnl_step_training(const void* model, const void* task, int B, const VnlState* in, const float* action, const VnlState* out,
                      const VnlOutputs* outputs, const VnlState* first, const float* first_obs, const VnlEpisode* episode,
                      void* stream) {
  if (B <= 0 || !in || !out || !outputs || !action || !first || !first->qpos || !episode) return -1;
  if (!episode->steps_in || !episode->done_in || !episode->steps_out || !episode->truncation_out) return -1;
  vnl::Params p;
  memset(&p, 0, sizeof(p));
  int rc = prepare(model, task, true, p);
  if (rc) return rc;
  Header ht;
  lookup(task, ht);
  p.B = B; p.nsteps = vnl_hdr_i(ht.w, VNL_TH_NFRAMES); p.in = *in; p.out = *out; p.ctrl = action; p.outputs = *outputs;
  p.first = *first; p.first_obs = first_obs; p.episode = *episode;
  return (int)vnl::any_launch(0, p, (cudaStream_t)stream);
}

// Developer hook: vnl_step with per-phase clock64 accumulation for CTA `block` into prof[32].
int vnl_step_profiled(const void* model, const void* task, int B, const VnlState* in, const float* action, const VnlState* out,
                      const VnlOutputs* outputs, void* stream, long long* prof, int block) {
  if (B <= 0 || !in || !out || !outputs || !action) return -1;
  vnl::Params p;
  memset(&p, 0, sizeof(p));
  int rc = prepare(model, task, true, p);
  if (rc) return rc;
  Header ht;
  lookup(task, ht);
  p.B = B; p.nsteps = vnl_hdr_i(ht.w, VNL_TH_NFRAMES); p.in = *in; p.out = *out; p.ctrl = action; p.outputs = *outputs;
  p.prof = prof; p.prof_env = block;
  return (int)vnl::any_launch(0, p, (cudaStream_t)stream);
}

int vnl_reset(const void* model, const void* task, int B, const VnlState* in, const VnlState* out, const VnlOutputs* outputs,
              void* stream) {
  if (B <= 0 || !in || !out || !outputs) return -1;
  vnl::Params p;
  memset(&p, 0, sizeof(p));
  int rc = prepare(model, task, true, p);
  if (rc) return rc;
  p.B = B; p.nsteps = 1; p.in = *in; p.out = *out; p.outputs = *outputs;
  return (int)vnl::any_launch(1, p, (cudaStream_t)stream);
}

int vnl_pipeline_step(const void* model, int B, int nsteps, const VnlState* in, const float* ctrl, const VnlState* out,
                      int32_t* stats, void* stream) {
  if (B <= 0 || nsteps <= 0 || !in || !out) return -1;
  vnl::Params p;
  memset(&p, 0, sizeof(p));
  int rc = prepare(model, nullptr, false, p);
  if (rc) return rc;
  p.B = B; p.nsteps = nsteps; p.in = *in; p.out = *out; p.ctrl = ctrl; p.stats = stats;
  return (int)vnl::any_launch(2, p, (cudaStream_t)stream);
}

int vnl_forward_dump(const void* model, int B, const VnlState* in, const float* ctrl, float* dump, void* stream) {
  if (B <= 0 || !in || !dump) return -1;
  vnl::Params p;
  memset(&p, 0, sizeof(p));
  int rc = prepare(model, nullptr, false, p);
  if (rc) return rc;
  p.B = B; p.nsteps = 1; p.in = *in; p.ctrl = ctrl; p.dump = dump;
  cudaError_t err = cudaMemsetAsync(dump, 0xFF, (size_t)B * p.dims.dump_total * sizeof(float), (cudaStream_t)stream);
  if (err != cudaSuccess) return (int)err;
  return (int)vnl::any_launch(3, p, (cudaStream_t)stream);
}

// Legacy XLA custom calls.  `opaque` = two little-endian int32: B, then the operand layout version (1).
// buffers: [model, task, qpos, qvel, act, warm, xpos, xquat, subtree_com, qfrc_actuator, cur_frame, sub_clip_frame, action,
//           (outputs) qpos', qvel', act', warm', xpos', xquat', subtree_com', qfrc_actuator', cur_frame', sub_clip_frame',
//           obs, traj, reward, done, metrics, stats]
static void unpack(void** b, int first, VnlState& s) {
  s.qpos = (float*)b[first]; s.qvel = (float*)b[first + 1]; s.act = (float*)b[first + 2]; s.qacc_warmstart = (float*)b[first + 3];
  s.xpos = (float*)b[first + 4]; s.xquat = (float*)b[first + 5]; s.subtree_com = (float*)b[first + 6];
  s.qfrc_actuator = (float*)b[first + 7]; s.cur_frame = (int32_t*)b[first + 8]; s.sub_clip_frame = (int32_t*)b[first + 9];
}
void vnl_xla_step(void* stream, void** buffers, const char* opaque, size_t opaque_len) {
  if (opaque_len < 4) return;
  int B;
  memcpy(&B, opaque, 4);
  VnlState in, out;
  unpack(buffers, 2, in);
  unpack(buffers, 13, out);
  VnlOutputs o;
  o.obs = (float*)buffers[23]; o.traj = (float*)buffers[24]; o.reward = (float*)buffers[25]; o.done = (float*)buffers[26];
  o.metrics = (float*)buffers[27]; o.stats = (int32_t*)buffers[28];
  vnl_step(buffers[0], buffers[1], B, &in, (const float*)buffers[12], &out, &o, stream);
}
void vnl_xla_reset(void* stream, void** buffers, const char* opaque, size_t opaque_len) {
  if (opaque_len < 4) return;
  int B;
  memcpy(&B, opaque, 4);
  VnlState in, out;
  unpack(buffers, 2, in);
  unpack(buffers, 13, out);
  VnlOutputs o;
  o.obs = (float*)buffers[23]; o.traj = (float*)buffers[24]; o.reward = (float*)buffers[25]; o.done = (float*)buffers[26];
  o.metrics = (float*)buffers[27]; o.stats = (int32_t*)buffers[28];
  vnl_reset(buffers[0], buffers[1], B, &in, &out, &o, stream);
}

// Measurement helper for bench.py: an FFMA-saturating microkernel (8 independent chains per thread) that gives the
// FP32 CUDA-core roofline denominator on the device the benchmark runs on.  flops = blocks * 256 * iters * 16 * 2.
int vnl_ffma_probe(int blocks, int iters, float* out, void* stream) {
  if (blocks <= 0 || iters <= 0 || !out) return -1;
  vnl_ffma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, out);
  return (int)cudaGetLastError();
}

}  // extern "C"

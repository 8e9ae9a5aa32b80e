// Intention-network policy forward: the step on the other side of env.step in the rollout
// (ppo_imitation/acting.py:47-48 -> ppo_networks.py:45-83 -> intention_policy_network.py:20-105), one launch.
//
// One CTA per 128 envs (tile M = 128 = the tcgen05 accumulator height: env r of the tile is tensor-memory lane r).
// Six dense layers run as tcgen05.mma.cta_group::1.kind::f16 (bf16 operands from shared memory, fp32 accumulators in
// tensor memory); everything between two layers happens in the epilogue of the first, on the SM:
//
//   L0  traj[795 -> 832]  x W[.,256]  -> +b, relu, LayerNorm -> bf16 A-operand of L1     (K streamed in 64-wide chunks through
//   L1  [256] x W[.,128]              -> +b, relu, LayerNorm                               a 3-slot ring: warps 1-7 load, thread 0 issues)
//   L2  [128] x [W_mean | W_logvar]   -> +b, z = mean + eps_z * exp(logvar / 2)
//   L3  [z | (obs - mu) / sigma] (296 -> 304) x W[.,128] -> +b, relu, LayerNorm
//   L4  [128] x W[.,256]              -> +b, relu, LayerNorm
//   L5  [256] x W[.,60 -> 64]         -> +b = logits; tanh-normal sample, log-prob, outputs
//
// Operand image (both A and B, K-major, no swizzle): 8-row x 16-byte core matrices; element (r, k) of an R-row operand
// sits at (k / 8) * LBO + r * 16 + (k % 8) * 2 with LBO = R * 16 + 16.  The 16 bytes of slack per K-group rotate the
// banks so that the coalesced activation loader (a lane owns 2 consecutive k of one row) stores conflict-free; the
// stride between 8-row groups (SBO) is 128.  Weights are packed into exactly this image on the host (vnl_policy_pack),
// so a layer's B operand (or a K chunk of it) is ONE contiguous TMA bulk copy (cp.async.bulk, completion counted in bytes
// on an mbarrier): written by the async proxy, read by the tensor core, no thread touches it.
//
// All 8 warps run the epilogues (warps w and w + 4 share the tensor-memory lanes of quadrant w % 4 = env rows 32 (w % 4) ..
// + 31 and split the accumulator columns in halves, 32x32b loads); the next layer's weights stream in by TMA while the
// epilogue computes; thread 0 issues the weight copies and the MMAs (committed to mbarriers).  In layer 0 warp 0 does
// nothing else: the seven other warps load, convert and store the activations and signal per-slot mbarriers, so the
// ~130 cycles each MMA takes to issue are in nobody's load iteration.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/vnl_policy.h"
#include "vnl_xla_status.h"

namespace {

constexpr int TILE_M = 128;
constexpr int THREADS = 256;
constexpr int KCHUNK = 64;
constexpr uint32_t LBO_A = TILE_M * 16 + 16;
constexpr uint32_t SBO = 128;
constexpr uint32_t R0_BYTES = 96 * 1024;  // A operands (activations), later the logits scratch
constexpr uint32_t R1_BYTES = 100 * 1024; // B operands (weights; 3 ring slots in L0)
constexpr int RING = 3;                   // L0 ring depth
constexpr int MAX_PARAM_FLOATS = 3072;
constexpr uint32_t SMEM_BYTES = R0_BYTES + R1_BYTES + MAX_PARAM_FLOATS * 4 + 2 * TILE_M * 8 + 16 * 8 + 64;
constexpr uint32_t POLICY_MAGIC = 0x4c4f5056u;  // "VPOL"
constexpr uint32_t HEADER_BYTES = 64;
constexpr int TMEM_COLS = 512;

struct Layout {
  int K[6], N[6], ln[6], tcol[6];
  uint32_t lboB[6], offW[6], bytesW[6], offP[6];
  uint32_t offParams, nParamFloats, total;
};

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

__host__ __device__ inline void make_layout(const VnlPolicyDims& d, Layout& L) {
  const int K[6] = {round_up(d.traj, KCHUNK), d.e1, d.e2, round_up(d.latent + d.obs, 16), d.d1, d.d2};
  const int N[6] = {d.e1, d.e2, 2 * d.latent, d.d1, d.d2, round_up(2 * d.nu, 32)};
  const int ln[6] = {1, 1, 0, 1, 1, 0};
  const int tcol[6] = {0, d.e1, d.e1 + d.e2, 0, d.d1, d.d1 + d.d2};
  uint32_t off = HEADER_BYTES, pf = 0;
  for (int i = 0; i < 6; ++i) {
    L.K[i] = K[i];
    L.N[i] = N[i];
    L.ln[i] = ln[i];
    L.tcol[i] = tcol[i];
    L.lboB[i] = (uint32_t)N[i] * 16 + 16;
    L.offW[i] = off;
    L.bytesW[i] = (uint32_t)(K[i] / 8) * L.lboB[i];
    off += L.bytesW[i];
    L.offP[i] = pf;
    pf += (uint32_t)N[i] * (ln[i] ? 3 : 1);
  }
  L.offParams = off;
  L.nParamFloats = pf;
  L.total = (off + pf * 4 + 15u) & ~15u;  // whole 16-byte granules (TMA bulk copies)
}

// K-group offset of the decoder input inside R0: its obs columns must clear the operand of L2 (e2 / 8 K-groups)
__host__ __device__ inline int a3_base_kg(const VnlPolicyDims& d) { return d.e2 / 8 > d.latent / 8 ? d.e2 / 8 - d.latent / 8 : 0; }

// where the action-noise draws of the tile are staged in R1: behind the weights of L4 (live when the copy is issued) and L5
__host__ __device__ inline uint32_t draw_stage_offset(const Layout& L) {
  const uint32_t m = L.bytesW[4] > L.bytesW[5] ? L.bytesW[4] : L.bytesW[5];
  return (m + 127u) & ~127u;
}

int check_dims(const VnlPolicyDims* d) {
  if (!d) return -1;
  const int h[4] = {d->e1, d->e2, d->d1, d->d2};
  for (int i = 0; i < 4; ++i)
    if (h[i] < 64 || h[i] > 256 || h[i] % 64) return -2;
  if (d->latent < 32 || d->latent % 32 || d->latent > 64) return -3;
  if (d->nu < 1 || d->nu > 64 || d->traj < 1 || d->obs < 0 || d->obs > 4 * KCHUNK) return -4;  // obs: two staging phases x two chunks
  Layout L;
  make_layout(*d, L);
  if (d->e1 + d->e2 + 2 * d->latent > TMEM_COLS || d->d1 + d->d2 + L.N[5] > TMEM_COLS) return -5;
  for (int i = 1; i < 6; ++i)
    if ((uint32_t)(L.K[i] / 8 + (i == 3 ? a3_base_kg(*d) : 0)) * LBO_A > R0_BYTES || L.bytesW[i] > R1_BYTES) return -6;
  if ((uint32_t)RING * (KCHUNK / 8) * LBO_A > R0_BYTES || (uint32_t)RING * (KCHUNK / 8) * L.lboB[0] > R1_BYTES) return -6;
  if ((uint32_t)TILE_M * (L.N[5] + 1) * 4 > R0_BYTES || draw_stage_offset(L) + 2u * TILE_M * d->nu * 4 > R1_BYTES) return -6;
  if (L.nParamFloats > (uint32_t)MAX_PARAM_FLOATS) return -7;
  return 0;
}

struct Args {
  VnlPolicyDims d;
  int B;
  const uint8_t* blob;
  const float *traj, *obs, *obs_mean, *obs_std, *eps_z, *eps_a, *rand_action;
  float *action, *raw_action, *logits, *log_prob, *rand_log_prob, *z_mean, *z_logvar;
  int dump_layer;
  float* dump;
};

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// smem matrix descriptor: K-major, no swizzle; LBO = stride between K-adjacent core matrices, SBO = between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// bulk L2 prefetch of a contiguous global range (16-byte granules; the range is shrunk to whole granules)
__device__ __noinline__ void l2_prefetch(const void* p, size_t bytes) {
  const uintptr_t b = ((uintptr_t)p + 15) & ~(uintptr_t)15, e = ((uintptr_t)p + bytes) & ~(uintptr_t)15;
  for (uintptr_t q = b; q < e; q += 65536) {
    const uint32_t n = (uint32_t)((e - q) < 65536 ? (e - q) : 65536);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q), "r"(n) : "memory");
  }
}
__device__ __noinline__ void cp_async_words(uint32_t dst, const float* src, int count, int tid) {
  for (int i = tid; i < count; i += THREADS)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * i), "l"(src + i) : "memory");
}
// TMA bulk copy global -> shared (contiguous, 16-byte granules), completion counted in bytes on an mbarrier
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int LROWS = 19;  // rows per loader warp in layer 0: 128 = 5 x 18 + 2 x 19 over warps 1-7 (warp 0 issues)
#define VNL_R8(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
// split form: the load is issued into `v`, other work proceeds, tmem_ld32_wait(v) makes the registers valid (it lists them
// as in/out operands, so no consumer of v can be scheduled before it)
#define VNL_RW8(v, o) "+r"(v[o + 0]), "+r"(v[o + 1]), "+r"(v[o + 2]), "+r"(v[o + 3]), "+r"(v[o + 4]), "+r"(v[o + 5]), "+r"(v[o + 6]), "+r"(v[o + 7])
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : VNL_R8(v, 0), VNL_R8(v, 8), VNL_R8(v, 16), VNL_R8(v, 24)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : VNL_RW8(v, 0), VNL_RW8(v, 8), VNL_RW8(v, 16), VNL_RW8(v, 24)::"memory");
}
// 16 consecutive accumulator columns, load and wait in one statement
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* f) {
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : VNL_R8(v, 0), VNL_R8(v, 8)
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
}

// Blackwell packed fp32 arithmetic (FADD2 / FMUL2 / FFMA2): two lanes per instruction on an aligned register pair
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// relu(v + b) on a pair
__device__ __forceinline__ f32x2 bias_relu2(uint32_t v0, uint32_t v1, float b0, float b1) {
  float x0, x1;
  upk2(add2(pk2(__uint_as_float(v0), __uint_as_float(v1)), pk2(b0, b1)), x0, x1);
  return pk2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float softplus(float x) { return fmaxf(x, 0.0f) + __logf(1.0f + __expf(-fabsf(x))); }
// tanh(x) and brax TanhBijector.forward_log_det_jacobian(x) = 2 (log 2 - x - softplus(-2x)) from one exponential
__device__ __forceinline__ void tanh_and_log_det(float x, float& th, float& ld) {
  const float t = __expf(-2.0f * fabsf(x));  // in (0, 1]
  th = copysignf(__fdividef(1.0f - t, 1.0f + t), x);
  ld = 2.0f * (0.69314718056f - x - (fmaxf(-2.0f * x, 0.0f) + __logf(1.0f + t)));
}

// ---------------------------------------------------------------------------------------------------------------------
// Coalesced activation loader: NR rows of the tile from r_begin, 64 source columns from col0; lane l owns columns
// col0 + 2l, col0 + 2l + 1 of a row and stores them as one bf16 pair into K-group kg0 + l / 4 of the operand image.
// Out-of-range rows / columns are zeros.  (src rows are only 4-byte aligned: traj rows are 795 floats.)  Split into a
// fetch (all loads in flight at once) and a store so that the fetch of the next chunk can overlap the current MMA.
template <int NR>
struct RowRegs {
  float x[NR][2];
};

template <int NR>
__device__ __forceinline__ void fetch_rows(RowRegs<NR>& rr, const float* __restrict__ src, int ncols, int col0, int row0, int B,
                                           int r_begin, int lane, int count = NR) {
  const int j = col0 + 2 * lane;
  const bool ok0 = j < ncols, ok1 = j + 1 < ncols;
#pragma unroll
  for (int u = 0; u < NR; ++u) {
    const int grow = row0 + r_begin + u;
    const float* p = src + (size_t)grow * ncols + j;
    rr.x[u][0] = (ok0 && grow < B && u < count) ? __ldg(p) : 0.0f;
    rr.x[u][1] = (ok1 && grow < B && u < count) ? __ldg(p + 1) : 0.0f;
  }
}

template <int NR>
__device__ __forceinline__ void store_rows(const RowRegs<NR>& rr, int ncols, int col0, int row0, int B, uint8_t* dst, int kg0,
                                           int kg_limit, int r_begin, const float* __restrict__ mu,
                                           const float* __restrict__ sigma, int lane, int count = NR) {
  const int j = col0 + 2 * lane;
  const bool ok0 = j < ncols, ok1 = j + 1 < ncols;
  const int kg = kg0 + (lane >> 2);
  if (kg >= kg_limit) return;
  float m0 = 0.0f, m1 = 0.0f, s0 = 1.0f, s1 = 1.0f;  // s = 1 / sigma: one division per lane, not one per element
  if (mu) {
    if (ok0) m0 = __ldg(mu + j), s0 = 1.0f / __ldg(sigma + j);
    if (ok1) m1 = __ldg(mu + j + 1), s1 = 1.0f / __ldg(sigma + j + 1);
  }
  uint8_t* out = dst + (uint32_t)kg * LBO_A + (lane & 3) * 4;
#pragma unroll
  for (int u = 0; u < NR; ++u) {
    const int grow = row0 + r_begin + u;
    float a = rr.x[u][0], b = rr.x[u][1];
    if (mu && grow < B) {
      a = ok0 ? (a - m0) * s0 : 0.0f;
      b = ok1 ? (b - m1) * s1 : 0.0f;
    }
    if (u < count) *reinterpret_cast<uint32_t*>(out + (r_begin + u) * 16) = pack_bf16(a, b);
  }
}

// Epilogues run on all 8 warps: warps w and w + 4 share the tensor-memory lanes of quadrant w % 4 (thread = env row
// 32 (w % 4) + lane) and split the accumulator columns in halves.

// +bias, relu, LayerNorm (flax: fast variance E[x^2] - E[x]^2 clipped at 0, eps 1e-6, scale and bias) of this thread's half
// row, written as the bf16 A operand of the next layer.  Two passes over tensor memory instead of a register row; the two
// halves of a row exchange their partial sums through `stats` (one CTA barrier).
__device__ __forceinline__ void epilogue_ln(uint32_t tb, int N, const float* __restrict__ P, uint8_t* anext, int row, int half,
                                            float2* stats) {
  const float* bias = P;
  const float* gamma = P + N;
  const float* beta = P + 2 * N;
  const int cb = half * (N >> 1), nchunk = N >> 6;  // 32-column chunks of this thread's half row
  uint32_t va[32], vb[32];
  f32x2 sum2 = pk2(0.0f, 0.0f), sq2 = pk2(0.0f, 0.0f);  // even / odd columns
  auto stat = [&](const uint32_t* v, int c0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias + c0 + 4 * q);  // warp-uniform address: one broadcast wavefront
      const f32x2 x01 = bias_relu2(v[4 * q], v[4 * q + 1], b4.x, b4.y), x23 = bias_relu2(v[4 * q + 2], v[4 * q + 3], b4.z, b4.w);
      sum2 = add2(sum2, add2(x01, x23));
      sq2 = fma2(x01, x01, fma2(x23, x23, sq2));
    }
  };
  tmem_ld32_issue(tb + cb, va);
  for (int c = 0; c < nchunk; c += 2) {  // the load of chunk c + 1 is in flight while chunk c is reduced
    tmem_ld32_wait(va);
    if (c + 1 < nchunk) tmem_ld32_issue(tb + cb + (c + 1) * 32, vb);
    stat(va, cb + c * 32);
    if (c + 1 < nchunk) {
      tmem_ld32_wait(vb);
      if (c + 2 < nchunk) tmem_ld32_issue(tb + cb + (c + 2) * 32, va);
      stat(vb, cb + (c + 1) * 32);
    }
  }
  tmem_ld32_issue(tb + cb, va);  // first chunk of the second pass, in flight across the barrier
  float sum, sq, sum_hi, sq_hi;
  upk2(sum2, sum, sum_hi);
  upk2(sq2, sq, sq_hi);
  sum += sum_hi;
  sq += sq_hi;
  stats[half * TILE_M + row] = make_float2(sum, sq);
  __syncthreads();
  const float2 p0 = stats[row], p1 = stats[TILE_M + row];
  const float inv_n = 1.0f / (float)N;
  const float mean = (p0.x + p1.x) * inv_n;
  const float var = fmaxf((p0.y + p1.y) * inv_n - mean * mean, 0.0f);
  const float rstd = rsqrtf(var + 1e-6f);
  const f32x2 nmean2 = pk2(-mean, -mean), rstd2 = pk2(rstd, rstd);
  auto emit = [&](const uint32_t* v, int c0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + q * 8 + h * 4;
        const float4 b4 = *reinterpret_cast<const float4*>(bias + c), g4 = *reinterpret_cast<const float4*>(gamma + c),
                     e4 = *reinterpret_cast<const float4*>(beta + c);
        const uint32_t* vv = v + q * 8 + h * 4;
        const f32x2 y01 = fma2(mul2(add2(bias_relu2(vv[0], vv[1], b4.x, b4.y), nmean2), rstd2), pk2(g4.x, g4.y), pk2(e4.x, e4.y));
        const f32x2 y23 = fma2(mul2(add2(bias_relu2(vv[2], vv[3], b4.z, b4.w), nmean2), rstd2), pk2(g4.z, g4.w), pk2(e4.z, e4.w));
        float y0, y1, y2, y3;
        upk2(y01, y0, y1);
        upk2(y23, y2, y3);
        w[2 * h] = pack_bf16(y0, y1);
        w[2 * h + 1] = pack_bf16(y2, y3);
      }
      *reinterpret_cast<uint4*>(anext + (uint32_t)((c0 >> 3) + q) * LBO_A + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  };
  for (int c = 0; c < nchunk; c += 2) {
    tmem_ld32_wait(va);
    if (c + 1 < nchunk) tmem_ld32_issue(tb + cb + (c + 1) * 32, vb);
    emit(va, cb + c * 32);
    if (c + 1 < nchunk) {
      tmem_ld32_wait(vb);
      if (c + 2 < nchunk) tmem_ld32_issue(tb + cb + (c + 2) * 32, va);
      emit(vb, cb + (c + 1) * 32);
    }
  }
}

// [mean | logvar] heads -> z = mean + eps * exp(logvar / 2) (intention_policy_network.py:76-79), K-groups 0 .. latent/8 of
// the decoder's A operand; this thread owns latent columns [half * latent / 2, (half + 1) * latent / 2)
__device__ __forceinline__ void epilogue_z(uint32_t tb, int latent, const float* __restrict__ P, uint8_t* anext, int row,
                                           int half, int grow, int B, const float4* ez, float* z_mean, float* z_logvar) {
  const int cb = half * (latent >> 1);
#pragma unroll
  for (int it = 0; it < 2; ++it) {  // latent / 2 <= 32 columns per thread, 16 at a time
    const int c0 = cb + it * 16;
    if (it * 16 < (latent >> 1)) {
      float m[16], lv[16], e[16];
      tmem_ld16(tb + c0, m);
      tmem_ld16(tb + latent + c0, lv);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        m[j] += P[c0 + j];
        lv[j] += P[latent + c0 + j];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = ez[it * 4 + q];
        e[4 * q] = t.x, e[4 * q + 1] = t.y, e[4 * q + 2] = t.z, e[4 * q + 3] = t.w;
      }
      if (grow < B) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (z_mean)
            reinterpret_cast<float4*>(z_mean + (size_t)grow * latent + c0)[q] = make_float4(m[4 * q], m[4 * q + 1], m[4 * q + 2], m[4 * q + 3]);
          if (z_logvar)
            reinterpret_cast<float4*>(z_logvar + (size_t)grow * latent + c0)[q] = make_float4(lv[4 * q], lv[4 * q + 1], lv[4 * q + 2], lv[4 * q + 3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int j = q * 8 + h * 2;
          w[h] = pack_bf16(m[j] + e[j] * __expf(0.5f * lv[j]), m[j + 1] + e[j + 1] * __expf(0.5f * lv[j + 1]));
        }
        *reinterpret_cast<uint4*>(anext + (uint32_t)((c0 >> 3) + q) * LBO_A + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}

// developer profile (vnl_policy_debug with layer -1): SM-clock stamps of CTA 0 at the phase boundaries, relative to entry
#define VNL_STAMP(i) do { if (prof && tid == 0) a.dump[(i)] = (float)(clock64() - t_entry); } while (0)

__global__ void __launch_bounds__(THREADS, 1) vnl_policy_kernel(const Args a) {
  const bool prof = a.dump && a.dump_layer == -1 && blockIdx.x == 0;
  const long long t_entry = prof ? clock64() : 0;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* R0 = smem;
  uint8_t* R1 = smem + R0_BYTES;
  float* P = reinterpret_cast<float*>(smem + R0_BYTES + R1_BYTES);
  float2* stats = reinterpret_cast<float2*>(P + MAX_PARAM_FLOATS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stats + 2 * TILE_M);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 16);

  // the layer table: a per-thread copy for the compile-time indices (stays in registers) and one in shared memory for the
  // layer loop's run-time indices (a dynamically indexed local array would live on the stack, behind L1)
  __shared__ Layout Ls;
  Layout L;
  make_layout(a.d, L);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) Ls = L;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;  // epilogue role
  const int row0 = blockIdx.x * TILE_M;
  const uint32_t r0_s = smem_u32(R0), r1_s = smem_u32(R1);
  // mbarriers: ring slot released (MMA commit) x3, layer done (MMA commit), ring slot filled (TMA bytes) x3, layer weights
  // filled, bias / LayerNorm vectors filled
  const uint32_t bar_ring = smem_u32(bars), bar_layer = smem_u32(bars + RING), bar_full = smem_u32(bars + RING + 1),
                 bar_wfull = smem_u32(bars + 2 * RING + 1), bar_pfull = smem_u32(bars + 2 * RING + 2),
                 bar_afull = smem_u32(bars + 2 * RING + 3);  // x3: the activations of a ring slot are stored (7 warp arrivals)
  constexpr int PRODUCER = 0;  // thread 0 issues the weight copies and the MMAs
  // layer-0 loader rows of this warp: warps 1-5 take 18 rows, warps 6-7 take 19, warp 0 none (it only issues)
  const int l0_begin = warp == 0 ? 0 : (warp <= 5 ? (warp - 1) * 18 : 90 + (warp - 6) * 19);
  const int l0_count = warp == 0 ? 0 : (warp <= 5 ? 18 : 19);

  // The tile's rows of every input are one contiguous range: ask for them in L2 now, so that the dependent phases below see
  // L2 latency instead of HBM latency.
  if (tid == 0) {
    const int rows = min(TILE_M, a.B - row0);
    l2_prefetch(a.traj + (size_t)row0 * a.d.traj, (size_t)rows * a.d.traj * 4);
    if (a.d.obs > 0) l2_prefetch(a.obs + (size_t)row0 * a.d.obs, (size_t)rows * a.d.obs * 4);
    l2_prefetch(a.eps_z + (size_t)row0 * a.d.latent, (size_t)rows * a.d.latent * 4);
    if (a.eps_a) l2_prefetch(a.eps_a + (size_t)row0 * a.d.nu, (size_t)rows * a.d.nu * 4);
    if (a.rand_action) l2_prefetch(a.rand_action + (size_t)row0 * a.d.nu, (size_t)rows * a.d.nu * 4);
  }

  // ---- L0 prologue: the first two weight chunks and the first activation chunk are requested before anything else ----
  const int nch = L.K[0] / KCHUNK;
  const uint32_t a_stride = (KCHUNK / 8) * LBO_A, b_stride = (KCHUNK / 8) * L.lboB[0];
  const uint8_t* w0 = a.blob + L.offW[0];
  RowRegs<LROWS> rr, rr2;  // activations of the even / odd chunks, fetched two iterations ahead
  if (tid == PRODUCER) {
    for (int i = 0; i < 2 * RING + 3; ++i) mbar_init(bar_ring + 8 * i, 1);
    for (int i = 0; i < RING; ++i) mbar_init(bar_afull + 8 * i, 7);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t pbytes = (L.nParamFloats * 4 + 15) & ~15u;
    mbar_expect_tx(bar_pfull, pbytes);
    bulk_g2s(smem_u32(P), a.blob + L.offParams, pbytes, bar_pfull);  // bias / LayerNorm vectors
    for (int c = 0; c < 2 && c < nch; ++c) {
      mbar_expect_tx(bar_full + 8 * c, b_stride);
      bulk_g2s(r1_s + c * b_stride, w0 + (size_t)c * b_stride, b_stride, bar_full + 8 * c);
    }
  }
  if (warp > 0) {
    fetch_rows<LROWS>(rr, a.traj, a.d.traj, 0, row0, a.B, l0_begin, lane, l0_count);
    if (nch > 1) fetch_rows<LROWS>(rr2, a.traj, a.d.traj, KCHUNK, row0, a.B, l0_begin, lane, l0_count);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  VNL_STAMP(0);

  // ---- L0: K streamed in 64-wide chunks through a 3-slot ring (A slots in R0, B slots in R1), two decoupled loops:
  // warps 1-7 convert and store the activations of chunk c (those of chunk c + 2 in flight in registers) and arrive on the
  // slot's `afull` barrier; thread 0 waits for the activations and the TMA'd weights of chunk c, issues its MMAs (4 x ~130
  // cycles that no loader waits for), and refills the weight slot of chunk c - 1 with chunk c + 2. ----
  if (warp > 0) {
    auto step = [&](int c, RowRegs<LROWS>& regs) {
      const int slot = c % RING;
      if (c >= RING) mbar_wait(bar_ring + 8 * slot, ((c - RING) / RING) & 1);  // the MMAs of chunk c - 3 have left this A slot
      store_rows<LROWS>(regs, a.d.traj, c * KCHUNK, row0, a.B, R0 + slot * a_stride, 0, KCHUNK / 8, l0_begin, nullptr, nullptr, lane,
                        l0_count);
      if (c + 2 < nch) fetch_rows<LROWS>(regs, a.traj, a.d.traj, (c + 2) * KCHUNK, row0, a.B, l0_begin, lane, l0_count);
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_afull + 8 * slot);
    };
    for (int c = 0; c < nch; c += 2) {
      step(c, rr);
      if (c + 1 < nch) step(c + 1, rr2);
    }
  } else if (tid == 0) {
    const uint32_t idesc = make_idesc(L.N[0]), d_tmem = tmem + L.tcol[0], lbo_b = L.lboB[0];
    const uint64_t sa = (2 * LBO_A) >> 4, sb = (2 * lbo_b) >> 4;
    for (int c = 0; c < nch; ++c) {
      const int slot = c % RING;
      mbar_wait(bar_afull + 8 * slot, (c / RING) & 1);  // activations stored by the 7 loader warps
      mbar_wait(bar_full + 8 * slot, (c / RING) & 1);   // weights landed (TMA)
      tc_fence_after();
      uint64_t da = make_desc(r0_s + slot * a_stride, LBO_A), db = make_desc(r1_s + slot * b_stride, lbo_b);
#pragma unroll
      for (int s = 0; s < KCHUNK / 16; ++s, da += sa, db += sb) mma_bf16(d_tmem, da, db, idesc, (c > 0 || s > 0) ? 1u : 0u);
      mma_commit(bar_ring + 8 * slot);
      if (c == nch - 1) mma_commit(bar_layer);
      if (c + 2 < nch) {
        const int ns = (c + 2) % RING;  // = slot of chunk c - 1: refill once its MMAs are done
        if (c >= 1) mbar_wait(bar_ring + 8 * ns, ((c - 1) / RING) & 1);
        mbar_expect_tx(bar_full + 8 * ns, b_stride);
        bulk_g2s(r1_s + ns * b_stride, w0 + (size_t)(c + 2) * b_stride, b_stride, bar_full + 8 * ns);
      }
      if (c < 4) VNL_STAMP(1 + c);
    }
  }
  __syncwarp();
  VNL_STAMP(5);

  // ---- L1 .. L5: the weights of layer n stream in (cp.async) while all warps run the epilogue of layer n - 1.
  // The decoder input [z | normalised obs | 0] sits a3_kg K-groups into R0 so that its obs columns can be staged early
  // (first half beside the epilogue of L1, second half beside the z epilogue) without touching the live operand of L2.
  const int kg_lat = a.d.latent / 8, kg_end3 = L.K[3] / 8;
  const int a3_kg = a3_base_kg(a.d);
  uint8_t* A3 = R0 + (uint32_t)a3_kg * LBO_A;
  const int nobs = (L.K[3] - a.d.latent + KCHUNK - 1) / KCHUNK;
  float4 ez[8];
  auto obs_fetch = [&](RowRegs<LROWS>& r, int ci) { fetch_rows<LROWS>(r, a.obs, a.d.obs, ci * KCHUNK, row0, a.B, warp * 16, lane, 16); };
  auto obs_store = [&](const RowRegs<LROWS>& r, int ci) {
    store_rows<LROWS>(r, a.d.obs, ci * KCHUNK, row0, a.B, A3, kg_lat + ci * (KCHUNK / 8), kg_end3, warp * 16, a.obs_mean, a.obs_std, lane, 16);
  };
  uint32_t layer_phase = 0;
  for (int n = 1; n < 6; ++n) {
    // global operands of the coming epilogue are requested before the wait on the MMAs of layer n - 1
    const int ob = (n == 2) ? 0 : nobs / 2, oe = (n == 2) ? nobs / 2 : ((n == 3) ? nobs : 0);
    if (n == 2 || n == 3) {
      if (ob < oe) obs_fetch(rr, ob);
      if (ob + 1 < oe) obs_fetch(rr2, ob + 1);
    }
    if (n == 3) {
      const int hw = a.d.latent >> 1, grow = row0 + row;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        ez[q] = (4 * q < hw && grow < a.B) ? __ldg(reinterpret_cast<const float4*>(a.eps_z + (size_t)grow * a.d.latent + half * hw) + q)
                                           : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    if (n == 5) {  // the tile's action-noise draws -> shared memory behind the weights (consumed after L5)
      const int cnt = min(TILE_M, a.B - row0) * a.d.nu;
      const uint32_t dst = r1_s + draw_stage_offset(L);
      if (a.eps_a) cp_async_words(dst, a.eps_a + (size_t)row0 * a.d.nu, cnt, tid);
      if (a.rand_action) cp_async_words(dst + TILE_M * a.d.nu * 4, a.rand_action + (size_t)row0 * a.d.nu, cnt, tid);
    }
    mbar_wait(bar_layer, layer_phase);
    layer_phase ^= 1;
    tc_fence_after();
    VNL_STAMP(6 + 3 * (n - 1));
    if (tid == PRODUCER) {  // the weights of layer n: one TMA bulk copy (R1 is free: the MMAs of layer n - 1 are done)
      mbar_expect_tx(bar_wfull, Ls.bytesW[n]);
      bulk_g2s(r1_s, a.blob + Ls.offW[n], Ls.bytesW[n], bar_wfull);
    }
    if (n == 1) mbar_wait(bar_pfull, 0);  // bias / LayerNorm vectors (landed long ago; the wait is the acquire)
    cp_async_commit();
    const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + Ls.tcol[n - 1];
    if (a.dump && a.dump_layer == n - 1 && blockIdx.x == 0) {
      const int hw = Ls.N[n - 1] >> 1;
      for (int c0 = half * hw; c0 < (half + 1) * hw; c0 += 16) {
        float v[16];
        tmem_ld16(tb + c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) a.dump[(size_t)row * Ls.N[n - 1] + c0 + j] = v[j];
      }
    }
    if (n == 3)
      epilogue_z(tb, a.d.latent, P + L.offP[2], A3, row, half, row0 + row, a.B, ez, a.z_mean, a.z_logvar);
    else
      epilogue_ln(tb, Ls.N[n - 1], P + Ls.offP[n - 1], R0, row, half, stats);
    if (n == 2 || n == 3) {
      if (ob < oe) obs_store(rr, ob);
      if (ob + 1 < oe) obs_store(rr2, ob + 1);
    }
    VNL_STAMP(7 + 3 * (n - 1));
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    VNL_STAMP(8 + 3 * (n - 1));
    if (tid == 0) {
      mbar_wait(bar_wfull, (n - 1) & 1);
      tc_fence_after();
      // everything the issue loop needs sits in registers before it starts (the asm statements clobber memory, so a table
      // read inside the loop would be a shared-memory round trip per MMA); the descriptors advance by constant increments
      const uint32_t idesc = make_idesc(Ls.N[n]), lbo_b = Ls.lboB[n], d_tmem = tmem + Ls.tcol[n];
      const int ksteps = Ls.K[n] / 16;
      uint64_t da = make_desc(r0_s + (n == 3 ? (uint32_t)a3_kg * LBO_A : 0u), LBO_A), db = make_desc(r1_s, lbo_b);
      const uint64_t sa = (2 * LBO_A) >> 4, sb = (2 * lbo_b) >> 4;  // one K step of 16 = two K-groups, in 16-byte units
      for (int s = 0; s < ksteps; ++s, da += sa, db += sb) mma_bf16(d_tmem, da, db, idesc, s > 0 ? 1u : 0u);
      mma_commit(bar_layer);
      VNL_STAMP(23 + n);
    }
  }

  // ---- logits -> NormalTanhDistribution sample / log-prob (ppo_networks.py:45-83).  Warp w owns tile rows w, w + 8, ..;
  // lane i owns action dimensions i (and i + 32), so a row's log-prob is one warp reduction (fixed order).  The loops are
  // kept rolled (4 rows per trip): this code runs once per CTA and its instruction-fetch footprint is what it costs. ----
  const int nu = a.d.nu, nlog = 2 * nu, sstride = L.N[5] + 1;
  mbar_wait(bar_layer, layer_phase);
  tc_fence_after();
  VNL_STAMP(21);
  float* S = reinterpret_cast<float*>(R0);  // [128][N5 + 1] logits
  const float* EA = reinterpret_cast<const float*>(R1 + draw_stage_offset(L));  // [128][nu] eps_a, then [128][nu] rand_action
  const float* UA = EA + TILE_M * nu;
  {
    const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + L.tcol[5];
    const float* bias = P + L.offP[5];
    const int hw = L.N[5] >> 1;
#pragma unroll 1
    for (int c0 = half * hw; c0 < (half + 1) * hw; c0 += 16) {
      float v[16];
      tmem_ld16(tb + c0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (a.dump && a.dump_layer == 5 && blockIdx.x == 0) a.dump[(size_t)row * L.N[5] + c0 + j] = v[j];
        S[row * sstride + c0 + j] = v[j] + bias[c0 + j];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  constexpr int RG = 8;  // rows of a warp in flight per trip (16 rows per warp: two trips)
#pragma unroll 1
  for (int g = 0; g < TILE_M / (8 * RG); ++g) {
    float lp[RG], lpr[RG];
#pragma unroll
    for (int u = 0; u < RG; ++u) lp[u] = lpr[u] = 0.0f;
#pragma unroll 1
    for (int i = lane; i < nu; i += 32) {
      float th[RG], raw[RG];
#pragma unroll
      for (int u = 0; u < RG; ++u) {  // independent rows in flight, no stores in between
        const int r = warp + 8 * (RG * g + u);
        const float loc = S[r * sstride + i];
        const float scale = softplus(S[r * sstride + nu + i]) + 1e-3f;  // brax NormalTanhDistribution min_std
        const float e = a.eps_a ? EA[r * nu + i] : 0.0f;
        raw[u] = loc + scale * e;
        const float log_scale = __logf(scale);
        float ld;
        tanh_and_log_det(raw[u], th[u], ld);
        lp[u] += -0.5f * e * e - log_scale - 0.91893853321f - ld;
        if (a.rand_action) {
          const float ua = UA[r * nu + i], zr = __fdividef(ua - loc, scale);
          float t2;
          tanh_and_log_det(ua, t2, ld);
          lpr[u] += -0.5f * zr * zr - log_scale - 0.91893853321f - ld;
        }
      }
#pragma unroll
      for (int u = 0; u < RG; ++u) {
        const int grow = row0 + warp + 8 * (RG * g + u);
        if (grow < a.B) {
          if (a.action) a.action[(size_t)grow * nu + i] = th[u];
          if (a.raw_action) a.raw_action[(size_t)grow * nu + i] = raw[u];
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < RG; ++u) {
        lp[u] += __shfl_xor_sync(0xffffffffu, lp[u], o);
        lpr[u] += __shfl_xor_sync(0xffffffffu, lpr[u], o);
      }
    // lane u keeps row u's sums, so that the scalar stores below are one predicated store per array
    float l = 0.0f, lr = 0.0f;
#pragma unroll
    for (int u = 0; u < RG; ++u)
      if (lane == u) l = lp[u], lr = lpr[u];
    if (lane < RG) {
      const int grow = row0 + warp + 8 * (RG * g + lane);
      if (grow < a.B) {
        if (a.log_prob) a.log_prob[grow] = l;
        if (a.rand_action && a.rand_log_prob) a.rand_log_prob[grow] = lr;
      }
    }
  }
  if (a.logits) {  // [128, 2 nu] row-major out of the scratch: a warp per row, coalesced
#pragma unroll 1
    for (int k = 0; k < TILE_M / 8; ++k) {
      const int r = warp + 8 * k, grow = row0 + r;
      if (grow < a.B)
#pragma unroll 2
        for (int i = lane; i < nlog; i += 32) a.logits[(size_t)grow * nlog + i] = S[r * sstride + i];
    }
  }
  __syncthreads();
  VNL_STAMP(22);
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

uint16_t bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// W [K_true, N_true] row-major (flax kernel) at output-column offset n0 of layer `li` -> operand image rows n0 .. n0 + N_true
void pack_weight(const Layout& L, int li, const float* W, int K_true, int N_true, int n0, int k0, uint8_t* blob) {
  for (int k = 0; k < K_true; ++k)
    for (int n = 0; n < N_true; ++n) {
      const int kk = k0 + k;
      uint16_t* dst = reinterpret_cast<uint16_t*>(blob + L.offW[li] + (size_t)(kk / 8) * L.lboB[li] + (size_t)(n0 + n) * 16) + (kk % 8);
      *dst = bf16_rne(W[(size_t)k * N_true + n]);
    }
}

int launch(const Args& a, void* stream) {
  static bool attr_set[64] = {};  // per device; idempotent, a race sets it twice
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -9;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(vnl_policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return -(int)e;
    attr_set[dev] = true;
  }
  const int grid = (a.B + TILE_M - 1) / TILE_M;
  vnl_policy_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(a);
  return -(int)cudaGetLastError();
}

}  // namespace

extern "C" {

int vnl_policy_check(const VnlPolicyDims* dims) { return check_dims(dims); }

size_t vnl_policy_blob_bytes(const VnlPolicyDims* dims) {
  if (check_dims(dims)) return 0;
  Layout L;
  make_layout(*dims, L);
  return L.total;
}

int vnl_policy_pack(const VnlPolicyDims* dims, const float* const* p, void* blob_host, size_t nbytes) {
  const int rc = check_dims(dims);
  if (rc) return rc;
  if (!p || !blob_host) return -1;
  Layout L;
  make_layout(*dims, L);
  if (nbytes < L.total) return -8;
  uint8_t* blob = static_cast<uint8_t*>(blob_host);
  memset(blob, 0, L.total);
  uint32_t hdr[16] = {POLICY_MAGIC, 1u};
  memcpy(hdr + 2, dims, sizeof(VnlPolicyDims));
  memcpy(blob, hdr, sizeof(hdr));
  const VnlPolicyDims& d = *dims;
  float* P = reinterpret_cast<float*>(blob + L.offParams);
  auto vec = [&](int li, int slot, const float* v, int n, int off) { memcpy(P + L.offP[li] + slot * L.N[li] + off, v, n * sizeof(float)); };
  // encoder
  pack_weight(L, 0, p[0], d.traj, d.e1, 0, 0, blob);
  vec(0, 0, p[1], d.e1, 0), vec(0, 1, p[2], d.e1, 0), vec(0, 2, p[3], d.e1, 0);
  pack_weight(L, 1, p[4], d.e1, d.e2, 0, 0, blob);
  vec(1, 0, p[5], d.e2, 0), vec(1, 1, p[6], d.e2, 0), vec(1, 2, p[7], d.e2, 0);
  pack_weight(L, 2, p[8], d.e2, d.latent, 0, 0, blob);
  pack_weight(L, 2, p[10], d.e2, d.latent, d.latent, 0, blob);
  vec(2, 0, p[9], d.latent, 0), vec(2, 0, p[11], d.latent, d.latent);
  // decoder
  pack_weight(L, 3, p[12], d.latent + d.obs, d.d1, 0, 0, blob);
  vec(3, 0, p[13], d.d1, 0), vec(3, 1, p[14], d.d1, 0), vec(3, 2, p[15], d.d1, 0);
  pack_weight(L, 4, p[16], d.d1, d.d2, 0, 0, blob);
  vec(4, 0, p[17], d.d2, 0), vec(4, 1, p[18], d.d2, 0), vec(4, 2, p[19], d.d2, 0);
  pack_weight(L, 5, p[20], d.d2, 2 * d.nu, 0, 0, blob);
  vec(5, 0, p[21], 2 * d.nu, 0);
  return 0;
}

int vnl_policy_forward(const void* blob_dev, const VnlPolicyDims* dims, int B, const float* traj, const float* obs,
                       const float* obs_mean, const float* obs_std, const float* eps_z, const float* eps_a,
                       const float* rand_action, float* action, float* raw_action, float* logits, float* log_prob,
                       float* rand_log_prob, float* z_mean, float* z_logvar, void* stream) {
  const int rc = check_dims(dims);
  if (rc) return rc;
  if (B == 0) return 0;  // an empty batch launches nothing (its buffers may be null)
  if (!blob_dev || !traj || !eps_z || B < 0 || (dims->obs > 0 && !obs) || ((obs_mean == nullptr) != (obs_std == nullptr))) return -1;
  if (reinterpret_cast<uintptr_t>(blob_dev) & 15) return -10;  // TMA bulk copies stream the blob in 16-byte granules
  Args a{*dims, B, static_cast<const uint8_t*>(blob_dev), traj, obs, obs_mean, obs_std, eps_z, eps_a, rand_action,
         action, raw_action, logits, log_prob, rand_log_prob, z_mean, z_logvar, -1, nullptr};
  return launch(a, stream);
}

int vnl_policy_debug(const void* blob_dev, const VnlPolicyDims* dims, int B, const float* traj, const float* obs,
                     const float* obs_mean, const float* obs_std, const float* eps_z, int layer, float* dump, void* stream) {
  const int rc = check_dims(dims);
  if (rc) return rc;
  if (!blob_dev || !traj || !eps_z || !dump || B <= 0 || layer < -1 || layer > 5) return -1;
  if (reinterpret_cast<uintptr_t>(blob_dev) & 15) return -10;
  Args a{*dims, B, static_cast<const uint8_t*>(blob_dev), traj, obs, obs_mean, obs_std, eps_z, nullptr, nullptr,
         nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, layer, dump};
  return launch(a, stream);
}

// Legacy XLA custom call (`void f(cudaStream_t, void** buffers, const char* opaque, size_t opaque_len)`).
// opaque = 9 little-endian int32: the VnlPolicyDims fields (traj, obs, latent, e1, e2, d1, d2, nu), then B.
// buffers: [blob, traj, obs, obs_mean, obs_std, eps_z, eps_a, rand_action,
//           (outputs) action, raw_action, logits, log_prob, rand_log_prob]
void vnl_xla_policy_forward(void* stream, void** b, const char* opaque, size_t opaque_len, void* status) {
  if (!b || !opaque || opaque_len < 9 * sizeof(int32_t)) { vnl::xla_report(status, "vnl_xla_policy_forward", -30); return; }
  VnlPolicyDims d;
  int32_t B;
  memcpy(&d, opaque, sizeof(d));
  memcpy(&B, opaque + sizeof(d), sizeof(B));
  const int rc = vnl_policy_forward(b[0], &d, B, (const float*)b[1], (const float*)b[2], (const float*)b[3], (const float*)b[4],
                                    (const float*)b[5], (const float*)b[6], (const float*)b[7], (float*)b[8], (float*)b[9],
                                    (float*)b[10], (float*)b[11], (float*)b[12], nullptr, nullptr, stream);
  vnl::xla_report(status, "vnl_xla_policy_forward", rc);
}

}  // extern "C"

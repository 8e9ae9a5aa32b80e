// Intention-network policy forward: the step on the other side of env.step in the rollout
// (ppo_imitation/acting.py:47-48 -> ppo_networks.py:45-83 -> intention_policy_network.py:20-105), one launch.
//
// One CTA per 128 envs (tile M = 128 = the tcgen05 accumulator height: env r of the tile is tensor-memory lane r).
// Six dense layers run as tcgen05.mma.cta_group::1.kind::f16 (bf16 operands from shared memory, fp32 accumulators in
// tensor memory); everything between two layers happens in the epilogue of the first, on the SM:
//
//   L0  traj[795 -> 832]  x W[.,256]  -> +b, relu, LayerNorm -> bf16 A-operand of L1     (K streamed in 64-wide chunks,
//   L1  [256] x W[.,128]              -> +b, relu, LayerNorm                               2-deep ring, loads overlap MMA)
//   L2  [128] x [W_mean | W_logvar]   -> +b, z = mean + eps_z * exp(logvar / 2)
//   L3  [z | (obs - mu) / sigma] (296 -> 304) x W[.,128] -> +b, relu, LayerNorm
//   L4  [128] x W[.,256]              -> +b, relu, LayerNorm
//   L5  [256] x W[.,60 -> 64]         -> +b = logits; tanh-normal sample, log-prob, outputs
//
// Operand image (both A and B, K-major, no swizzle): 8-row x 16-byte core matrices; element (r, k) of an R-row operand
// sits at (k / 8) * LBO + r * 16 + (k % 8) * 2 with LBO = R * 16 + 16.  The 16 bytes of slack per K-group rotate the
// banks so that the coalesced activation loader (a lane owns 2 consecutive k of one row) stores conflict-free; the
// stride between 8-row groups (SBO) is 128.  Weights are packed into exactly this image on the host (vnl_policy_pack),
// so a layer's B operand is one contiguous cp.async stream.
//
// Warps 0-3 own the epilogues (thread t = env t of the tile = its tensor-memory lane, 32x32b loads), warps 4-7
// stream the next layer's weights (and the normalised obs columns of L3's A operand) while the epilogue runs; thread 0
// issues the MMAs and commits them to an mbarrier.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/vnl_policy.h"

namespace {

constexpr int TILE_M = 128;
constexpr int THREADS = 256;
constexpr int KCHUNK = 64;
constexpr uint32_t LBO_A = TILE_M * 16 + 16;
constexpr uint32_t SBO = 128;
constexpr uint32_t R0_BYTES = 80 * 1024;  // A operands (activations), later the logits scratch
constexpr uint32_t R1_BYTES = 80 * 1024;  // B operands (weights), later the log-prob scratch
constexpr int MAX_PARAM_FLOATS = 3072;
constexpr uint32_t SMEM_BYTES = R0_BYTES + R1_BYTES + MAX_PARAM_FLOATS * 4 + 64;
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
  L.total = off + pf * 4;
}

int check_dims(const VnlPolicyDims* d) {
  if (!d) return -1;
  const int h[4] = {d->e1, d->e2, d->d1, d->d2};
  for (int i = 0; i < 4; ++i)
    if (h[i] < 32 || h[i] > 256 || h[i] % 32) return -2;
  if (d->latent < 16 || d->latent % 16 || 2 * d->latent > 256) return -3;
  if (d->nu < 1 || d->nu > 64 || d->traj < 1 || d->obs < 0) return -4;
  Layout L;
  make_layout(*d, L);
  if (d->e1 + d->e2 + 2 * d->latent > TMEM_COLS || d->d1 + d->d2 + L.N[5] > TMEM_COLS) return -5;
  for (int i = 1; i < 6; ++i)
    if ((uint32_t)(L.K[i] / 8) * LBO_A > R0_BYTES || L.bytesW[i] > R1_BYTES) return -6;
  if (2u * (KCHUNK / 8) * LBO_A > R0_BYTES || 2u * (KCHUNK / 8) * L.lboB[0] > R1_BYTES) return -6;
  if ((uint32_t)TILE_M * (L.N[5] + 1) * 4 > R0_BYTES || (uint32_t)TILE_M * d->nu * 8 > R1_BYTES) return -6;
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
  int desc_mode;  // developer knob (VNL_POLICY_DESC_MODE): 1 swaps the two descriptor strides
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
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, int mode = 0) {
  const uint32_t lead = mode ? SBO : lbo, stride = mode ? lbo : SBO;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lead >> 4) << 16) | ((uint64_t)(stride >> 4) << 32) | (1ull << 46);
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
__device__ __forceinline__ void cp_async_bytes(uint32_t dst, const uint8_t* src, uint32_t bytes, int tid, int nthreads) {
  for (uint32_t o = (uint32_t)tid * 16; o < bytes; o += (uint32_t)nthreads * 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

#define VNL_R8(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
// 32 (16) consecutive accumulator columns of this thread's tensor-memory lane; the wait is part of the statement so
// that no consumer can be scheduled between the load and its completion
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* f) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : VNL_R8(v, 0), VNL_R8(v, 8), VNL_R8(v, 16), VNL_R8(v, 24)
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
}
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

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float softplus(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }
// brax TanhBijector.forward_log_det_jacobian
__device__ __forceinline__ float tanh_log_det(float x) { return 2.0f * (0.69314718056f - x - softplus(-2.0f * x)); }

// ---------------------------------------------------------------------------------------------------------------------
// Coalesced activation loader: rows [r_begin, r_end) of the tile, 64 source columns from col0; lane l owns columns
// col0 + 2l, col0 + 2l + 1 of a row and stores them as one bf16 pair into K-group kg0 + l / 4 of the operand image.
// Out-of-range rows / columns store zeros.  (src rows are only 4-byte aligned: traj rows are 795 floats.)
__device__ __forceinline__ void load_rows(const float* __restrict__ src, int ncols, int col0, int row0, int B, uint8_t* dst,
                                          int kg0, int kg_limit, int r_begin, int r_end, const float* __restrict__ mu,
                                          const float* __restrict__ sigma, int lane) {
  const int j = col0 + 2 * lane;
  const bool ok0 = j < ncols, ok1 = j + 1 < ncols;
  const int kg = kg0 + (lane >> 2);
  if (kg >= kg_limit) return;
  float m0 = 0.0f, m1 = 0.0f, s0 = 1.0f, s1 = 1.0f;
  if (mu) {
    if (ok0) m0 = __ldg(mu + j), s0 = __ldg(sigma + j);
    if (ok1) m1 = __ldg(mu + j + 1), s1 = __ldg(sigma + j + 1);
  }
  uint8_t* out = dst + (uint32_t)kg * LBO_A + (lane & 3) * 4;
  for (int r = r_begin; r < r_end; r += 4) {
    float x[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int grow = row0 + r + u;
      const float* p = src + (size_t)grow * ncols + j;
      x[u][0] = (ok0 && grow < B) ? __ldg(p) : 0.0f;
      x[u][1] = (ok1 && grow < B) ? __ldg(p + 1) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int grow = row0 + r + u;
      float a = x[u][0], b = x[u][1];
      if (mu && grow < B) {
        a = ok0 ? (a - m0) / s0 : 0.0f;
        b = ok1 ? (b - m1) / s1 : 0.0f;
      }
      *reinterpret_cast<uint32_t*>(out + (r + u) * 16) = pack_bf16(a, b);
    }
  }
}

// +bias, relu, LayerNorm (flax: fast variance E[x^2] - E[x]^2 clipped at 0, eps 1e-6, scale and bias) of this thread's row,
// written as the bf16 A operand of the next layer.  Two passes over tensor memory instead of a 256-float register row.
__device__ __forceinline__ void epilogue_ln(uint32_t tb, int N, const float* __restrict__ P, uint8_t* anext, int row) {
  const float* bias = P;
  const float* gamma = P + N;
  const float* beta = P + 2 * N;
  float sum = 0.0f, sq = 0.0f;
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tb + c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x = fmaxf(v[j] + bias[c0 + j], 0.0f);
      sum += x;
      sq = fmaf(x, x, sq);
    }
  }
  const float inv_n = 1.0f / (float)N;
  const float mean = sum * inv_n;
  const float var = fmaxf(sq * inv_n - mean * mean, 0.0f);
  const float rstd = rsqrtf(var + 1e-6f);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tb + c0, v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int c = c0 + q * 8 + h * 2;
        const float x0 = fmaxf(v[q * 8 + h * 2] + bias[c], 0.0f), x1 = fmaxf(v[q * 8 + h * 2 + 1] + bias[c + 1], 0.0f);
        w[h] = pack_bf16((x0 - mean) * rstd * gamma[c] + beta[c], (x1 - mean) * rstd * gamma[c + 1] + beta[c + 1]);
      }
      *reinterpret_cast<uint4*>(anext + (uint32_t)((c0 >> 3) + q) * LBO_A + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// [mean | logvar] heads -> z = mean + eps * exp(logvar / 2) (intention_policy_network.py:76-79), K-groups 0 .. latent/8 of
// the decoder's A operand
__device__ __forceinline__ void epilogue_z(uint32_t tb, int latent, const float* __restrict__ P, uint8_t* anext, int row,
                                           int grow, int B, const float* __restrict__ eps_z, float* z_mean, float* z_logvar) {
  for (int c0 = 0; c0 < latent; c0 += 16) {
    float m[16], lv[16], e[16];
    tmem_ld16(tb + c0, m);
    tmem_ld16(tb + latent + c0, lv);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      m[j] += P[c0 + j];
      lv[j] += P[latent + c0 + j];
      e[j] = 0.0f;
    }
    if (grow < B) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(eps_z + (size_t)grow * latent + c0) + q);
        e[4 * q] = t.x, e[4 * q + 1] = t.y, e[4 * q + 2] = t.z, e[4 * q + 3] = t.w;
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
        w[h] = pack_bf16(m[j] + e[j] * expf(0.5f * lv[j]), m[j + 1] + e[j + 1] * expf(0.5f * lv[j + 1]));
      }
      *reinterpret_cast<uint4*>(anext + (uint32_t)((c0 >> 3) + q) * LBO_A + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) vnl_policy_kernel(const Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* R0 = smem;
  uint8_t* R1 = smem + R0_BYTES;
  float* P = reinterpret_cast<float*>(smem + R0_BYTES + R1_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(P + MAX_PARAM_FLOATS);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);

  Layout L;
  make_layout(a.d, L);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TILE_M;
  const uint32_t r0_s = smem_u32(R0), r1_s = smem_u32(R1);
  const uint32_t bar_ring[2] = {smem_u32(bars), smem_u32(bars + 1)};
  const uint32_t bar_layer = smem_u32(bars + 2);

  {
    const float* src = reinterpret_cast<const float*>(a.blob + L.offParams);
    for (uint32_t i = tid; i < L.nParamFloats; i += THREADS) P[i] = __ldg(src + i);
  }
  if (tid == 0) {
    mbar_init(bar_ring[0], 1);
    mbar_init(bar_ring[1], 1);
    mbar_init(bar_layer, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  // ---- L0: K streamed in 64-wide chunks through a 2-deep ring (A halves of R0, B halves of R1) ----
  {
    const int nch = L.K[0] / KCHUNK;
    const uint32_t a_stride = (KCHUNK / 8) * LBO_A, b_stride = (KCHUNK / 8) * L.lboB[0];
    const uint32_t idesc = make_idesc(L.N[0]);
    uint32_t ring_phase[2] = {0, 0};
    for (int c = 0; c < nch; ++c) {
      const int b = c & 1;
      if (c >= 2) {  // the MMAs of chunk c - 2 have finished reading this slot
        mbar_wait(bar_ring[b], ring_phase[b]);
        ring_phase[b] ^= 1;
      }
      cp_async_bytes(r1_s + b * b_stride, a.blob + L.offW[0] + (size_t)c * b_stride, b_stride, tid, THREADS);
      load_rows(a.traj, a.d.traj, c * KCHUNK, row0, a.B, R0 + b * a_stride, 0, KCHUNK / 8, warp * 16, warp * 16 + 16, nullptr,
                nullptr, lane);
      cp_async_wait_all();
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < KCHUNK / 16; ++s)
          mma_bf16(tmem + L.tcol[0], make_desc(r0_s + b * a_stride + s * 2 * LBO_A, LBO_A, a.desc_mode),
                   make_desc(r1_s + b * b_stride + s * 2 * L.lboB[0], L.lboB[0], a.desc_mode), idesc, (c > 0 || s > 0) ? 1u : 0u);
        mma_commit(bar_ring[b]);
        if (c == nch - 1) mma_commit(bar_layer);
      }
    }
  }

  // ---- L1 .. L5: epilogue of layer n - 1 (warps 0-3) beside the weight stream of layer n (warps 4-7) ----
  uint32_t layer_phase = 0;
  for (int n = 1; n < 6; ++n) {
    mbar_wait(bar_layer, layer_phase);
    layer_phase ^= 1;
    tc_fence_after();
    if (warp >= 4) {
      cp_async_bytes(r1_s, a.blob + L.offW[n], L.bytesW[n], tid - 128, 128);
      if (n == 3) {  // decoder input = [z | normalised obs | 0]: the obs columns and the zero tail
        const int w = warp - 4, kg_lat = a.d.latent / 8, kg_end = L.K[3] / 8;
        for (int col0 = 0; kg_lat + col0 / 8 < kg_end; col0 += KCHUNK)
          load_rows(a.obs, a.d.obs, col0, row0, a.B, R0, kg_lat + col0 / 8, kg_end, w * 32, w * 32 + 32, a.obs_mean, a.obs_std, lane);
      }
      cp_async_wait_all();
      fence_proxy_async();
    } else {
      const uint32_t tb = tmem + ((uint32_t)(warp * 32) << 16) + L.tcol[n - 1];
      if (a.dump && a.dump_layer == n - 1 && blockIdx.x == 0) {
        for (int c0 = 0; c0 < L.N[n - 1]; c0 += 16) {
          float v[16];
          tmem_ld16(tb + c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) a.dump[(size_t)tid * L.N[n - 1] + c0 + j] = v[j];
        }
      }
      if (L.ln[n - 1])
        epilogue_ln(tb, L.N[n - 1], P + L.offP[n - 1], R0, tid);
      else
        epilogue_z(tb, a.d.latent, P + L.offP[n - 1], R0, tid, row0 + tid, a.B, a.eps_z, a.z_mean, a.z_logvar);
      fence_proxy_async();
      tc_fence_before();
    }
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc(L.N[n]);
      const int ksteps = L.K[n] / 16;
      for (int s = 0; s < ksteps; ++s)
        mma_bf16(tmem + L.tcol[n], make_desc(r0_s + s * 2 * LBO_A, LBO_A, a.desc_mode), make_desc(r1_s + s * 2 * L.lboB[n], L.lboB[n], a.desc_mode), idesc,
                 s > 0 ? 1u : 0u);
      mma_commit(bar_layer);
    }
  }

  // ---- logits -> NormalTanhDistribution sample / log-prob (ppo_networks.py:45-83) ----
  mbar_wait(bar_layer, layer_phase);
  tc_fence_after();
  const int nu = a.d.nu, nlog = 2 * nu, sstride = L.N[5] + 1;
  float* S = reinterpret_cast<float*>(R0);        // [128][N5 + 1] logits
  float* LP = reinterpret_cast<float*>(R1);       // [128][nu] per-dimension log-prob terms
  float* LPR = LP + TILE_M * nu;                  // same for the uniform draw
  if (warp < 4) {
    const uint32_t tb = tmem + ((uint32_t)(warp * 32) << 16) + L.tcol[5];
    const float* bias = P + L.offP[5];
    for (int c0 = 0; c0 < L.N[5]; c0 += 32) {
      float v[32];
      tmem_ld32(tb + c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (a.dump && a.dump_layer == 5 && blockIdx.x == 0) a.dump[(size_t)tid * L.N[5] + c0 + j] = v[j];
        S[tid * sstride + c0 + j] = v[j] + bias[c0 + j];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  for (int item = tid; item < TILE_M * nu; item += THREADS) {
    const int r = item / nu, i = item - r * nu, grow = row0 + r;
    if (grow >= a.B) continue;
    const float loc = S[r * sstride + i];
    const float scale = softplus(S[r * sstride + nu + i]) + 1e-3f;  // brax NormalTanhDistribution min_std
    const float e = a.eps_a ? __ldg(a.eps_a + (size_t)grow * nu + i) : 0.0f;
    const float raw = loc + scale * e;
    const float log_scale = logf(scale);
    if (a.action) a.action[(size_t)grow * nu + i] = tanhf(raw);
    if (a.raw_action) a.raw_action[(size_t)grow * nu + i] = raw;
    LP[item] = -0.5f * e * e - log_scale - 0.91893853321f - tanh_log_det(raw);
    if (a.rand_action) {
      const float ra = __ldg(a.rand_action + (size_t)grow * nu + i);
      const float zr = (ra - loc) / scale;
      LPR[item] = -0.5f * zr * zr - log_scale - 0.91893853321f - tanh_log_det(ra);
    }
  }
  if (a.logits)
    for (int item = tid; item < TILE_M * nlog; item += THREADS) {
      const int r = item / nlog, i = item - r * nlog, grow = row0 + r;
      if (grow < a.B) a.logits[(size_t)grow * nlog + i] = S[r * sstride + i];
    }
  __syncthreads();
  if (tid < TILE_M && row0 + tid < a.B) {
    float s = 0.0f, sr = 0.0f;
    for (int i = 0; i < nu; ++i) s += LP[tid * nu + i];
    if (a.log_prob) a.log_prob[row0 + tid] = s;
    if (a.rand_action && a.rand_log_prob) {
      for (int i = 0; i < nu; ++i) sr += LPR[tid * nu + i];
      a.rand_log_prob[row0 + tid] = sr;
    }
  }
  __syncthreads();
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

int launch(Args a, void* stream) {
  static bool attr_set[64] = {};  // per device; idempotent, a race sets it twice
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -9;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(vnl_policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return -(int)e;
    attr_set[dev] = true;
  }
  static const int desc_mode = getenv("VNL_POLICY_DESC_MODE") ? atoi(getenv("VNL_POLICY_DESC_MODE")) : 0;
  a.desc_mode = desc_mode;
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
  if (!blob_dev || !traj || !eps_z || B < 0 || (dims->obs > 0 && !obs) || ((obs_mean == nullptr) != (obs_std == nullptr))) return -1;
  if (B == 0) return 0;
  Args a{*dims, B, static_cast<const uint8_t*>(blob_dev), traj, obs, obs_mean, obs_std, eps_z, eps_a, rand_action,
         action, raw_action, logits, log_prob, rand_log_prob, z_mean, z_logvar, -1, nullptr, 0};
  return launch(a, stream);
}

int vnl_policy_debug(const void* blob_dev, const VnlPolicyDims* dims, int B, const float* traj, const float* obs,
                     const float* obs_mean, const float* obs_std, const float* eps_z, int layer, float* dump, void* stream) {
  const int rc = check_dims(dims);
  if (rc) return rc;
  if (!blob_dev || !traj || !eps_z || !dump || B <= 0 || layer < 0 || layer > 5) return -1;
  Args a{*dims, B, static_cast<const uint8_t*>(blob_dev), traj, obs, obs_mean, obs_std, eps_z, nullptr, nullptr,
         nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, layer, dump, 0};
  return launch(a, stream);
}

}  // extern "C"

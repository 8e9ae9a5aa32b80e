// vnl_gemm.cu -- the dense contractions of the PPO update on the 5th-generation tensor cores.
//
//     C[M, N] (+)= sum over pairs p of  A_p[M, K] . B_p[N, K]^T   (+ bias[N])
//
// tcgen05.mma.cta_group::1.kind::tf32 (fp32 operands read as tf32, fp32 accumulation in tensor memory), operands staged
// in shared memory by TMA tensor maps (cp.async.bulk.tensor.2d, 128-byte swizzle) through a 4-stage mbarrier ring, warp
// specialised: warp 0 = TMA producer (one elected lane), warp 1 = tensor-memory allocation + MMA issue (one lane),
// warps 2-5 = epilogue (tcgen05.ld 32x32b: thread = accumulator row; bias; fp32 stores or red.add for split-K).
//
// Tiles: 128 x {64, 128, 256}, and 256 x 256 (two 128-row accumulators side by side in tensor memory, all 512 columns; 3 stages of
// 64 KB; 8 epilogue warps) for the wide value-network products.  These GEMMs are bound by the L2 -> shared-memory operand
// stream, not by the tensor pipe: fp32 operands cost 4 bytes per element, the chip's L2 delivers ~6300 B / cycle (B300_MICROARCH)
// = 43 B / cycle / SM with every SM pulling, and a 128 x 256 x 32 block needs 48 KB for 2.1 MFLOP -- 1126 cycles of transfer against
// 525 cycles of MMA.  The 256 x 256 tile moves 64 KB for twice the work (1.5 x the intensity) and turns 168 tiles on 148 SMs (two
// waves, the second 14 % full) into 84 in one wave.
//
// Either operand may be K-major (reduction index contiguous in memory) or MN-major (row index contiguous): the three
// products of a dense layer y = x W (flax kernel W[in, out], activations [rows, features]) need no transposed copies:
//     forward  y  = x . W      A = x  [rows, in]  K-major,   B = W  [in, out]   MN-major (N = out contiguous)
//     dgrad    dx = dy . W^T   A = dy [rows, out] K-major,   B = W  [in, out]   K-major  (N = in, K = out contiguous)
//     wgrad    dW = x^T . dy   A = x  [rows, in]  MN-major (M = in),  B = dy [rows, out] MN-major (N = out), K = rows
// Up to three operand pairs accumulate into the same tile: the 3xTF32 split (hi.hi + hi.lo + lo.hi with x = hi + lo, hi =
// x truncated to tf32) gives fp32-class accuracy on the tensor cores for the parity tests; one pair is the production
// mode and is also what XLA runs for the reference on an NVIDIA GPU (f32 dots at default precision use TF32).
//
// Reference replaced: the matmuls inside `jax.value_and_grad(compute_ppo_intention_loss)` (ppo_imitation/train.py:251-268
// -> intention_losses.py:91-202 -> IntentionNetwork / brax MLP `linen.Dense`).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/vnl_train.h"

namespace {

constexpr int BM = 128;      // accumulator lanes of one MMA; a tile is MT x 128 rows
constexpr int BK = 32;       // tf32 elements per K block = one 128-byte swizzle row
__host__ __device__ constexpr int threads_for(int mt) { return 64 + 128 * mt; }  // producer warp, issuer warp, 4 epilogue warps per 128 rows
__host__ __device__ constexpr int stages_for(int bn, int mt) { return (bn == 256 && mt == 2) ? 3 : 4; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
               "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
               : "memory");
}
// shared-memory matrix descriptor, descriptor version 1 (Blackwell), 128-byte swizzle.
//   K-major : layout type 2 (SWIZZLE_128B: 16-byte chunks permuted by row % 8; TMA CU_TENSOR_MAP_SWIZZLE_128B).  Rows of 128
//             bytes = 32 tf32 along K, 8-row groups SBO = 1024 bytes apart, LBO unused (1).
//   MN-major: 32-bit operands only exist in layout type 1 (SWIZZLE_128B_BASE32B: 32-byte chunks permuted by row % 4; TMA
//             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  Rows of 128 bytes = 32 elements along MN, one row per k; an atom is 4 rows
//             (512 bytes); SBO = next 4 along K, LBO = next 32 along MN.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout_type << 61);
}
// instruction descriptor: D fp32 (bit 4), A / B tf32 (format 2 at bits 7 / 10), majors at bits 15 / 16, N >> 3 at 17, M >> 4 at 24
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}
// linen.swish and its derivative, as the row kernels of vnl_train.cu compute them
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float dswish(float x) { const float s = sigmoidf(x); return s + x * s * (1.0f - s); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

struct GemmArgs {
  int M, N, K, npairs, a_mn, b_mn, ldc, kb_per_split, atomic;
  float* C;
  const float* bias;
  int epilogue, ldaux;  // 0 plain; 1: aux = swish(C) written beside C; 2: C = acc * swish'(aux)  (value MLP activations)
  float* aux;
};

template <int BN, int MT>
__global__ void __launch_bounds__(threads_for(MT), 1)
vnl_gemm_tf32_kernel(const __grid_constant__ CUtensorMap ta0, const __grid_constant__ CUtensorMap ta1, const __grid_constant__ CUtensorMap ta2,
                     const __grid_constant__ CUtensorMap tb0, const __grid_constant__ CUtensorMap tb1, const __grid_constant__ CUtensorMap tb2,
                     const GemmArgs g) {
  constexpr int STAGES = stages_for(BN, MT), TM = BM * MT;
  constexpr uint32_t A_BYTES = TM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN * MT < 32 ? 32 : BN * MT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);  // swizzle atoms: 1024-byte aligned
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]), accum_bar = smem_u32(&bars[2 * STAGES]);
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TM;
  const int kblocks = (g.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * g.kb_per_split;
  int kb1 = kb0 + g.kb_per_split;
  if (kb1 > kblocks) kb1 = kblocks;
  const int nkb = kb1 > kb0 ? kb1 - kb0 : 0;
  const int iters = nkb * g.npairs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // tensor memory: MT accumulators of BN fp32 columns x 128 lanes (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES, round = it / STAGES;
        if (round > 0) mbar_wait(empty0 + 8 * s, (round - 1) & 1);
        const int p = it / nkb, kb = kb0 + it % nkb;
        const CUtensorMap* ta = p == 0 ? &ta0 : (p == 1 ? &ta1 : &ta2);
        const CUtensorMap* tb = p == 0 ? &tb0 : (p == 1 ? &tb1 : &tb2);
        const uint32_t a_s = smem_u32(smem + (size_t)s * STAGE_BYTES), b_s = a_s + A_BYTES, bar = full0 + 8 * s;
        mbar_expect_tx(bar, STAGE_BYTES);
        if (!g.a_mn) tma_load_2d(a_s, ta, bar, kb * BK, m0);  // [TM rows][32 k]
        else
          for (int j = 0; j < TM / 32; ++j) tma_load_2d(a_s + j * 4096, ta, bar, m0 + 32 * j, kb * BK);  // TM / 32 x [32 k][32 m]
        if (!g.b_mn) tma_load_2d(b_s, tb, bar, kb * BK, n0);
        else
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(b_s + j * 4096, tb, bar, n0 + 32 * j, kb * BK);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = make_idesc(BN, g.a_mn, g.b_mn);
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES, round = it / STAGES;
        mbar_wait(full0 + 8 * s, round & 1);
        tc_fence_after();
        const uint32_t a_s = smem_u32(smem + (size_t)s * STAGE_BYTES), b_s = a_s + A_BYTES;
        // one instruction = 8 tf32 along K: K-major advances 32 bytes inside the swizzle row, MN-major 8 rows = 1024 bytes
        const uint64_t da0 = g.a_mn ? make_desc(a_s, 4096, 512, 1) : make_desc(a_s, 16, 1024, 2);
        const uint64_t db0 = g.b_mn ? make_desc(b_s, 4096, 512, 1) : make_desc(b_s, 16, 1024, 2);
        const uint64_t sa = g.a_mn ? (1024 >> 4) : (32 >> 4), sb = g.b_mn ? (1024 >> 4) : (32 >> 4);
        // the second 128 rows of a 256-row tile sit 16 KB further in either layout and accumulate BN columns further in tensor memory
#pragma unroll
        for (int h = 0; h < MT; ++h)
#pragma unroll
          for (int j = 0; j < BK / 8; ++j)
            mma_tf32(tmem + (uint32_t)(h * BN), da0 + (uint64_t)(h * (16384 >> 4)) + j * sa, db0 + j * sb, idesc, (it > 0 || j > 0) ? 1u : 0u);
        mma_commit(empty0 + 8 * s);  // frees the slot when these MMAs have read it
      }
      mma_commit(accum_bar);  // all MMAs of the tile done: accumulators readable
    }
  } else {  // ===== epilogue: a warp may read the tensor-memory lane quadrant warp % 4; warps 2-5 take the first 128 rows, 6-9 the second =====
    const int q = warp & 3, h = (warp - 2) >> 2, row = m0 + BM * h + 32 * q + lane;
    if (iters > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    const bool add_bias = g.bias != nullptr && blockIdx.z == 0;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= g.N) break;
      uint32_t v[32];
      if (iters > 0) tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(h * BN + c0), v);
      else
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      if (row < g.M) {
        float* dst = g.C + (size_t)row * g.ldc + n0 + c0;
        const int ncol = g.N - (n0 + c0) < 32 ? g.N - (n0 + c0) : 32;
        if (g.atomic && ncol == 32 && (g.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          // split-K partials: 16-byte vector reductions (a quarter of the L2 atomic operations of scalar red.add)
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (add_bias) {
              const float4 b = *reinterpret_cast<const float4*>(g.bias + n0 + c0 + j);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
          }
        } else if (g.atomic) {
#pragma unroll  // (static indices: the accumulator chunk stays in registers)
          for (int j = 0; j < 32; ++j)
            if (j < ncol) atomicAdd(dst + j, __uint_as_float(v[j]) + (add_bias ? g.bias[n0 + c0 + j] : 0.0f));
        } else if (ncol == 32 && (g.ldc & 3) == 0) {
          float* aux = g.epilogue ? g.aux + (size_t)row * g.ldaux + n0 + c0 : nullptr;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (add_bias) {
              const float4 b = *reinterpret_cast<const float4*>(g.bias + n0 + c0 + j);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (g.epilogue == 1) {  // the layer's activation, written beside its pre-activation (kept for the backward pass)
              *reinterpret_cast<float4*>(aux + j) = make_float4(o.x * sigmoidf(o.x), o.y * sigmoidf(o.y), o.z * sigmoidf(o.z), o.w * sigmoidf(o.w));
            } else if (g.epilogue == 2) {  // dgrad straight through the activation in front: dpre = dh * swish'(pre)
              const float4 x = *reinterpret_cast<const float4*>(aux + j);
              o.x *= dswish(x.x); o.y *= dswish(x.y); o.z *= dswish(x.z); o.w *= dswish(x.w);
            }
            *reinterpret_cast<float4*>(dst + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncol) dst[j] = __uint_as_float(v[j]) + (add_bias ? g.bias[n0 + c0 + j] : 0.0f);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// ---- host: tensor maps -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  // resolved through the runtime each time it is needed (no libcuda link dependency, no cached global)
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<EncodeFn>(fn);
}

// operand [rows_mn, K]: K-major = memory [mn][k] (ld elements between mn rows); MN-major = memory [k][mn]
int make_map(EncodeFn enc, CUtensorMap* map, const float* ptr, int mn, int k, int ld, int mn_major, int tile_mn) {
  if (!ptr || (reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 3) || ld < (mn_major ? mn : k)) return -2;
  cuuint64_t dims[2], strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2], estr[2] = {1, 1};
  if (!mn_major) { dims[0] = (cuuint64_t)k; dims[1] = (cuuint64_t)mn; box[0] = BK; box[1] = (cuuint32_t)tile_mn; }
  else { dims[0] = (cuuint64_t)mn; dims[1] = (cuuint64_t)k; box[0] = 32; box[1] = BK; }
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

template <int BN, int MT>
int launch(const CUtensorMap* ta, const CUtensorMap* tb, const GemmArgs& g, int splits, cudaStream_t stream) {
  const size_t smem = (size_t)stages_for(BN, MT) * (BM * MT * 128 + BN * 128) + 1024;
  cudaError_t err = cudaFuncSetAttribute(vnl_gemm_tf32_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM * MT - 1) / (BM * MT), splits);
  vnl_gemm_tf32_kernel<BN, MT><<<grid, threads_for(MT), smem, stream>>>(ta[0], ta[1], ta[2], tb[0], tb[1], tb[2], g);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

int vnl_gemm_tf32(int M, int N, int K, int npairs, const float* const* A, int lda, int a_mn_major, const float* const* B, int ldb,
                  int b_mn_major, float* C, int ldc, const float* bias, int splitk, void* stream) {
  return vnl_gemm_tf32_ex(M, N, K, npairs, A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, bias, splitk, 0, nullptr, 0, stream);
}

int vnl_gemm_tf32_ex(int M, int N, int K, int npairs, const float* const* A, int lda, int a_mn_major, const float* const* B, int ldb,
                     int b_mn_major, float* C, int ldc, const float* bias, int splitk, int epilogue, float* aux, int ldaux, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || npairs < 1 || npairs > 3 || !A || !B || !C || ldc < N) return -1;
  // fused activations need whole 16-byte row segments and a single pass over K
  if (epilogue < 0 || epilogue > 2 || (epilogue && (!aux || ldaux < N || (ldaux & 3) || (ldc & 3) || (N & 31) || splitk > 1 ||
                                                    (reinterpret_cast<uintptr_t>(aux) & 15) || (reinterpret_cast<uintptr_t>(C) & 15))))
    return -5;
  EncodeFn enc = encode_fn();
  if (!enc) return -4;
  const int kblocks = (K + BK - 1) / BK;
  int splits = splitk < 1 ? 1 : splitk;
  if (splits > kblocks) splits = kblocks;
  // tile: 128 x 256 where that still fills the machine (>= 148 tiles), else 128 x 128; 128 x 64 for narrow outputs; 256 x 256 once
  // the 128 x 256 tiling would need a second wave (or, with split-K, when the big tiles and their splits still occupy the machine)
  const long tiles256 = (long)((N + 255) / 256) * ((M + BM - 1) / BM), tiles_big = (long)((N + 255) / 256) * ((M + 2 * BM - 1) / (2 * BM));
  const bool big = N % 256 == 0 && M >= 2 * BM && (splits <= 1 ? tiles256 > 148 : tiles_big * splits >= 96);
  const int bn = N <= 64 ? 64 : ((big || (N % 256 == 0 && tiles256 >= 148 && splits <= 1)) ? 256 : 128);
  CUtensorMap ta[3], tb[3];
  for (int p = 0; p < 3; ++p) {
    const int q = p < npairs ? p : 0;
    int rc = make_map(enc, &ta[p], A[q], M, K, lda, a_mn_major, big ? 2 * BM : BM);
    if (rc) return rc;
    rc = make_map(enc, &tb[p], B[q], N, K, ldb, b_mn_major, bn);
    if (rc) return rc - 10;
  }
  GemmArgs g;
  g.M = M; g.N = N; g.K = K; g.npairs = npairs; g.a_mn = a_mn_major ? 1 : 0; g.b_mn = b_mn_major ? 1 : 0; g.ldc = ldc;
  g.kb_per_split = (kblocks + splits - 1) / splits;
  splits = (kblocks + g.kb_per_split - 1) / g.kb_per_split;
  g.atomic = splits > 1 ? 1 : 0;  // split-K partial tiles are added with red.global.add: the caller zeroes C first
  g.C = C; g.bias = bias;
  g.epilogue = epilogue; g.aux = aux; g.ldaux = ldaux;
  if (big) return launch<256, 2>(ta, tb, g, splits, (cudaStream_t)stream);
  if (bn == 256) return launch<256, 1>(ta, tb, g, splits, (cudaStream_t)stream);
  return bn == 128 ? launch<128, 1>(ta, tb, g, splits, (cudaStream_t)stream) : launch<64, 1>(ta, tb, g, splits, (cudaStream_t)stream);
}

}  // extern "C"

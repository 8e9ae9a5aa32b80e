// Observation-normaliser update (brax running_statistics.update at ppo_imitation/train.py:330-334), see
// include/vnl_normalizer.h.  One HBM pass: the batch is read once, fully coalesced; everything else is width-sized.
//
// Partial kernel: a CTA is `rpi` row lanes x `width` (or width / 4 with 16-byte loads) feature lanes (986 threads = 17 rows x
// 58 float4 for the rodent's 232 features), so consecutive threads read consecutive words of `rpi` consecutive rows and
// every thread keeps the same features throughout: its
// accumulators (sum d, sum d^2) live in registers for the whole kernel.  16 row groups (59 KB, contiguous) are in flight
// per CTA.  The grid is one CTA per SM, each owning one contiguous block of rows; per-CTA partials go to the workspace
// and the last CTA to arrive (ticket counter) adds them in index order: deterministic, one launch.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/vnl_normalizer.h"
#include "vnl_xla_status.h"

namespace {

constexpr int MAX_WIDTH = 1024;
constexpr int GRID = 148;   // one CTA per SM: few partials for the tail, the loads in flight come from the unroll
constexpr int UNROLL = 16;

template <int V>  // V = 4: a thread owns four adjacent features and moves 16 bytes per load (width % 4 == 0, aligned rows)
__global__ void __launch_bounds__(1024, 1) obs_stats_partial_kernel(const float* __restrict__ batch, long long rows, int width,
                                                                   int rpi, const float* __restrict__ mean,
                                                                   float* __restrict__ part, unsigned* ticket,
                                                                   float* __restrict__ sums) {
  extern __shared__ float sm[];  // [2][rpi][width]
  __shared__ bool last;
  constexpr int UN = V == 4 ? 8 : UNROLL;
  const int wv = width / V;
  const int t = threadIdx.x, c = (t % wv) * V, q = t / wv;
  float s1[V], s2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) s1[j] = s2[j] = 0.0f;
  if (q < rpi) {
    float m[V];
#pragma unroll
    for (int j = 0; j < V; ++j) m[j] = __ldg(mean + c + j);
    // each CTA owns one contiguous block of rows (long DRAM bursts); UN row groups of rpi rows are in flight per CTA
    const long long per = ((rows + gridDim.x - 1) / gridDim.x + rpi - 1) / rpi * rpi;
    const long long end = min(rows, ((long long)blockIdx.x + 1) * per);
    const long long stride = rpi;
    long long r = (long long)blockIdx.x * per + q;
    auto acc = [&](const float* x) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float d = x[j] - m[j];
        s1[j] += d;
        s2[j] = fmaf(d, d, s2[j]);
      }
    };
    auto load = [&](long long row, float* x) {  // streamed once
      if constexpr (V == 4) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(batch + row * width + c));
        x[0] = v.x, x[1] = v.y, x[2] = v.z, x[3] = v.w;
      } else {
        x[0] = __ldcs(batch + row * width + c);
      }
    };
    for (; r + (UN - 1) * stride < end; r += UN * stride) {
      float x[UN][V];
#pragma unroll
      for (int u = 0; u < UN; ++u) load(r + u * stride, x[u]);
#pragma unroll
      for (int u = 0; u < UN; ++u) acc(x[u]);
    }
    for (; r < end; r += stride) {
      float x[V];
      load(r, x);
      acc(x);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      sm[q * width + c + j] = s1[j];
      sm[(rpi + q) * width + c + j] = s2[j];
    }
  }
  __syncthreads();
  if (t < width) {
    float a = 0.0f, b = 0.0f;
    for (int k = 0; k < rpi; ++k) a += sm[k * width + t], b += sm[(rpi + k) * width + t];
    __stcg(part + (size_t)blockIdx.x * 2 * width + t, a);
    __stcg(part + (size_t)blockIdx.x * 2 * width + width + t, b);
  }
  // the last CTA to arrive adds the per-CTA partials in index order (which CTA that is does not change the result)
  __threadfence();
  __syncthreads();
  if (t == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  // two threads per output (even / odd partials), eight loads in flight each; fixed combination order
  const int nparts = gridDim.x;
  for (int base = 0; base < 2 * width; base += blockDim.x / 2) {
    const int i = base + (t >> 1), h = t & 1;
    float tot = 0.0f;
    if (i < 2 * width) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      int k = h;
      for (; k + 14 < nparts; k += 16) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += __ldcg(part + (size_t)(k + 2 * u) * 2 * width + i);
      }
      for (; k < nparts; k += 2) acc[0] += __ldcg(part + (size_t)k * 2 * width + i);
      tot = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    }
    const float other = __shfl_xor_sync(0xffffffffu, tot, 1);
    if (i < 2 * width && h == 0) sums[i] = tot + other;
  }
  if (t == 0) {
    sums[2 * width] = (float)rows;
    *ticket = 0;  // ready for the next call
  }
}

__global__ void obs_stats_finish_kernel(const float* __restrict__ sums, int width, float* count, float* mean, float* sv, float* sd,
                                        float std_min, float std_max) {
  const float cnt = *count + sums[2 * width];
  __syncthreads();  // every thread has read the old count before thread 0 replaces it
  if (!(cnt > 0.0f)) return;  // nothing seen yet (all ranks sent empty batches): keep init_state, no 0 / 0
  for (int c = threadIdx.x; c < width; c += blockDim.x) {
    const float s1 = sums[c], s2 = sums[width + c];
    const float mu = s1 / cnt;  // mean' - mean
    mean[c] += mu;
    const float v = sv[c] + (s2 - mu * s1);
    sv[c] = v;
    sd[c] = fminf(fmaxf(sqrtf(fmaxf(v, 0.0f) / cnt), std_min), std_max);
  }
  if (threadIdx.x == 0) *count = cnt;
}

int grid_for(long long rows, int rpi) {
  const long long groups = (rows + rpi - 1) / rpi;
  return (int)(groups < GRID ? (groups < 1 ? 1 : groups) : GRID);
}

}  // namespace

extern "C" {

size_t vnl_obs_stats_workspace_bytes(int width) {
  if (width < 1 || width > MAX_WIDTH) return 0;
  return (size_t)GRID * 2 * width * sizeof(float) + 16;  // per-CTA partials + the arrival ticket
}

int vnl_obs_stats_partial(const float* batch, long long rows, int width, const float* mean, void* workspace, float* sums,
                          void* stream) {
  if (width < 1 || width > MAX_WIDTH || rows < 0 || !mean || !workspace || !sums || (rows > 0 && !batch)) return -1;
  const bool vec = width % 4 == 0 && (reinterpret_cast<uintptr_t>(batch) & 15) == 0;
  const int wv = vec ? width / 4 : width;
  int rpi = 1024 / wv;
  if (rpi > 32) rpi = 32;  // bounds the shared-memory block reduction (2 * rpi * width floats)
  const int threads = (rpi * wv + 31) / 32 * 32;
  const int grid = grid_for(rows, rpi);
  float* part = static_cast<float*>(workspace);
  unsigned* ticket = reinterpret_cast<unsigned*>(part + (size_t)GRID * 2 * width);
  const size_t smem = 2 * (size_t)rpi * width * sizeof(float);
  if (vec)
    obs_stats_partial_kernel<4><<<grid, threads, smem, (cudaStream_t)stream>>>(batch, rows, width, rpi, mean, part, ticket, sums);
  else
    obs_stats_partial_kernel<1><<<grid, threads, smem, (cudaStream_t)stream>>>(batch, rows, width, rpi, mean, part, ticket, sums);
  return -(int)cudaGetLastError();
}

int vnl_obs_stats_finish(const float* sums, int width, float* count, float* mean, float* summed_variance, float* std,
                         float std_min_value, float std_max_value, void* stream) {
  if (width < 1 || width > MAX_WIDTH || !sums || !count || !mean || !summed_variance || !std) return -1;
  obs_stats_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sums, width, count, mean, summed_variance, std, std_min_value,
                                                              std_max_value);
  return -(int)cudaGetLastError();
}

// Legacy XLA custom calls.  XLA hands out uninitialised result buffers and never aliases operands with results, so the
// trampolines zero the arrival ticket of the scratch result and copy the state operands into the state results first.
// partial: opaque = int64 rows, int32 width; buffers = [batch, mean, (outputs) sums, workspace]
void vnl_xla_obs_stats_partial(void* stream, void** b, const char* opaque, size_t opaque_len, void* status) {
  if (!b || !opaque || opaque_len < 12) { vnl::xla_report(status, "vnl_xla_obs_stats_partial", -30); return; }
  long long rows;
  int32_t width;
  memcpy(&rows, opaque, 8);
  memcpy(&width, opaque + 8, 4);
  if (width < 1 || width > MAX_WIDTH) { vnl::xla_report(status, "vnl_xla_obs_stats_partial", -1); return; }
  cudaMemsetAsync(static_cast<float*>(b[3]) + (size_t)GRID * 2 * width, 0, 16, (cudaStream_t)stream);
  vnl::xla_report(status, "vnl_xla_obs_stats_partial",
                  vnl_obs_stats_partial((const float*)b[0], rows, width, (const float*)b[1], b[3], (float*)b[2], stream));
}
// finish: opaque = int32 width, float std_min, float std_max; buffers = [sums, count, mean, summed_variance, std,
//         (outputs) count', mean', summed_variance', std']
void vnl_xla_obs_stats_finish(void* stream, void** b, const char* opaque, size_t opaque_len, void* status) {
  if (!b || !opaque || opaque_len < 12) { vnl::xla_report(status, "vnl_xla_obs_stats_finish", -30); return; }
  int32_t width;
  float lo, hi;
  memcpy(&width, opaque, 4);
  memcpy(&lo, opaque + 4, 4);
  memcpy(&hi, opaque + 8, 4);
  if (width < 1 || width > MAX_WIDTH) { vnl::xla_report(status, "vnl_xla_obs_stats_finish", -1); return; }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemcpyAsync(b[5], b[1], sizeof(float), cudaMemcpyDeviceToDevice, st);
  for (int i = 0; i < 3; ++i) cudaMemcpyAsync(b[6 + i], b[2 + i], (size_t)width * sizeof(float), cudaMemcpyDeviceToDevice, st);
  vnl::xla_report(status, "vnl_xla_obs_stats_finish",
                  vnl_obs_stats_finish((const float*)b[0], width, (float*)b[5], (float*)b[6], (float*)b[7], (float*)b[8], lo, hi, stream));
}

}  // extern "C"

// PPO data-side ops (include/vnl_ppo.h): generalised advantage estimation, the reverse scan of
// ppo_imitation/intention_losses.py:26-89, one thread per env.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/vnl_ppo.h"
#include "vnl_xla_status.h"

namespace {

__global__ void __launch_bounds__(256) gae_kernel(int T, int B, const float* __restrict__ truncation,
                                                  const float* __restrict__ termination, const float* __restrict__ rewards,
                                                  const float* __restrict__ values, const float* __restrict__ bootstrap,
                                                  float lambda_, float discount, float* __restrict__ vs,
                                                  float* __restrict__ advantages) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float v_next = __ldg(bootstrap + b);  // values_{t+1}
  float vs_next = v_next;               // vs_{t+1}
  float acc = 0.0f;
  // the four operand rows of step t - 1 are requested while step t is reduced
  size_t i = (size_t)(T - 1) * B + b;
  float tr = __ldcs(truncation + i), te = __ldcs(termination + i), r = __ldcs(rewards + i), v = __ldcs(values + i);
  for (int t = T - 1; t >= 0; --t) {
    float ntr = 0.0f, nte = 0.0f, nr = 0.0f, nv = 0.0f;
    if (t > 0) {
      const size_t j = i - B;
      ntr = __ldcs(truncation + j), nte = __ldcs(termination + j), nr = __ldcs(rewards + j), nv = __ldcs(values + j);
    }
    const float mask = 1.0f - tr, cont = discount * (1.0f - te);
    const float delta = (r + cont * v_next - v) * mask;
    acc = delta + cont * mask * lambda_ * acc;
    const float vs_t = acc + v;
    if (advantages) advantages[i] = (r + cont * vs_next - v) * mask;
    if (vs) vs[i] = vs_t;
    v_next = v;
    vs_next = vs_t;
    tr = ntr, te = nte, r = nr, v = nv;
    i -= B;
  }
}

}  // namespace

extern "C" {

int vnl_gae(int T, int B, const float* truncation, const float* termination, const float* rewards, const float* values,
            const float* bootstrap_value, float lambda_, float discount, float* vs, float* advantages, void* stream) {
  if (T < 0 || B < 0) return -1;
  if (T == 0 || B == 0) return 0;
  if (!truncation || !termination || !rewards || !values || !bootstrap_value || (!vs && !advantages)) return -1;
  gae_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(T, B, truncation, termination, rewards, values, bootstrap_value,
                                                               lambda_, discount, vs, advantages);
  return -(int)cudaGetLastError();
}

void vnl_xla_gae(void* stream, void** b, const char* opaque, size_t opaque_len, void* status) {
  if (!b || !opaque || opaque_len < 16) { vnl::xla_report(status, "vnl_xla_gae", -30); return; }
  int32_t T, B;
  float lam, disc;
  memcpy(&T, opaque, 4);
  memcpy(&B, opaque + 4, 4);
  memcpy(&lam, opaque + 8, 4);
  memcpy(&disc, opaque + 12, 4);
  vnl::xla_report(status, "vnl_xla_gae", vnl_gae(T, B, (const float*)b[0], (const float*)b[1], (const float*)b[2], (const float*)b[3],
                                                 (const float*)b[4], lam, disc, (float*)b[5], (float*)b[6], stream));
}

}  // extern "C"

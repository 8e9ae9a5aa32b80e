/* C ABI of the PPO data-side ops of libvnl_b200.so (the part that carries no gradient).
 *
 * vnl_gae replaces `compute_gae` (ppo_imitation/intention_losses.py:26-89; called at :167-175 inside
 * compute_ppo_intention_loss).  Both results are wrapped in `jax.lax.stop_gradient` there (:89), so a custom call without a
 * differentiation rule is a drop-in.  Time-major [T, B] fp32 operands as the loss holds them after its swapaxes (:133):
 *     deltas_t   = (rewards_t + discount * (1 - termination_t) * values_{t+1} - values_t) * (1 - truncation_t)
 *     acc_t      = deltas_t + discount * (1 - termination_t) * (1 - truncation_t) * lambda * acc_{t+1}      (reverse scan)
 *     vs_t       = acc_t + values_t
 *     advantages_t = (rewards_t + discount * (1 - termination_t) * vs_{t+1} - values_t) * (1 - truncation_t)
 * with values_T = vs_T = bootstrap_value.  One thread per env walks its T steps backwards; consecutive threads read
 * consecutive envs (coalesced); every operand is read once and every result written once.
 *
 * Caller-owned device buffers, enqueue-only on `stream`, no allocation, no sync, CUDA-graph capturable.
 */
#ifndef VNL_PPO_H_
#define VNL_PPO_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int vnl_gae(int T, int B, const float* truncation, const float* termination, const float* rewards, const float* values,
            const float* bootstrap_value, float lambda_, float discount, float* vs, float* advantages, void* stream);

/* XLA custom call (status-returning legacy ABI, api_version 2): opaque = int32 T, int32 B, float lambda, float discount;
 * buffers = [truncation, termination, rewards, values, bootstrap_value, (outputs) vs, advantages]. */
void vnl_xla_gae(void* stream, void** buffers, const char* opaque, size_t opaque_len, void* status);

#ifdef __cplusplus
}
#endif
#endif /* VNL_PPO_H_ */

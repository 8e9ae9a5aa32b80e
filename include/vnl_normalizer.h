/* C ABI of the observation-normaliser update of libvnl_b200.so: the exchange step of the rollout path.
 *
 * Replaces `running_statistics.update(normalizer_params, data.observation, pmap_axis_name=...)` of brax
 * (brax/training/acme/running_statistics.py), called once per training step at ppo_imitation/train.py:330-334; its
 * result feeds the policy's `preprocess_observations_fn` (ppo_imitation/train.py:220-229 -> intention_policy_network.py:
 * 124-126 -> `obs_mean` / `obs_std` of vnl_policy_forward).  brax computes, per feature c over the batch rows x,
 *     count' = count + psum(rows)
 *     d = x - mean;  mean' = mean + psum(sum d) / count'
 *     summed_variance' = summed_variance + psum(sum d * (x - mean'))
 *     std' = clip(sqrt(max(summed_variance', 0) / count'), std_min, std_max)
 * Since x - mean' = d - (mean' - mean), the second sum is  sum d^2 - (mean' - mean) * sum d:  ONE pass over the batch
 * yields S1 = sum d and S2 = sum d^2 per feature, the only cross-GPU step is one all-reduce of [S1 | S2 | rows]
 * (2 * width + 1 floats, NCCL over NVLink, done by the caller between the two entry points), and the state update is a
 * width-sized epilogue.  Reduction orders are fixed (per-CTA partials combined in index order): bit-identical across runs.
 *
 * All pointers are caller-owned device memory; enqueue-only on `stream`; no allocation, no sync; CUDA-graph capturable.
 */
#ifndef VNL_NORMALIZER_H_
#define VNL_NORMALIZER_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Bytes of scratch vnl_obs_stats_partial needs for this feature width (per-CTA partial sums + an arrival ticket); 0 if
 * width is unsupported (1 <= width <= 1024).  The caller zero-fills the workspace ONCE after allocating it; calls leave
 * it ready for the next call.  One workspace serves one stream at a time. */
size_t vnl_obs_stats_workspace_bytes(int width);

/* sums[0:width] = sum over rows of (batch - mean), sums[width:2 width] = sum of squares of the same, sums[2 width] = rows.
 * batch [rows, width] fp32 row-major, mean [width] (the state's mean BEFORE the update). */
int vnl_obs_stats_partial(const float* batch, long long rows, int width, const float* mean, void* workspace,
                          float* sums, void* stream);

/* State update from the (all-reduced) sums, in place: count [1], mean / summed_variance / std [width]. */
int vnl_obs_stats_finish(const float* sums, int width, float* count, float* mean, float* summed_variance, float* std,
                         float std_min_value, float std_max_value, void* stream);

/* XLA custom-call entry points, status-returning legacy ABI (`void f(cudaStream_t, void** buffers, const char* opaque, size_t opaque_len, XlaCustomCallStatus* status)`, api_version 2; a non-zero code of the underlying call is reported through XlaCustomCallStatusSetFailure).
 * XLA result buffers are uninitialised and never alias operands: `partial` zeroes the scratch result's ticket itself,
 * `finish` copies the state operands into the state results before updating them in place.
 *   partial: opaque = int64 rows, int32 width;                 buffers = [batch, mean, (outputs) sums, workspace]
 *   finish:  opaque = int32 width, float std_min, float std_max; buffers = [sums, count, mean, summed_variance, std,
 *                                                                          (outputs) count', mean', summed_variance', std'] */
void vnl_xla_obs_stats_partial(void* stream, void** buffers, const char* opaque, size_t opaque_len, void* status);
void vnl_xla_obs_stats_finish(void* stream, void** buffers, const char* opaque, size_t opaque_len, void* status);

#ifdef __cplusplus
}
#endif
#endif /* VNL_NORMALIZER_H_ */

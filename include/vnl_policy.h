/* C ABI of the intention-network policy forward of libvnl_b200.so.
 *
 * Replaces, on the rollout path of the reference, what `actor_step` calls before `env.step`
 * (ppo_imitation/acting.py:47-48):
 *   make_inference_fn(...).policy             ppo_imitation/ppo_networks.py:45-83
 *     IntentionNetwork.__call__               ppo_imitation/intention_policy_network.py:82-105
 *       Encoder  (Dense -> relu -> LayerNorm per hidden layer; fc2_mean / fc2_logvar)   :20-45
 *       reparameterize                                                                   :76-79
 *       Decoder  (Dense -> relu -> LayerNorm, last layer plain) on concat([z, obs])      :48-73
 *     running_statistics.normalize on obs only (the `apply` closure, :124-126; traj is fed raw)
 *     brax NormalTanhDistribution: sample_no_postprocessing, log_prob, postprocess (tanh)
 *
 * One launch per call, one CTA per 128 envs: six dense layers as tcgen05.mma (bf16 operands, fp32 accumulation in
 * tensor memory), bias / relu / LayerNorm / reparameterisation / tanh-normal sampling in the epilogues (fp32), the
 * activations never leave the SM.  Random draws are operands (eps_z, eps_a), so the caller owns the RNG stream, as
 * the reference's `key` argument does.
 *
 * All pointers are caller-owned; `blob_dev` (16-byte aligned: it is streamed by TMA bulk copies), the inputs and outputs
 * are device memory, `dims` is host memory.
 * Enqueue-only on `stream`, no allocation, no synchronisation, CUDA-graph capturable.  Returns 0 or a negative error
 * code / cudaError.
 */
#ifndef VNL_POLICY_H_
#define VNL_POLICY_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct VnlPolicyDims {
  int32_t traj;   /* traj_size (rodent 795)                 encoder input                      */
  int32_t obs;    /* observation_size (rodent 232)          decoder side input                 */
  int32_t latent; /* intention_latent_size (64)             configs/train_config.yaml:15       */
  int32_t e1, e2; /* encoder_layer_sizes ([256, 128])       configs/train_config.yaml:16       */
  int32_t d1, d2; /* decoder_layer_sizes ([128, 256])       configs/train_config.yaml:17       */
  int32_t nu;     /* action_size; the decoder emits 2 * nu  ppo_networks.py:99-113             */
} VnlPolicyDims;

/* 0 if the kernel supports these sizes (hidden sizes multiples of 64 and <= 256, latent a multiple of 32 and
 * 2 * latent <= 256, nu <= 64, obs <= 256, tensor-memory and shared-memory budgets), else a negative code. */
int vnl_policy_check(const VnlPolicyDims* dims);

/* Size of the packed parameter blob for these sizes. */
size_t vnl_policy_blob_bytes(const VnlPolicyDims* dims);

/* Host-side packing of the flax parameter tree into the blob the kernel streams (bf16 weights in the tensor-core
 * operand image, fp32 bias / LayerNorm vectors).  `params` holds 22 host fp32 arrays in this order, kernels row-major
 * [in, out] as flax stores them:
 *   encoder: hidden_0.kernel [traj,e1], hidden_0.bias, LayerNorm_0.scale, LayerNorm_0.bias,
 *            hidden_1.kernel [e1,e2],   hidden_1.bias, LayerNorm_1.scale, LayerNorm_1.bias,
 *            fc2_mean.kernel [e2,latent], fc2_mean.bias, fc2_logvar.kernel [e2,latent], fc2_logvar.bias,
 *   decoder: hidden_0.kernel [latent+obs,d1], hidden_0.bias, LayerNorm_0.scale, LayerNorm_0.bias,
 *            hidden_1.kernel [d1,d2],         hidden_1.bias, LayerNorm_1.scale, LayerNorm_1.bias,
 *            hidden_2.kernel [d2,2*nu],       hidden_2.bias. */
int vnl_policy_pack(const VnlPolicyDims* dims, const float* const* params, void* blob_host, size_t nbytes);

/* The policy forward for B envs.
 *   traj [B,traj], obs [B,obs]            env outputs (info["traj"], State.obs), fp32
 *   obs_mean, obs_std [obs]               running-statistics normaliser; NULL = identity
 *   eps_z [B,latent], eps_a [B,nu]        standard-normal draws (reparameterize; sample_no_postprocessing)
 *   rand_action [B,nu] or NULL            the uniform draw of ppo_networks.py:66-71
 * outputs (any may be NULL):
 *   action [B,nu] = tanh(raw_action), raw_action [B,nu], logits [B,2*nu], log_prob [B], rand_log_prob [B],
 *   z_mean [B,latent], z_logvar [B,latent] (the encoder heads, needed by the KL term of the loss). */
int vnl_policy_forward(const void* blob_dev, const VnlPolicyDims* dims, int B, const float* traj, const float* obs,
                       const float* obs_mean, const float* obs_std, const float* eps_z, const float* eps_a,
                       const float* rand_action, float* action, float* raw_action, float* logits, float* log_prob,
                       float* rand_log_prob, float* z_mean, float* z_logvar, void* stream);

/* Test hook: the pre-activation of layer `layer` (0..5: the raw tcgen05 accumulators, no bias) for the first
 * min(B,128) envs, written row-major [128, N_layer] to `dump`.  layer -1: SM-clock stamps of CTA 0 at the phase
 * boundaries (32 floats; developer profile, tools/gpu_policy.py). */
int vnl_policy_debug(const void* blob_dev, const VnlPolicyDims* dims, int B, const float* traj, const float* obs,
                     const float* obs_mean, const float* obs_std, const float* eps_z, int layer, float* dump,
                     void* stream);

/* XLA custom-call entry point, status-returning legacy ABI (`void f(cudaStream_t, void** buffers, const char* opaque,
 * size_t opaque_len, XlaCustomCallStatus* status)`, api_version 2; failures reported through XlaCustomCallStatusSetFailure):
 * opaque = 9 little-endian int32 (the VnlPolicyDims fields in order, then B); buffers = [blob, traj, obs, obs_mean,
 * obs_std, eps_z, eps_a, rand_action, (outputs) action, raw_action, logits, log_prob, rand_log_prob]. */
void vnl_xla_policy_forward(void* stream, void** buffers, const char* opaque, size_t opaque_len, void* status);

#ifdef __cplusplus
}
#endif
#endif /* VNL_POLICY_H_ */

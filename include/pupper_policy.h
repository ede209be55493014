/*
 * pupper_policy.h — C ABI of the policy-MLP forward pass used by rollout collection (libpupper_env.so).
 *
 * What this replaces in the reference (rishihahs/pupperv3-mjx): the reference trains with Brax PPO, whose policy
 * is a dense MLP evaluated in JAX between two env steps; the layer list this library consumes is the one the
 * reference's exporter writes (pupperv3_mjx/export.py:31-62: per layer {"type": "dense", "activation", "weights":
 * [kernel [in, out], bias [out]]}, observation normalisation folded into layer 0 :7-10, mean half of the Gaussian
 * head :39-41).  BASELINE configs[4] ("rollout collection, policy MLP + env step") runs policy and env back to
 * back on the device; this entry point is the policy half.
 *
 * Same conventions as pupper_env.h: 0 or a negative PUPPER_E* code, nothing throws, nothing synchronises, work is
 * only enqueued on `stream`, device buffers are caller-owned, calls are CUDA-graph capturable.
 */
#ifndef PUPPER_POLICY_H_
#define PUPPER_POLICY_H_

#include <stdint.h>

#include "pupper_env.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PUPPER_POLICY_MAX_LAYERS 8
#define PUPPER_POLICY_MAX_WIDTH 1024 /* widest layer input / output */
#define PUPPER_POLICY_MAX_OUT 256    /* widest layer OUTPUT (accumulators live in registers) */

/* Activation codes (names as in the reference's utils.activation_fn_map / export.py). */
enum {
  PUPPER_ACT_LINEAR = 0,
  PUPPER_ACT_RELU = 1,
  PUPPER_ACT_SIGMOID = 2,
  PUPPER_ACT_ELU = 3,
  PUPPER_ACT_TANH = 4,
  PUPPER_ACT_SWISH = 5, /* = silu */
  PUPPER_ACT_GELU = 6,  /* erf form */
  PUPPER_ACT_LEAKY_RELU = 7
};

/* Arithmetic of the matrix products (accumulation is always float32):
 *   PUPPER_POLICY_TF32   operands rounded to TF32 (what XLA's default float32 matmul precision does on this GPU class);
 *                        layers up to 256 wide run on the tcgen05 / tensor-memory kernel (set PUPPER_POLICY_LEGACY=1 in the
 *                        environment at create time to force the mma.sync kernel);
 *   PUPPER_POLICY_3XTF32 each product as three TF32 products of hi/lo splits: float32-level accuracy (default). */
enum { PUPPER_POLICY_TF32 = 1, PUPPER_POLICY_3XTF32 = 3 };

typedef struct PupperPolicy PupperPolicy;

/* Builds the device-side weight tables.  weights[l] is a HOST pointer to a row-major [in_dims[l], out_dims[l]]
 * float32 matrix, biases[l] to [out_dims[l]] floats; in_dims[l + 1] must equal out_dims[l]. */
int pupper_policy_create(int n_layers, const int32_t *in_dims, const int32_t *out_dims, const int32_t *activations,
                         const float *const *weights, const float *const *biases, int device, int precision,
                         PupperPolicy **out);
int pupper_policy_destroy(PupperPolicy *policy);

/* action[n, out_dims[last]] = MLP(obs[n, in_dims[0]]); both device pointers, row-major, 16-byte aligned rows are NOT
 * required.  One kernel launch. */
int pupper_policy_forward(const PupperPolicy *policy, int n, const float *obs, float *action, pupper_stream_t stream);
/* Same, and the observation rows the kernel reads are also written to obs_record[n, in_dims[0]] (device pointer, may be NULL,
 * must not alias obs): the per-step rollout files obs[t] this way instead of with a separate copy. */
int pupper_policy_forward_record(const PupperPolicy *policy, int n, const float *obs, float *action, float *obs_record,
                                 pupper_stream_t stream);

/* Rollout collection, ONE launch per unroll (BASELINE configs[4]).  What it replaces in the reference's
 * training stack: Brax PPO's unroll (`brax.training.acting.generate_unroll` [3P]: a `lax.scan` of unroll_length x
 * (policy forward, env.step) over the wrapped env -- the reference hands it PupperV3Env through `brax.envs.training.wrap`,
 * pupperv3_mjx/environment.py:348 is the step it scans).  For t = 0 .. unroll_length - 1:
 *     traj_obs[t]    = state->obs                        [n_envs][observation_history * 36]
 *     traj_action[t] = policy(traj_obs[t])               [n_envs][12]   (deterministic head, as export.py:39-41 exports it)
 *     state, traj_reward[t], traj_done[t] = pupper_step(state, traj_action[t])   (episode / auto-reset block included)
 * All trajectory pointers are device memory, [unroll_length] slices back to back.  After the call `state` (and out->reward /
 * done / metrics) are what unroll_length successive pupper_step calls with those actions would have left -- bit for bit,
 * it is the same device code.  The policy's input width must be observation_history * 36, its output width 12, its layers
 * at most 256 wide; debug taps of `out` must be NULL; unroll_length <= 65535.
 * One kernel launch: a grid of (env groups of 32) x (steps); the CTA of (group, step) waits for the CTA of (group, step - 1)
 * through a per-group counter in device memory (allocated by the first call for a batch size: make that call outside a
 * stream capture; later calls only enqueue and are CUDA-graph capturable).  Groups never wait for each other.  Setting
 * PUPPER_ROLLOUT_PER_STEP in the environment before the first call switches to one launch per step (diagnostics). */
int pupper_rollout(const PupperModel *model, const PupperPolicy *policy, int n_envs, int unroll_length, const PupperDR *dr,
                   PupperState *state, PupperStepOut *out, PupperEpisode *episode, float *traj_obs, float *traj_action,
                   float *traj_reward, float *traj_done, pupper_stream_t stream);
/* Number of waits of the chained launch that timed out so far (always 0 unless the device dispatched CTAs out of order;
 * synchronises the device).  Negative: PUPPER_E* code. */
int pupper_rollout_timeouts(const PupperModel *model);

#ifdef __cplusplus
}
#endif
#endif

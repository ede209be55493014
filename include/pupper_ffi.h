/*
 * pupper_ffi.h -- the JAX side of the drop-in boundary: XLA FFI custom-call handlers over pupper_reset / pupper_step, and the
 * flat "blob" layouts that let JAX carry PupperState / PupperDR / PupperEpisode as plain float32 arrays.
 *
 * What this replaces in the reference (file:line into rishihahs/pupperv3-mjx): the bodies of PupperV3Env.reset / step
 * (pupperv3_mjx/environment.py:314, :348) under Brax's vmap, with the domain-randomised `sys_v` leaves of
 * domain_randomize(sys, rng) -> (sys_v, in_axes) (pupperv3_mjx/domain_randomization.py:93-112) packed into the DR blob and the
 * Brax EpisodeWrapper / AutoResetWrapper state packed into the episode blob.  INTEGRATION.md shows the Python side.
 *
 * jaxlib (which ships xla/ffi/api/c_api.h) is not installable in this image: the handlers are compiled against
 * csrc/xla_ffi_stub.h, a restatement of the few C structs they touch, and exercised through a hand-built call frame
 * (tests/test_ffi.py).  Build with -DPUPPER_XLA_FFI_HEADER='"xla/ffi/api/c_api.h"' to use the real header instead.
 */
#ifndef PUPPER_FFI_H_
#define PUPPER_FFI_H_

#include "pupper_env.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- blob layouts (host-side pointer arithmetic only; no CUDA call) -----------------------------------------------------
 * Every blob is an array of 4-byte words (float32 on the JAX side; integer fields keep their bit patterns).
 *   state blob   : the 14 SoA fields of PupperState in declaration order, field f as [rows_f][stride] (rows from
 *                  pupper_state_rows), then obs [n_envs][H*36] env-major
 *   DR blob      : [58][stride]: friction, kp, kd, base_ipos (3), body_inertia (39), body_mass (13)
 *   episode blob : [79][stride]: first_qpos (19), first_qvel (18), first_warmstart (18), steps, truncation, sum_reward,
 *                  length, sum_metrics (19), episode_done; then first_obs [n_envs][H*36]; then totals [24]
 *   rand blob    : [PUPPER_NRAND][stride] (PupperRand.u)
 * stride = n_envs rounded up to a multiple of 32. */
int pupper_blob_stride(int n_envs);
int64_t pupper_state_blob_words(const PupperEnvCfg *cfg, int n_envs);
int64_t pupper_dr_blob_words(int n_envs);
int64_t pupper_episode_blob_words(const PupperEnvCfg *cfg, int n_envs);
int64_t pupper_rand_blob_words(int n_envs);
/* "unpack": fill the pointer structs with addresses inside a blob (device or host memory alike) */
int pupper_state_blob_bind(const PupperEnvCfg *cfg, int n_envs, void *blob, PupperState *out);
int pupper_dr_blob_bind(int n_envs, const void *blob, PupperDR *out);
int pupper_episode_blob_bind(const PupperEnvCfg *cfg, int n_envs, void *blob, PupperEpisode *out);
/* "pack": gather the 6 batched DR leaves as domain_randomize returns them (host, env-major) into a host DR blob */
int pupper_dr_blob_pack(int n_envs, const float *friction /*[n]*/, const float *kp /*[n]*/, const float *kd /*[n]*/,
                        const float *base_ipos /*[n][3]*/, const float *body_inertia /*[n][13][3]*/,
                        const float *body_mass /*[n][13]*/, float *blob);

/* ---- per-device model registry: XLA calls the handlers from one host thread per device (8 under pmap); each looks up the
 * handle registered for the device its stream belongs to.  No process-global "current model". */
int pupper_ffi_register_model(int device, const PupperModel *model, const PupperEnvCfg *cfg);
int pupper_ffi_unregister_model(int device);
/* the policy of the rollout handlers, same convention (a handle of pupper_policy_create, include/pupper_policy.h; in_dim /
 * out_dim are its first layer's input and last layer's output width) */
struct PupperPolicy;
int pupper_ffi_register_policy(int device, const struct PupperPolicy *policy, int in_dim, int out_dim);
int pupper_ffi_unregister_policy(int device);

/* ---- XLA FFI handlers (typed-FFI C ABI: XLA_FFI_Error* handler(XLA_FFI_CallFrame*)) ---------------------------------------
 * PupperStepFfi   args: action f32[n,12], state blob, DR blob (or 0 elements), episode blob (or 0), rand blob (or 0)
 *                 rets: state blob (alias of arg 1), reward f32[n], done f32[n], metrics f32[n,19], episode blob (alias of arg 3)
 * PupperResetFfi  args: keys u32[n,2], DR blob (or 0), rand blob (or 0)
 *                 rets: state blob, reward f32[n], done f32[n], metrics f32[n,19], episode blob (or 0 elements)
 * PupperPolicyFfi  args: obs f32[n, in]                         rets: action f32[n, out]        (pupper_policy_forward)
 * PupperRolloutFfi args: state blob, DR blob (or 0 elements), episode blob (or 0)
 *                  rets: state blob (alias of arg 0), episode blob (alias of arg 2), obs f32[T,n,H*36], action f32[T,n,12],
 *                        reward f32[T,n], done f32[T,n], metrics f32[n,19] of the last step    (pupper_rollout: the whole unroll
 *                        that brax.training.acting.generate_unroll scans, as one custom call; T and n come from the action result)
 * A result that XLA did not alias to its argument is first filled with a device copy of the argument. */
struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;
struct XLA_FFI_Error *PupperStepFfi(struct XLA_FFI_CallFrame *call_frame);
struct XLA_FFI_Error *PupperResetFfi(struct XLA_FFI_CallFrame *call_frame);
struct XLA_FFI_Error *PupperPolicyFfi(struct XLA_FFI_CallFrame *call_frame);
struct XLA_FFI_Error *PupperRolloutFfi(struct XLA_FFI_CallFrame *call_frame);

#ifdef __cplusplus
}
#endif
#endif /* PUPPER_FFI_H_ */

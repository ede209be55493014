// pupper_ffi.cc -- XLA FFI custom-call handlers over pupper_reset / pupper_step and the blob layouts of include/pupper_ffi.h.
// A thin adapter on purpose: every buffer it touches is caller (XLA) owned, the only state it keeps is the per-device
// registry of model handles, and all real work is the two C-ABI calls.
#include <cuda_runtime.h>
#include <string.h>
#include <mutex>

#ifdef PUPPER_XLA_FFI_HEADER
#include PUPPER_XLA_FFI_HEADER
#else
#include "xla_ffi_stub.h"
#endif
#include "../../include/pupper_ffi.h"
#include "../../include/pupper_policy.h"

namespace {

constexpr int kMaxDevices = 64;
struct Registered { const PupperModel *model; PupperEnvCfg cfg; bool on; };
Registered g_models[kMaxDevices];
struct RegisteredPolicy { const PupperPolicy *policy; int in_dim, out_dim; bool on; };
RegisteredPolicy g_policies[kMaxDevices];
std::mutex g_mu;

constexpr int kDrRows = 1 + 1 + 1 + 3 + 39 + 13;                                     // 58
constexpr int kEpRows = PUPPER_NQ + PUPPER_NV + PUPPER_NV + 1 + 1 + 1 + 1 + PUPPER_NMETRIC + 1;  // 79
constexpr int kTotals = 24;

XLA_FFI_Error *fail(const XLA_FFI_CallFrame *cf, XLA_FFI_Error_Code code, const char *msg) {
  XLA_FFI_Error_Create_Args a;
  memset(&a, 0, sizeof(a));
  a.struct_size = sizeof(a);
  a.message = msg;
  a.errc = code;
  return cf->api->XLA_FFI_Error_Create(&a);
}

int64_t elements(const XLA_FFI_Buffer *b) {
  int64_t n = 1;
  for (int64_t i = 0; i < b->rank; i++) n *= b->dims[i];
  return n;
}

// Looks up the model of the device the call runs on and the stream XLA hands out.
XLA_FFI_Error *context(XLA_FFI_CallFrame *cf, cudaStream_t *stream, Registered *reg) {
  XLA_FFI_Stream_Get_Args s;
  memset(&s, 0, sizeof(s));
  s.struct_size = sizeof(s);
  s.ctx = cf->ctx;
  if (XLA_FFI_Error *e = cf->api->XLA_FFI_Stream_Get(&s)) return e;
  *stream = static_cast<cudaStream_t>(s.stream);
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return fail(cf, XLA_FFI_Error_Code_INTERNAL, "pupper: no current CUDA device");
  std::lock_guard<std::mutex> lock(g_mu);
  if (!g_models[dev].on) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper: no model registered for this device (pupper_ffi_register_model)");
  *reg = g_models[dev];
  return nullptr;
}

// The stream XLA hands out, the device the call runs on and the policy registered for it.
XLA_FFI_Error *policy_context(XLA_FFI_CallFrame *cf, cudaStream_t *stream, RegisteredPolicy *reg) {
  XLA_FFI_Stream_Get_Args s;
  memset(&s, 0, sizeof(s));
  s.struct_size = sizeof(s);
  s.ctx = cf->ctx;
  if (XLA_FFI_Error *e = cf->api->XLA_FFI_Stream_Get(&s)) return e;
  *stream = static_cast<cudaStream_t>(s.stream);
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return fail(cf, XLA_FFI_Error_Code_INTERNAL, "pupper: no current CUDA device");
  std::lock_guard<std::mutex> lock(g_mu);
  if (!g_policies[dev].on) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper: no policy registered for this device (pupper_ffi_register_policy)");
  *reg = g_policies[dev];
  return nullptr;
}

// ret <- arg when XLA did not alias them
XLA_FFI_Error *carry(const XLA_FFI_CallFrame *cf, const XLA_FFI_Buffer *arg, const XLA_FFI_Buffer *ret, cudaStream_t stream) {
  if (elements(arg) != elements(ret)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper: aliased argument / result sizes differ");
  if (arg->data != ret->data && elements(arg) > 0 &&
      cudaMemcpyAsync(ret->data, arg->data, (size_t)elements(arg) * 4, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return fail(cf, XLA_FFI_Error_Code_INTERNAL, "pupper: copying a non-aliased blob failed");
  return nullptr;
}

}  // namespace

extern "C" {

int pupper_blob_stride(int n_envs) { return n_envs <= 0 ? PUPPER_EINVAL : (n_envs + 31) / 32 * 32; }

int64_t pupper_state_blob_words(const PupperEnvCfg *cfg, int n_envs) {
  int32_t rows[14];
  if (!cfg || n_envs <= 0 || pupper_state_rows(cfg, rows) != PUPPER_OK) return PUPPER_EINVAL;
  int64_t r = 0;
  for (int i = 0; i < 14; i++) r += rows[i];
  return r * pupper_blob_stride(n_envs) + (int64_t)n_envs * cfg->observation_history * PUPPER_OBS_DIM;
}
int64_t pupper_dr_blob_words(int n_envs) { return n_envs <= 0 ? PUPPER_EINVAL : (int64_t)kDrRows * pupper_blob_stride(n_envs); }
int64_t pupper_episode_blob_words(const PupperEnvCfg *cfg, int n_envs) {
  if (!cfg || n_envs <= 0) return PUPPER_EINVAL;
  return (int64_t)kEpRows * pupper_blob_stride(n_envs) + (int64_t)n_envs * cfg->observation_history * PUPPER_OBS_DIM + kTotals;
}
int64_t pupper_rand_blob_words(int n_envs) { return n_envs <= 0 ? PUPPER_EINVAL : (int64_t)PUPPER_NRAND * pupper_blob_stride(n_envs); }

int pupper_state_blob_bind(const PupperEnvCfg *cfg, int n_envs, void *blob, PupperState *st) {
  int32_t rows[14];
  if (!cfg || !blob || !st || n_envs <= 0 || pupper_state_rows(cfg, rows) != PUPPER_OK) return PUPPER_EINVAL;
  const int64_t s = pupper_blob_stride(n_envs);
  float *p = static_cast<float *>(blob);
  st->stride = (int32_t)s;
  st->qpos = p; p += rows[0] * s;
  st->qvel = p; p += rows[1] * s;
  st->qacc_warmstart = p; p += rows[2] * s;
  st->rng = reinterpret_cast<uint32_t *>(p); p += rows[3] * s;
  st->last_act = p; p += rows[4] * s;
  st->action_buffer = p; p += rows[5] * s;
  st->imu_buffer = p; p += rows[6] * s;
  st->last_vel = p; p += rows[7] * s;
  st->command = p; p += rows[8] * s;
  st->desired_world_z = p; p += rows[9] * s;
  st->last_contact = reinterpret_cast<uint32_t *>(p); p += rows[10] * s;
  st->feet_air_time = p; p += rows[11] * s;
  st->step = reinterpret_cast<int32_t *>(p); p += rows[12] * s;
  st->kick = p; p += rows[13] * s;
  st->obs = p;
  return PUPPER_OK;
}

int pupper_dr_blob_bind(int n_envs, const void *blob, PupperDR *dr) {
  if (!blob || !dr || n_envs <= 0) return PUPPER_EINVAL;
  const int64_t s = pupper_blob_stride(n_envs);
  const float *p = static_cast<const float *>(blob);
  dr->stride = (int32_t)s;
  dr->friction = p; dr->kp = p + s; dr->kd = p + 2 * s; dr->base_ipos = p + 3 * s; dr->body_inertia = p + 6 * s; dr->body_mass = p + 45 * s;
  return PUPPER_OK;
}

int pupper_episode_blob_bind(const PupperEnvCfg *cfg, int n_envs, void *blob, PupperEpisode *ep) {
  if (!cfg || !blob || !ep || n_envs <= 0) return PUPPER_EINVAL;
  const int64_t s = pupper_blob_stride(n_envs);
  float *p = static_cast<float *>(blob);
  ep->stride = (int32_t)s;
  ep->first_qpos = p; p += PUPPER_NQ * s;
  ep->first_qvel = p; p += PUPPER_NV * s;
  ep->first_warmstart = p; p += PUPPER_NV * s;
  ep->steps = reinterpret_cast<int32_t *>(p); p += s;
  ep->truncation = p; p += s;
  ep->sum_reward = p; p += s;
  ep->length = p; p += s;
  ep->sum_metrics = p; p += PUPPER_NMETRIC * s;
  ep->episode_done = p; p += s;
  ep->first_obs = p; p += (int64_t)n_envs * cfg->observation_history * PUPPER_OBS_DIM;
  ep->totals = p;
  return PUPPER_OK;
}

int pupper_dr_blob_pack(int n_envs, const float *friction, const float *kp, const float *kd, const float *base_ipos, const float *body_inertia,
                        const float *body_mass, float *blob) {
  if (n_envs <= 0 || !friction || !kp || !kd || !base_ipos || !body_inertia || !body_mass || !blob) return PUPPER_EINVAL;
  const int64_t s = pupper_blob_stride(n_envs);
  memset(blob, 0, sizeof(float) * kDrRows * s);
  for (int e = 0; e < n_envs; e++) {
    blob[e] = friction[e]; blob[s + e] = kp[e]; blob[2 * s + e] = kd[e];
    for (int i = 0; i < 3; i++) blob[(3 + i) * s + e] = base_ipos[3 * e + i];
    for (int i = 0; i < 39; i++) blob[(6 + i) * s + e] = body_inertia[39 * e + i];
    for (int i = 0; i < 13; i++) blob[(45 + i) * s + e] = body_mass[13 * e + i];
  }
  return PUPPER_OK;
}

int pupper_ffi_register_model(int device, const PupperModel *model, const PupperEnvCfg *cfg) {
  if (device < 0 || device >= kMaxDevices || !model || !cfg) return PUPPER_EINVAL;
  std::lock_guard<std::mutex> lock(g_mu);
  g_models[device].model = model;
  g_models[device].cfg = *cfg;
  g_models[device].on = true;
  return PUPPER_OK;
}
int pupper_ffi_unregister_model(int device) {
  if (device < 0 || device >= kMaxDevices) return PUPPER_EINVAL;
  std::lock_guard<std::mutex> lock(g_mu);
  g_models[device].on = false;
  return PUPPER_OK;
}

int pupper_ffi_register_policy(int device, const PupperPolicy *policy, int in_dim, int out_dim) {
  if (device < 0 || device >= kMaxDevices || !policy || in_dim < 1 || out_dim < 1) return PUPPER_EINVAL;
  std::lock_guard<std::mutex> lock(g_mu);
  g_policies[device] = RegisteredPolicy{policy, in_dim, out_dim, true};
  return PUPPER_OK;
}
int pupper_ffi_unregister_policy(int device) {
  if (device < 0 || device >= kMaxDevices) return PUPPER_EINVAL;
  std::lock_guard<std::mutex> lock(g_mu);
  g_policies[device].on = false;
  return PUPPER_OK;
}

XLA_FFI_Error *PupperPolicyFfi(XLA_FFI_CallFrame *cf) {
  if (cf->stage != XLA_FFI_ExecutionStage_EXECUTE) return nullptr;
  if (cf->args.size != 1 || cf->rets.size != 1) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_policy: expected 1 argument and 1 result");
  const XLA_FFI_Buffer *obs = static_cast<XLA_FFI_Buffer *>(cf->args.args[0]), *action = static_cast<XLA_FFI_Buffer *>(cf->rets.rets[0]);
  if (obs->rank != 2 || action->rank != 2 || obs->dtype != XLA_FFI_DataType_F32 || action->dtype != XLA_FFI_DataType_F32 || obs->dims[0] != action->dims[0])
    return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_policy: obs must be f32[n, in] and the result f32[n, out]");
  cudaStream_t stream;
  RegisteredPolicy reg;
  if (XLA_FFI_Error *e = policy_context(cf, &stream, &reg)) return e;
  if (obs->dims[1] != reg.in_dim || action->dims[1] != reg.out_dim)
    return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_policy: widths differ from the registered policy's");
  const int rc = pupper_policy_forward(reg.policy, (int)obs->dims[0], static_cast<const float *>(obs->data), static_cast<float *>(action->data), stream);
  return rc == PUPPER_OK ? nullptr : fail(cf, XLA_FFI_Error_Code_INTERNAL, pupper_strerror(rc));
}

XLA_FFI_Error *PupperRolloutFfi(XLA_FFI_CallFrame *cf) {
  if (cf->stage != XLA_FFI_ExecutionStage_EXECUTE) return nullptr;
  if (cf->args.size != 3 || cf->rets.size != 7) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_rollout: expected 3 arguments and 7 results");
  auto arg = [&](int i) { return static_cast<XLA_FFI_Buffer *>(cf->args.args[i]); };
  auto ret = [&](int i) { return static_cast<XLA_FFI_Buffer *>(cf->rets.rets[i]); };
  const XLA_FFI_Buffer *state_in = arg(0), *dr_b = arg(1), *ep_in = arg(2);
  const XLA_FFI_Buffer *state_out = ret(0), *ep_out = ret(1), *t_obs = ret(2), *t_act = ret(3), *t_rew = ret(4), *t_done = ret(5), *metrics = ret(6);
  if (t_act->rank != 3 || t_act->dims[2] != PUPPER_NU || t_act->dtype != XLA_FFI_DataType_F32)
    return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_rollout: the action trajectory must be f32[T, n, 12]");
  const int T = (int)t_act->dims[0], n = (int)t_act->dims[1];
  cudaStream_t stream;
  Registered reg;
  RegisteredPolicy pol;
  if (XLA_FFI_Error *e = context(cf, &stream, &reg)) return e;
  if (XLA_FFI_Error *e = policy_context(cf, &stream, &pol)) return e;
  const int64_t w = (int64_t)reg.cfg.observation_history * PUPPER_OBS_DIM;
  if (T < 1 || n < 1 || elements(t_obs) != (int64_t)T * n * w || elements(t_rew) != (int64_t)T * n || elements(t_done) != (int64_t)T * n ||
      elements(metrics) != (int64_t)n * PUPPER_NMETRIC)
    return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_rollout: trajectory / metrics results have the wrong size");
  if (elements(state_in) != pupper_state_blob_words(&reg.cfg, n)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_rollout: state blob has the wrong size");
  if (XLA_FFI_Error *e = carry(cf, state_in, state_out, stream)) return e;
  PupperState st;
  pupper_state_blob_bind(&reg.cfg, n, state_out->data, &st);
  PupperDR dr, *drp = nullptr;
  if (elements(dr_b) > 0) {
    if (elements(dr_b) != pupper_dr_blob_words(n)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_rollout: DR blob has the wrong size");
    pupper_dr_blob_bind(n, dr_b->data, &dr);
    drp = &dr;
  }
  PupperEpisode ep, *epp = nullptr;
  if (elements(ep_in) > 0) {
    if (elements(ep_in) != pupper_episode_blob_words(&reg.cfg, n)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_rollout: episode blob has the wrong size");
    if (XLA_FFI_Error *e = carry(cf, ep_in, ep_out, stream)) return e;
    pupper_episode_blob_bind(&reg.cfg, n, ep_out->data, &ep);
    epp = &ep;
  }
  PupperStepOut out;
  memset(&out, 0, sizeof(out));
  float *rew = static_cast<float *>(t_rew->data), *dn = static_cast<float *>(t_done->data);
  out.reward = rew + (size_t)(T - 1) * n;  // "the last step's" single-step outputs are the last trajectory slices themselves
  out.done = dn + (size_t)(T - 1) * n;
  out.metrics = static_cast<float *>(metrics->data);
  const int rc = pupper_rollout(reg.model, pol.policy, n, T, drp, &st, &out, epp, static_cast<float *>(t_obs->data), static_cast<float *>(t_act->data),
                                rew, dn, stream);
  return rc == PUPPER_OK ? nullptr : fail(cf, XLA_FFI_Error_Code_INTERNAL, pupper_strerror(rc));
}

XLA_FFI_Error *PupperStepFfi(XLA_FFI_CallFrame *cf) {
  if (cf->stage != XLA_FFI_ExecutionStage_EXECUTE) return nullptr;
  if (cf->args.size != 5 || cf->rets.size != 5) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_step: expected 5 arguments and 5 results");
  auto arg = [&](int i) { return static_cast<XLA_FFI_Buffer *>(cf->args.args[i]); };
  auto ret = [&](int i) { return static_cast<XLA_FFI_Buffer *>(cf->rets.rets[i]); };
  const XLA_FFI_Buffer *action = arg(0), *state_in = arg(1), *dr_b = arg(2), *ep_in = arg(3), *rand_b = arg(4);
  const XLA_FFI_Buffer *state_out = ret(0), *reward = ret(1), *done = ret(2), *metrics = ret(3), *ep_out = ret(4);
  if (action->rank != 2 || action->dims[1] != PUPPER_NU || action->dtype != XLA_FFI_DataType_F32)
    return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_step: action must be f32[n, 12]");
  const int n = (int)action->dims[0];
  cudaStream_t stream;
  Registered reg;
  if (XLA_FFI_Error *e = context(cf, &stream, &reg)) return e;
  if (elements(state_in) != pupper_state_blob_words(&reg.cfg, n)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_step: state blob has the wrong size");
  if (elements(reward) != n || elements(done) != n || elements(metrics) != (int64_t)n * PUPPER_NMETRIC)
    return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_step: reward / done / metrics results have the wrong size");
  if (XLA_FFI_Error *e = carry(cf, state_in, state_out, stream)) return e;
  PupperState st;
  pupper_state_blob_bind(&reg.cfg, n, state_out->data, &st);
  PupperDR dr, *drp = nullptr;
  if (elements(dr_b) > 0) {
    if (elements(dr_b) != pupper_dr_blob_words(n)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_step: DR blob has the wrong size");
    pupper_dr_blob_bind(n, dr_b->data, &dr);
    drp = &dr;
  }
  PupperEpisode ep, *epp = nullptr;
  if (elements(ep_in) > 0) {
    if (elements(ep_in) != pupper_episode_blob_words(&reg.cfg, n)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_step: episode blob has the wrong size");
    if (XLA_FFI_Error *e = carry(cf, ep_in, ep_out, stream)) return e;
    pupper_episode_blob_bind(&reg.cfg, n, ep_out->data, &ep);
    epp = &ep;
  }
  PupperRand rnd, *rp = nullptr;
  if (elements(rand_b) > 0) {
    if (elements(rand_b) != pupper_rand_blob_words(n)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_step: rand blob has the wrong size");
    rnd.stride = pupper_blob_stride(n);
    rnd.u = static_cast<const float *>(rand_b->data);
    rp = &rnd;
  }
  PupperStepOut out;
  memset(&out, 0, sizeof(out));
  out.reward = static_cast<float *>(reward->data);
  out.done = static_cast<float *>(done->data);
  out.metrics = static_cast<float *>(metrics->data);
  const int rc = pupper_step(reg.model, n, drp, &st, static_cast<const float *>(action->data), rp, &out, epp, stream);
  return rc == PUPPER_OK ? nullptr : fail(cf, XLA_FFI_Error_Code_INTERNAL, pupper_strerror(rc));
}

XLA_FFI_Error *PupperResetFfi(XLA_FFI_CallFrame *cf) {
  if (cf->stage != XLA_FFI_ExecutionStage_EXECUTE) return nullptr;
  if (cf->args.size != 3 || cf->rets.size != 5) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_reset: expected 3 arguments and 5 results");
  auto arg = [&](int i) { return static_cast<XLA_FFI_Buffer *>(cf->args.args[i]); };
  auto ret = [&](int i) { return static_cast<XLA_FFI_Buffer *>(cf->rets.rets[i]); };
  const XLA_FFI_Buffer *keys = arg(0), *dr_b = arg(1), *rand_b = arg(2);
  const XLA_FFI_Buffer *state_out = ret(0), *reward = ret(1), *done = ret(2), *metrics = ret(3), *ep_out = ret(4);
  if (keys->rank != 2 || keys->dims[1] != 2 || keys->dtype != XLA_FFI_DataType_U32)
    return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_reset: keys must be u32[n, 2] (jax.random.key_data)");
  const int n = (int)keys->dims[0];
  cudaStream_t stream;
  Registered reg;
  if (XLA_FFI_Error *e = context(cf, &stream, &reg)) return e;
  if (elements(state_out) != pupper_state_blob_words(&reg.cfg, n)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_reset: state blob has the wrong size");
  if (elements(reward) != n || elements(done) != n || elements(metrics) != (int64_t)n * PUPPER_NMETRIC)
    return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_reset: reward / done / metrics results have the wrong size");
  PupperState st;
  pupper_state_blob_bind(&reg.cfg, n, state_out->data, &st);
  PupperDR dr, *drp = nullptr;
  if (elements(dr_b) > 0) {
    if (elements(dr_b) != pupper_dr_blob_words(n)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_reset: DR blob has the wrong size");
    pupper_dr_blob_bind(n, dr_b->data, &dr);
    drp = &dr;
  }
  PupperEpisode ep, *epp = nullptr;
  if (elements(ep_out) > 0) {
    if (elements(ep_out) != pupper_episode_blob_words(&reg.cfg, n)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_reset: episode blob has the wrong size");
    pupper_episode_blob_bind(&reg.cfg, n, ep_out->data, &ep);
    if (cudaMemsetAsync(ep.totals, 0, sizeof(float) * kTotals, stream) != cudaSuccess) return fail(cf, XLA_FFI_Error_Code_INTERNAL, "pupper_reset: clearing the totals failed");
    epp = &ep;
  }
  PupperRand rnd, *rp = nullptr;
  if (elements(rand_b) > 0) {
    if (elements(rand_b) != pupper_rand_blob_words(n)) return fail(cf, XLA_FFI_Error_Code_INVALID_ARGUMENT, "pupper_reset: rand blob has the wrong size");
    rnd.stride = pupper_blob_stride(n);
    rnd.u = static_cast<const float *>(rand_b->data);
    rp = &rnd;
  }
  PupperStepOut out;
  memset(&out, 0, sizeof(out));
  out.reward = static_cast<float *>(reward->data);
  out.done = static_cast<float *>(done->data);
  out.metrics = static_cast<float *>(metrics->data);
  const int rc = pupper_reset(reg.model, n, static_cast<const uint32_t *>(keys->data), drp, &st, &out, epp, rp, stream);
  return rc == PUPPER_OK ? nullptr : fail(cf, XLA_FFI_Error_Code_INTERNAL, pupper_strerror(rc));
}

}  // extern "C"

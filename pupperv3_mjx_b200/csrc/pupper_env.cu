// pupper_env.cu -- env-level kernel (reset / step) and the C ABI declared in include/pupper_env.h.
// Env level = reference pupperv3_mjx/environment.py:314-543 + rewards.py:9-138 + utils.py:49-69;
// physics = pupper_kernel.cuh.  No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <new>
#include <algorithm>

#include "pupper_kernel.cuh"

namespace pupper {

__device__ __forceinline__ uint2 qbcast2(uint2 v, int src, unsigned qm, int qbase) {
  return make_uint2(__shfl_sync(qm, v.x, qbase + src), __shfl_sync(qm, v.y, qbase + src));
}

// jax.random.choice index from the uniform draw u: searchsorted(cumsum(p), cumsum(p)[-1] * (1 - u))
// (out of line, rolled loops: env-level code is kept small, see threefry2x32)
__device__ __noinline__ int choice_from_u(float u, const float *p, int n) {
  float total = 0.f;
#pragma unroll 1
  for (int i = 0; i < n; i++) total = __fadd_rn(total, p[i]);
  const float r = __fmul_rn(total, __fsub_rn(1.0f, u));
  float acc = 0.f;
  int idx = 0;
#pragma unroll 1
  for (int i = 0; i < n; i++) { acc = __fadd_rn(acc, p[i]); idx += (acc < r) ? 1 : 0; }  // same running sums as cumsum
  return min(idx, n - 1);
}

// utils.sample_lagged_value on one row of a lag buffer held in global memory (row elements `stride` apart):
// push `newest` at the front, return the element at column `pick` of the updated row.
// Batched form: all loads are issued before the first store so they overlap (the compiler cannot reorder them itself: the
// pointers may alias).  A rolled, out-of-line form is 190 instructions smaller but serialises the loads: measured -2.6 % at
// 65,536 envs, so the batched form stays.
__device__ __forceinline__ float push_front_pick(float *row, int stride, int L, float newest, int pick, bool fresh, float fresh_value, bool valid) {
  float old[PUPPER_MAX_LAT - 1];
#pragma unroll
  for (int l = 0; l < PUPPER_MAX_LAT - 1; l++) old[l] = (l < L - 1) ? (fresh ? fresh_value : row[(size_t)l * stride]) : 0.f;
  float lag = newest;
#pragma unroll
  for (int l = 0; l < PUPPER_MAX_LAT - 1; l++) if (l + 1 == pick) lag = old[l];
  if (valid) {
    row[0] = newest;
#pragma unroll
    for (int l = 0; l < PUPPER_MAX_LAT - 1; l++) if (l < L - 1) row[(size_t)(l + 1) * stride] = old[l];
  }
  return lag;
}
// jax.random.uniform's affine map of a raw [0, 1) draw (three separately rounded operations, as in `uniform`)
__device__ __forceinline__ float affine_u(float f, float lo, float hi) {
  return fmaxf(lo, __fadd_rn(__fmul_rn(f, __fsub_rn(hi, lo)), lo));
}
// One uniform draw: from the external-randoms table (row `row` of PupperRand, this env's column `ext`, rows `es` apart)
// when one is given, else element `idx` of jax.random.uniform(key, ...) (threefry).
__device__ __forceinline__ float draw(const float *ext, int es, int row, uint2 key, uint32_t idx, float lo, float hi) {
  return ext ? affine_u(ext[(size_t)row * es], lo, hi) : uniform(key, idx, lo, hi);
}

// environment.py:246-272 (evaluated by one lane)
__device__ __noinline__ void sample_command(const PupperEnvCfg &c, uint2 key, float cmd[3], const float *ext, int es) {
  uint2 k1 = split_key(key, 1), k2 = split_key(key, 2), k3 = split_key(key, 3), k4 = split_key(key, 4), k5 = split_key(key, 5);
  float c0 = draw(ext, es, 35, k1, 0, c.lin_vel_x[0], c.lin_vel_x[1]);
  float c1 = draw(ext, es, 36, k2, 0, c.lin_vel_y[0], c.lin_vel_y[1]);
  float c2 = draw(ext, es, 37, k3, 0, c.ang_vel_yaw[0], c.ang_vel_yaw[1]);
  float zp = draw(ext, es, 38, k4, 0, 0.f, 1.f);
  float thr = c.stand_still_command_threshold;
  if (zp < c.zero_command_probability) {
    cmd[0] = draw(ext, es, 39, k5, 0, -thr, thr); cmd[1] = draw(ext, es, 40, k5, 1, -thr, thr); cmd[2] = draw(ext, es, 41, k5, 2, -thr, thr);
  } else { cmd[0] = c0; cmd[1] = c1; cmd[2] = c2; }
}

// environment.py:274-298 (evaluated by one lane)
__device__ __noinline__ void sample_body_orientation(const PupperEnvCfg &c, uint2 key, float out[3], const float *ext, int es) {
  uint2 kp = split_key(key, 1), kr = split_key(key, 2);
  float pitch = __fmul_rn(draw(ext, es, 42, kp, 0, -1.f, 1.f), c.maximum_pitch_command);
  float roll = __fmul_rn(draw(ext, es, 43, kr, 0, -1.f, 1.f), c.maximum_roll_command);
  const float pi = 3.14159274101257324f;
  float a1 = roll * pi / 360.f, a2 = pitch * pi / 360.f;
  float s1, c1, s2, c2;
  sincosf(a1, &s1, &c1);
  sincosf(a2, &s2, &c2);
  const float s3 = 0.f, c3 = 1.f;
  Q4 q = Q4{c1 * c2 * c3 - s1 * s2 * s3, s1 * c2 * c3 + c1 * s2 * s3, c1 * s2 * c3 - s1 * c2 * s3, c1 * c2 * s3 + s1 * s2 * c3};
  V3 z = rotate(V3{c.desired_world_z_in_body_frame[0], c.desired_world_z_in_body_frame[1], c.desired_world_z_in_body_frame[2]}, q);
  out[0] = z.x; out[1] = z.y; out[2] = z.z;
}

struct ObsCtx {
  float command[3], desired_z[3];
  float last_act[3];   // this leg's channels
  float ql[3];
};

// _get_obs (environment.py:485-543). `rng` is info["rng"] on entry; returns the new info["rng"].
// Updates the IMU buffer and the observation history in global memory.
__device__ __forceinline__ uint2 get_obs(const BlockShared &sh, const KParams &p, int e, int k, unsigned qm, int qbase, uint2 rng,
                                         const StaleOut &so, const ObsCtx &oc, bool zero_history, bool valid, const float *ext, int es) {
  const PupperEnvCfg &c = sh.c;
  const int stride = p.st.stride;
  // split(rng, 6): lane 0 -> new rng, lane 1 -> ang key, lane 2 -> gravity key, lane 3 -> imu sample key
  uint2 kmine = split_key(rng, k == 3 ? 5u : (uint32_t)k);
  uint2 k_motor = split_key(rng, 3u), k_act = split_key(rng, 4u);
  uint2 new_rng = qbcast2(kmine, 0, qm, qbase), k_ang = qbcast2(kmine, 1, qm, qbase), k_grav = qbcast2(kmine, 2, qm, qbase);
  // one draw per lane: lanes 0-2 angular-velocity noise component k, lane 3 the IMU latency draw
  float uA = draw(ext, es, k < 3 ? 4 + k : 34, k < 3 ? k_ang : kmine, k < 3 ? (uint32_t)k : 0u, k < 3 ? -1.f : 0.f, 1.f);
  float uB = draw(ext, es, k < 3 ? 7 + k : 7, k_grav, (uint32_t)k, -1.f, 1.f);
  float an_mine = __fmul_rn(uA, c.angular_velocity_noise), gn_mine = __fmul_rn(uB, c.gravity_noise);
  float an[3], gn[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { an[i] = __shfl_sync(qm, an_mine, qbase + i); gn[i] = __shfl_sync(qm, gn_mine, qbase + i); }
  int imu_idx = k == 3 ? choice_from_u(uA, c.imu_latency_distribution, c.n_imu_latency) : 0;
  imu_idx = __shfl_sync(qm, imu_idx, qbase + 3);

  Q4 inv = Q4{1.f, 0.f, 0.f, 0.f};
  V3 ang = V3{0.f, 0.f, 0.f};
  if (c.use_imu) { inv = qinv(so.torso_rot); ang = rotate(so.torso_ang, inv); }
  V3 g = rotate(V3{0.f, 0.f, -1.f}, inv);
  g = V3{g.x + gn[0], g.y + gn[1], g.z + gn[2]};
  float gnorm = sqrtf(g.x * g.x + g.y * g.y + g.z * g.z);
  float imu[6] = {ang.x + an[0], ang.y + an[1], ang.z + an[2], g.x / gnorm, g.y / gnorm, g.z / gnorm};

  // IMU lag buffer (6, L): this lane handles rows k and k+4
  const int Li = c.n_imu_latency;
  float lag_imu[2] = {0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 2; t++) {
    int row = k + 4 * t;
    if (row < 6) {
      float newest = 0.f;
#pragma unroll
      for (int i = 0; i < 6; i++) if (i == row) newest = imu[i];
      lag_imu[t] = push_front_pick(p.st.imu_buffer + (size_t)(row * Li) * stride + e, stride, Li, newest, imu_idx, zero_history,
                                   row == 5 ? -1.f : 0.f, valid);
    }
  }

  // history roll: slot h <- slot h-1 (each lane moves the entries i = k (mod 4))
  const int H = c.observation_history;
  float *obs = p.st.obs + (size_t)e * H * PUPPER_OBS_DIM;
  for (int h = H - 1; h >= 1; h--) {
    float tmp[PUPPER_OBS_DIM / 4];
#pragma unroll
    for (int t = 0; t < PUPPER_OBS_DIM / 4; t++) tmp[t] = zero_history ? 0.f : obs[(h - 1) * PUPPER_OBS_DIM + k + 4 * t];
    if (valid) {
#pragma unroll
      for (int t = 0; t < PUPPER_OBS_DIM / 4; t++) obs[h * PUPPER_OBS_DIM + k + 4 * t] = tmp[t];
    }
  }
  __syncwarp(qm);
  auto clip100 = [](float x) { return fminf(fmaxf(x, -100.f), 100.f); };
  // entries 0..11: imu(6), command(3), desired_z(3); this lane writes i = k, k+4, k+8
  {
    if (valid) obs[k] = clip100(lag_imu[0]);
    int i1 = k + 4;
    float v1 = i1 < 6 ? lag_imu[1] : (i1 == 6 ? oc.command[0] : oc.command[1]);
    if (valid) obs[i1] = clip100(v1);
    int i2 = k + 8;
    float v2 = i2 == 8 ? oc.command[2] : (i2 == 9 ? oc.desired_z[0] : (i2 == 10 ? oc.desired_z[1] : oc.desired_z[2]));
    if (valid) obs[i2] = clip100(v2);
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const int u = 3 * k + j;
    float mn = __fmul_rn(draw(ext, es, 10 + u, k_motor, (uint32_t)u, -1.f, 1.f), c.motor_angle_noise);
    float ln = __fmul_rn(draw(ext, es, 22 + u, k_act, (uint32_t)u, -1.f, 1.f), c.last_action_noise);
    if (valid) obs[12 + u] = clip100(oc.ql[j] - c.default_pose[u] + mn);
    if (valid) obs[24 + u] = clip100(oc.last_act[j] + ln);
  }
  return new_rng;
}

#ifndef PUPPER_MIN_BLOCKS
#define PUPPER_MIN_BLOCKS 2
#endif
#ifdef PUPPER_NO_BOUNDS  // experiments: let -maxrregcount decide
#define PUPPER_LB
#else
#define PUPPER_LB __launch_bounds__(kBlock, PUPPER_MIN_BLOCKS)
#endif
// Per-call pointers of one env step: what a single-step launch takes from KParams, and what the persistent rollout kernel
// (pupper_rollout.cuh) moves from step to step of its unroll: there the three pointers are the [T][n][...] trajectory bases
// and `t` selects the slice.
struct StepIO {
  const float *action;  // [n][12]
  float *reward, *done; // [n]
  int t;                // step of the unroll (0 for single-step launches)
};

// The whole reset / step of the CTA's 32 envs (CTA index `block`, thread index `tid`).  STAGED: the model constants, the per-env
// DR leaves and the cleared contact slots are already in shared memory (the rollout kernel stages them itself, ahead of its
// wait for the previous step), and io.t selects the trajectory slice.
template <bool RESET, bool DBG, bool STAGED>
__device__ __forceinline__ void env_body(const KParams &p, BlockShared &sh, const StepIO io, const int block, const int tid) {
  if (!STAGED) {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.consts);
    uint4 *dst = reinterpret_cast<uint4 *>(static_cast<ConstBlock *>(&sh));
    constexpr int n16 = (int)(sizeof(ConstBlock) / 16);
#pragma unroll
    for (int i0 = 0; i0 < n16; i0 += kBlock) {
      const int i = i0 + tid;
      if (i < n16) dst[i] = __ldg(src + i);
    }
    __syncthreads();
  }
  const int lane = tid & 31, k = tid & 3;
  const int el = tid >> 2;
  // Every shuffle uses the full warp mask, so all 32 lanes stay convergent: quads past the end of the
  // batch recompute the last env (identical values) and their lanes write nothing different.
  const bool valid = block * kEnvsPerBlock + el < p.n_envs;
  const int e = min(block * kEnvsPerBlock + el, p.n_envs - 1);
  const unsigned qm = 0xffffffffu;
  const int qbase = lane & 28;
  const PupperModelDesc &m = sh.m;
  const PupperEnvCfg &c = sh.c;
  EnvShared &es = sh.env[el];
  const int stride = p.st.stride;
  // slice `t` of a trajectory array (single-step launches: t = 0, the arrays are the step's own)
  auto at_step = [&](auto *base, int per_step) { return STAGED ? base + (size_t)io.t * per_step : base; };
  // external randoms: only the debug / parity instantiation looks at them (pupper_step routes calls that pass them there)
  const float *ext = (DBG && p.has_rand) ? p.rand.u + e : nullptr;
  const int ext_s = (DBG && p.has_rand) ? p.rand.stride : 0;

  // ---- stage the per-env DR leaves (or the nominal values) in shared memory --------------------------
  // Rolled loops on purpose (code size); the per-env leaves travel global -> shared as 4-byte async copies, so the
  // up-to-15 loads of a lane are all in flight at once without passing through registers.
  if (!STAGED) {
    float *ef = reinterpret_cast<float *>(&es);  // mass[13] inertia[39] ipos[3] friction kp kd are contiguous
    if (p.has_dr) {
      const int ds = p.dr.stride;
#pragma unroll 1
      for (int i = k; i < 58; i += 4) {
        const float *src = i < 13 ? p.dr.body_mass + (size_t)i * ds
                         : i < 52 ? p.dr.body_inertia + (size_t)(i - 13) * ds
                         : i < 55 ? p.dr.base_ipos + (size_t)(i - 52) * ds
                         : i == 55 ? p.dr.friction : (i == 56 ? p.dr.kp : p.dr.kd);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(ef + i)), "l"(src + e) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
#pragma unroll 1
      for (int i = k; i < 58; i += 4)
        ef[i] = i < 13 ? m.body_mass[1 + i] : i < 52 ? m.body_inertia[1 + (i - 13) / 3][(i - 13) % 3]
              : i < 55 ? m.body_ipos[1][i - 52] : i == 55 ? -1.f : (i == 56 ? m.act_gain[0] : -m.act_bias2[0]);
    }
  }
  // Contact slots are read (and multiplied by zero weights) by quads that have fewer contacts than the warp
  // maximum, so they must never hold non-finite garbage: clear them once per launch.
  if (!STAGED) {
    float *cz = reinterpret_cast<float *>(es.con);
    for (int i = k; i < (int)(sizeof(es.con) / 4); i += 4) cz[i] = 0.f;
    if (k == 0) es.ncon = 0;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp(qm);
  }

  LaneState L;
  StaleOut so;
  DbgOut dbg;
  ObsCtx oc;
  float ab[6], al[3];
  uint2 rng;
  const int H = c.observation_history;

  if (RESET) {
    // ---- reset (environment.py:314-346), replicated over the quad where it is per-env ------------------
    uint2 key = make_uint2(p.keys[2 * e], p.keys[2 * e + 1]);
    uint2 k_rng = split_key(key, 0), k_cmd = split_key(key, 1), k_ori = split_key(key, 2), k_pos = split_key(key, 3);
    uint2 kp = split_key(k_pos, 1), ky = split_key(k_pos, 2);
#pragma unroll
    for (int i = 0; i < 7; i++) L.qb[i] = c.init_q[i];
#pragma unroll
    for (int i = 0; i < 3; i++) L.qb[i] = draw(ext, ext_s, i, kp, (uint32_t)i, c.start_pos_min[i], c.start_pos_max[i]);
    {
      float yaw = draw(ext, ext_s, 3, ky, 0u, -3.14159274101257324f, 3.14159274101257324f);
      float s, cs;
      sincosf(yaw / 2.f, &s, &cs);
      L.qb[3] = cs; L.qb[4] = 0.f; L.qb[5] = 0.f; L.qb[6] = s;
    }
#pragma unroll
    for (int j = 0; j < 3; j++) { L.ql[j] = c.init_q[7 + 3 * k + j]; L.vl[j] = 0.f; L.wl[j] = 0.f; L.ctrl[j] = 0.f; }
#pragma unroll
    for (int d = 0; d < 6; d++) { L.vb[d] = 0.f; L.wb[d] = 0.f; }
    forward<DBG>(sh, es, sh.rows, L, k, qm, qbase, ab, al, true, &dbg, tid);
    __syncwarp(qm);
    so = load_stale(es, k);  // pipeline_init = make_data + forward
#pragma unroll
    for (int d = 0; d < 6; d++) L.wb[d] = ab[d];
#pragma unroll
    for (int j = 0; j < 3; j++) L.wl[j] = al[j];
    float cmd[3] = {0.f, 0.f, 0.f}, dz[3] = {0.f, 0.f, 0.f};
    if (k == 0) { sample_command(c, k_cmd, cmd, ext, ext_s); sample_body_orientation(c, k_ori, dz, ext, ext_s); }
#pragma unroll
    for (int i = 0; i < 3; i++) { oc.command[i] = __shfl_sync(qm, cmd[i], qbase); oc.desired_z[i] = __shfl_sync(qm, dz[i], qbase); }
#pragma unroll
    for (int j = 0; j < 3; j++) { oc.last_act[j] = 0.f; oc.ql[j] = L.ql[j]; }
    rng = get_obs(sh, p, e, k, qm, qbase, k_rng, so, oc, true, valid, ext, ext_s);
    if (ext) rng = key;  // external randoms: info["rng"] is never consumed; it keeps the env's key
    // info / state
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int u = 3 * k + j;
      if (valid) p.st.last_act[(size_t)u * stride + e] = 0.f;
      if (valid) p.st.last_vel[(size_t)u * stride + e] = 0.f;
      for (int l = 0; l < c.n_latency; l++) if (valid) p.st.action_buffer[(size_t)(u * c.n_latency + l) * stride + e] = 0.f;
    }
    if (valid) p.st.feet_air_time[(size_t)k * stride + e] = 0.f;
    if (k == 0) {
      if (valid) p.st.last_contact[e] = 0u; if (valid) p.st.step[e] = 0;
      if (valid) p.st.kick[e] = 0.f; if (valid) p.st.kick[stride + e] = 0.f;
      if (valid) io.reward[e] = 0.f; if (valid) io.done[e] = 0.f;
    }
    for (int i = k; i < PUPPER_NMETRIC; i += 4) if (valid) p.out.metrics[(size_t)e * PUPPER_NMETRIC + i] = 0.f;
  } else {
    // ---- step (environment.py:348-483) ---------------------------------------------------------------------
    // Everything the epilogue reads (lag buffers, obs history, bookkeeping, episode state) is cold in L2/DRAM and
    // would be fetched one dependent round trip at a time after the physics: start those fetches now.
    {
      auto pf = [](const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); };
      const int Li = c.n_imu_latency;
#pragma unroll 1
      for (int r = k; r < 6; r += 4)
#pragma unroll 1
        for (int l = 0; l < Li; l++) pf(p.st.imu_buffer + (size_t)(r * Li + l) * stride + e);
#pragma unroll 1
      for (int i = k * 32; i < H * PUPPER_OBS_DIM; i += 128) pf(p.st.obs + (size_t)e * H * PUPPER_OBS_DIM + i);
#pragma unroll 1
      for (int j = 0; j < 3; j++) { pf(p.st.last_vel + (size_t)(3 * k + j) * stride + e); pf(p.st.last_act + (size_t)(3 * k + j) * stride + e); }
      if (k < 3) { pf(p.st.command + (size_t)k * stride + e); pf(p.st.desired_world_z + (size_t)k * stride + e); }
      pf(p.st.feet_air_time + (size_t)k * stride + e);
      if (k == 0) { pf(p.st.last_contact + e); pf(p.st.step + e); }
      if (p.has_ep) {
        const int es_ = p.ep.stride;
#pragma unroll 1
        for (int i = k; i < PUPPER_NMETRIC; i += 4) pf(p.ep.sum_metrics + (size_t)i * es_ + e);
        if (k == 1) { pf(p.ep.episode_done + e); pf(p.ep.steps + e); }
        if (k == 2) { pf(p.ep.sum_reward + e); pf(p.ep.length + e); }
      }
    }
    rng = make_uint2(p.st.rng[e], p.st.rng[stride + e]);
#pragma unroll
    for (int i = 0; i < 7; i++) L.qb[i] = p.st.qpos[(size_t)i * stride + e];
#pragma unroll
    for (int d = 0; d < 6; d++) { L.vb[d] = p.st.qvel[(size_t)d * stride + e]; L.wb[d] = p.st.qacc_warmstart[(size_t)d * stride + e]; }
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int u = 3 * k + j;
      L.ql[j] = p.st.qpos[(size_t)(7 + u) * stride + e];
      L.vl[j] = p.st.qvel[(size_t)(6 + u) * stride + e];
      L.wl[j] = p.st.qacc_warmstart[(size_t)(6 + u) * stride + e];
    }
  }

  float act[3] = {0.f, 0.f, 0.f};
  float kick0 = 0.f, kick1 = 0.f;
  uint2 cmd_rng = make_uint2(0u, 0u);
  if (!RESET) {
    // S1 split(rng, 5): lane k computes key k; lane 3 also needs key 4 (latency)
    uint2 kmine = split_key(rng, (uint32_t)k);
    uint2 k4 = split_key(rng, 4u);
    uint2 K0 = qbcast2(kmine, 0, qm, qbase), K1 = qbcast2(kmine, 1, qm, qbase), K2 = qbcast2(kmine, 2, qm, qbase);
    cmd_rng = K1;
    // S2/S3 draws, one per lane: kick x, kick y, kick Bernoulli, latency pick
    // (lane 0: kick x, lane 1: kick y, lane 2: latency pick with key 4, lane 3: Bernoulli with its own key 3)
    float u = draw(ext, ext_s, k < 2 ? k : (k == 2 ? 3 : 2), k < 2 ? K2 : (k == 2 ? k4 : kmine), k == 1 ? 1u : 0u, k < 2 ? -1.f : 0.f, 1.f);
    float kv = __fmul_rn(u, c.kick_vel);
    float hit = __shfl_sync(qm, (u < c.kick_probability) ? 1.f : 0.f, qbase + 3);
    kick0 = __fmul_rn(__shfl_sync(qm, kv, qbase + 0), hit);
    kick1 = __fmul_rn(__shfl_sync(qm, kv, qbase + 1), hit);
    int aidx = k == 2 ? choice_from_u(u, c.latency_distribution, c.n_latency) : 0;
    aidx = __shfl_sync(qm, aidx, qbase + 2);
    L.vb[0] = kick0 + L.vb[0];
    L.vb[1] = kick1 + L.vb[1];
    // action latency buffer (12, L): push front, take column aidx; S4 motor targets
    const int La = c.n_latency;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int u_ = 3 * k + j;
      act[j] = at_step(io.action, p.n_envs * PUPPER_NU)[(size_t)e * PUPPER_NU + u_];
      float lag = push_front_pick(p.st.action_buffer + (size_t)(u_ * La) * stride + e, stride, La, act[j], aidx, false, 0.f, valid);
      float t = c.default_pose[u_] + lag * c.action_scale;
      L.ctrl[j] = fminf(fmaxf(t, c.joint_lower[u_]), c.joint_upper[u_]);
      es.lv_act[u_] = act[j];
    }
    if (k == 0) { es.lv_kick[0] = kick0; es.lv_kick[1] = kick1; es.lv_cmd_rng[0] = cmd_rng.x; es.lv_cmd_rng[1] = cmd_rng.y; }
    rng = K0;
    // S5 physics: n_frames x (forward ; semi-implicit Euler)
    const float dt = m.timestep;
    for (int f = 0; f < c.n_frames; f++) {
      forward<DBG>(sh, es, sh.rows, L, k, qm, qbase, ab, al, f == c.n_frames - 1, &dbg, tid);
#pragma unroll
      for (int d = 0; d < 6; d++) { L.wb[d] = ab[d]; L.vb[d] = fmaf(ab[d], dt, L.vb[d]); }
#pragma unroll
      for (int j = 0; j < 3; j++) { L.wl[j] = al[j]; L.vl[j] = fmaf(al[j], dt, L.vl[j]); L.ql[j] = fmaf(dt, L.vl[j], L.ql[j]); }
      L.qb[0] = fmaf(dt, L.vb[0], L.qb[0]); L.qb[1] = fmaf(dt, L.vb[1], L.qb[1]); L.qb[2] = fmaf(dt, L.vb[2], L.qb[2]);
      V3 w = V3{L.vb[3], L.vb[4], L.vb[5]};
      float nrm = normalize3(w);
      float sn, cs;
      sincos_small(dt * nrm * 0.5f, &sn, &cs);
      Q4 qn = qnormalize(qmul(Q4{L.qb[3], L.qb[4], L.qb[5], L.qb[6]}, Q4{cs, w.x * sn, w.y * sn, w.z * sn}));
      L.qb[3] = qn.w; L.qb[4] = qn.x; L.qb[5] = qn.y; L.qb[6] = qn.z;
    }
    __syncwarp(qm);
    so = load_stale(es, k);
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int u = 3 * k + j;
      oc.ql[j] = L.ql[j];
      act[j] = es.lv_act[u];
      oc.last_act[j] = p.st.last_act[(size_t)u * stride + e];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) { oc.command[i] = p.st.command[(size_t)i * stride + e]; oc.desired_z[i] = p.st.desired_world_z[(size_t)i * stride + e]; }
    kick0 = es.lv_kick[0]; kick1 = es.lv_kick[1];
    cmd_rng = make_uint2(es.lv_cmd_rng[0], es.lv_cmd_rng[1]);
    // S6 observation (reads the not-yet-updated last_act / command / desired_z)
    rng = get_obs(sh, p, e, k, qm, qbase, rng, so, oc, false, valid, ext, ext_s);
  }

  // ---- write back the physics state -------------------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const int u = 3 * k + j;
    if (valid) p.st.qpos[(size_t)(7 + u) * stride + e] = L.ql[j];
    if (valid) p.st.qvel[(size_t)(6 + u) * stride + e] = L.vl[j];
    if (valid) p.st.qacc_warmstart[(size_t)(6 + u) * stride + e] = L.wl[j];
  }
  for (int i = k; i < 7; i += 4) {
    float v = 0.f;
#pragma unroll
    for (int t = 0; t < 7; t++) if (t == i) v = L.qb[t];
    if (valid) p.st.qpos[(size_t)i * stride + e] = v;
  }
  for (int i = k; i < 6; i += 4) {
    float v = 0.f, w = 0.f;
#pragma unroll
    for (int t = 0; t < 6; t++) if (t == i) { v = L.vb[t]; w = L.wb[t]; }
    if (valid) p.st.qvel[(size_t)i * stride + e] = v;
    if (valid) p.st.qacc_warmstart[(size_t)i * stride + e] = w;
  }
  if (k == 0 && (RESET || !ext)) { if (valid) p.st.rng[e] = rng.x; if (valid) p.st.rng[stride + e] = rng.y; }  // a step with external randoms leaves the key alone

  if (DBG) {
    if (p.out.dbg_x_pos) {
      float *xp = p.out.dbg_x_pos + (size_t)e * 39, *xr = p.out.dbg_x_rot + (size_t)e * 52;
      float *xv = p.out.dbg_xd_vel + (size_t)e * 39, *xa = p.out.dbg_xd_ang + (size_t)e * 39;
      if (k == 0) {
        if (valid) xp[0] = so.torso_pos.x; if (valid) xp[1] = so.torso_pos.y; if (valid) xp[2] = so.torso_pos.z;
        if (valid) xr[0] = so.torso_rot.w; if (valid) xr[1] = so.torso_rot.x; if (valid) xr[2] = so.torso_rot.y; if (valid) xr[3] = so.torso_rot.z;
        if (valid) xv[0] = so.torso_vel.x; if (valid) xv[1] = so.torso_vel.y; if (valid) xv[2] = so.torso_vel.z;
        if (valid) xa[0] = so.torso_ang.x; if (valid) xa[1] = so.torso_ang.y; if (valid) xa[2] = so.torso_ang.z;
      }
      for (int j = 0; j < 3; j++) {
        int b = 1 + 3 * k + j;
        if (valid) xp[3 * b] = dbg.pos[j].x; if (valid) xp[3 * b + 1] = dbg.pos[j].y; if (valid) xp[3 * b + 2] = dbg.pos[j].z;
        if (valid) xr[4 * b] = dbg.rot[j].w; if (valid) xr[4 * b + 1] = dbg.rot[j].x; if (valid) xr[4 * b + 2] = dbg.rot[j].y; if (valid) xr[4 * b + 3] = dbg.rot[j].z;
        if (valid) xv[3 * b] = dbg.vel[j].x; if (valid) xv[3 * b + 1] = dbg.vel[j].y; if (valid) xv[3 * b + 2] = dbg.vel[j].z;
        if (valid) xa[3 * b] = dbg.ang[j].x; if (valid) xa[3 * b + 1] = dbg.ang[j].y; if (valid) xa[3 * b + 2] = dbg.ang[j].z;
      }
    }
    if (p.out.dbg_qfrc_actuator) {
      float *o = p.out.dbg_qfrc_actuator + (size_t)e * 18;
      for (int j = 0; j < 3; j++) if (valid) o[6 + 3 * k + j] = so.frc[j];
      if (k == 0) for (int d = 0; d < 6; d++) if (valid) o[d] = 0.f;
    }
    if (p.out.dbg_qacc) {
      float *o = p.out.dbg_qacc + (size_t)e * 18;
      for (int j = 0; j < 3; j++) if (valid) o[6 + 3 * k + j] = dbg.qacc_l[j];
      if (k == 0) for (int d = 0; d < 6; d++) if (valid) o[d] = dbg.qacc_b[d];
    }
    if (p.out.dbg_solver && !RESET && k == 0) {
      for (int i = 0; i < 8; i++) if (valid) p.out.dbg_solver[(size_t)e * 8 + i] = dbg.solver[i];
    }
    if (p.out.dbg_efc && !RESET) {
      float *o = p.out.dbg_efc + (size_t)e * (PUPPER_NEFC_TAP * 2);
      for (int j = 0; j < 3; j++) {
        const int r = 3 * k + j;
        if (valid) { o[2 * r] = dbg.efc_fD[j]; o[2 * r + 1] = dbg.efc_fA[j]; o[2 * (12 + r)] = dbg.efc_lD[j]; o[2 * (12 + r) + 1] = dbg.efc_lA[j]; }
      }
      for (int c = 0; c < kMaxCon; c++)
        if (valid) { o[2 * (24 + 4 * c + k)] = dbg.efc_cD[c]; o[2 * (24 + 4 * c + k) + 1] = dbg.efc_cA[c]; }
    }
    if (p.out.dbg_site_xpos) {
      float *o = p.out.dbg_site_xpos + (size_t)e * 15;
      if (valid) o[3 + 3 * k] = so.foot_site.x; if (valid) o[4 + 3 * k] = so.foot_site.y; if (valid) o[5 + 3 * k] = so.foot_site.z;
      if (k == 0) {
        V3 s = so.torso_pos + rotate(V3{m.site_pos[0][0], m.site_pos[0][1], m.site_pos[0][2]}, so.torso_rot);
        if (valid) o[0] = s.x; if (valid) o[1] = s.y; if (valid) o[2] = s.z;
      }
    }
    if (p.out.dbg_contact_dist && k == 0) {
      const int mc = m.max_contact_points;
      for (int i = 0; i < mc; i++) {
        bool on = i < es.ncon;
        if (valid) p.out.dbg_contact_dist[(size_t)e * mc + i] = on ? es.con[i].dist : 1.0f;
        int g1 = -1, g2 = -1;
        if (on) {
          const ContactSlot &s = es.con[i];
          g1 = s.s1 >= 0 ? m.sphere_geomid[s.s1] : m.floor_geomid;
          g2 = s.s2 >= 0 ? m.sphere_geomid[s.s2] : -2;  // -2: a world box (id not tracked)
        }
        if (valid) p.out.dbg_contact_geom[((size_t)e * mc + i) * 2] = g1;
        if (valid) p.out.dbg_contact_geom[((size_t)e * mc + i) * 2 + 1] = g2;
      }
    }
  }

  if (RESET) {
    // persistent info written by reset; fused wrapper state
#pragma unroll
    for (int i = 0; i < 3; i++) if (k == i) { if (valid) p.st.command[(size_t)i * stride + e] = oc.command[i]; if (valid) p.st.desired_world_z[(size_t)i * stride + e] = oc.desired_z[i]; }
    if (p.has_ep) {
      const int es_ = p.ep.stride;
      __syncwarp(qm);
      for (int i = k; i < PUPPER_NQ; i += 4) if (valid) p.ep.first_qpos[(size_t)i * es_ + e] = p.st.qpos[(size_t)i * stride + e];
      for (int i = k; i < PUPPER_NV; i += 4) {
        if (valid) p.ep.first_qvel[(size_t)i * es_ + e] = p.st.qvel[(size_t)i * stride + e];
        if (valid) p.ep.first_warmstart[(size_t)i * es_ + e] = p.st.qacc_warmstart[(size_t)i * stride + e];
      }
      for (int i = k; i < H * PUPPER_OBS_DIM; i += 4) if (valid) p.ep.first_obs[(size_t)e * H * PUPPER_OBS_DIM + i] = p.st.obs[(size_t)e * H * PUPPER_OBS_DIM + i];
      for (int i = k; i < PUPPER_NMETRIC; i += 4) if (valid) p.ep.sum_metrics[(size_t)i * es_ + e] = 0.f;
      if (k == 0) { if (valid) p.ep.steps[e] = 0; if (valid) p.ep.truncation[e] = 0.f; if (valid) p.ep.sum_reward[e] = 0.f; if (valid) p.ep.length[e] = 0.f; if (valid) p.ep.episode_done[e] = 0.f; }
    }
    return;
  }

  // ---- S7 foot contacts (own foot) -----------------------------------------------------------------------------
  const float dtenv = c.dt;
  const uint32_t last_contact = p.st.last_contact[e];
  const bool lc = (last_contact >> k) & 1u;
  const float fz = so.foot_site.z - c.foot_radius;
  const bool contact = fz < 1e-3f;
  const bool filt_mm = contact || lc;
  const bool filt_cm = (fz < 3e-2f) || lc;
  float air = p.st.feet_air_time[(size_t)k * stride + e];
  const bool first = (air > 0.f) && filt_mm;
  air += dtenv;

  // ---- S8 termination ------------------------------------------------------------------------------------------
  const V3 up = V3{0.f, 0.f, 1.f};
  const V3 rup = rotate(up, so.torso_rot);
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 3; j++) bad |= (L.ql[j] < c.joint_lower[3 * k + j]) || (L.ql[j] > c.joint_upper[3 * k + j]);
  bool done = ((__ballot_sync(qm, bad) >> qbase) & 0xFu) != 0u;
  done |= rup.z < c.cos_terminal_body_angle;
  done |= so.torso_pos.z < c.terminal_body_z;
  const int step0 = p.st.step[e];

  // ---- S9 rewards (rewards.py) ----------------------------------------------------------------------------------
  float rw[PUPPER_NREWARD];
  {
    const Q4 inv = qinv(so.torso_rot);
    const V3 lv = rotate(so.torso_vel, inv), av = rotate(so.torso_ang, inv), wz = rotate(up, inv);
    const float sigma = c.tracking_sigma;
    float e0 = oc.command[0] - lv.x, e1 = oc.command[1] - lv.y, e2 = oc.command[2] - av.z;
    rw[PUPPER_R_TRACKING_LIN_VEL] = expf(-(e0 * e0 + e1 * e1) / sigma);
    rw[PUPPER_R_TRACKING_ANG_VEL] = expf(-(e2 * e2) / sigma);
    float d0 = wz.x - oc.desired_z[0], d1 = wz.y - oc.desired_z[1], d2 = wz.z - oc.desired_z[2];
    rw[PUPPER_R_TRACKING_ORIENTATION] = expf(-(d0 * d0 + d1 * d1 + d2 * d2) / sigma);
    rw[PUPPER_R_LIN_VEL_Z] = so.torso_vel.z * so.torso_vel.z;
    rw[PUPPER_R_ANG_VEL_XY] = so.torso_ang.x * so.torso_ang.x + so.torso_ang.y * so.torso_ang.y;
    rw[PUPPER_R_ORIENTATION] = rup.x * rup.x + rup.y * rup.y;
    float tq = 0.f, ja = 0.f, mw = 0.f, ar = 0.f, ss = 0.f, sv = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int u = 3 * k + j;
      float lvj = p.st.last_vel[(size_t)u * stride + e];
      tq = fmaf(so.frc[j], so.frc[j], tq);
      float a = (L.vl[j] - lvj) / c.env_dt;
      ja = fmaf(a, a, ja);
      mw += fabsf(so.frc[j] * L.vl[j]);
      float da = act[j] - oc.last_act[j];
      ar = fmaf(da, da, ar);
      ss += fabsf(L.ql[j] - c.default_pose[u]);
      sv += fabsf(L.vl[j]);
    }
    float abd = L.ql[1] - c.desired_abduction[k];
    abd = abd * abd;
    float airt = first ? (air - 0.1f) : 0.f;
    // foot slip: lower-leg xd shifted to the foot site, xy only
    V3 off = so.foot_site - so.lower_pos;
    V3 cr = cross(off, so.lower_ang);
    float vx = so.lower_vel.x - cr.x, vy = so.lower_vel.y - cr.y;
    float slip = filt_cm ? (vx * vx + vy * vy) : 0.f;
    tq = qsum(tq, qm); ja = qsum(ja, qm); mw = qsum(mw, qm); ar = qsum(ar, qm); ss = qsum(ss, qm); sv = qsum(sv, qm);
    abd = qsum(abd, qm); airt = qsum(airt, qm); slip = qsum(slip, qm);
    const float cn = brax_norm(V3{oc.command[0], oc.command[1], oc.command[2]});
    rw[PUPPER_R_TORQUES] = tq;
    rw[PUPPER_R_JOINT_ACCELERATION] = ja;
    rw[PUPPER_R_MECHANICAL_WORK] = mw;
    rw[PUPPER_R_ACTION_RATE] = ar;
    rw[PUPPER_R_STAND_STILL] = ss * (cn < 0.1f ? 1.f : 0.f);
    rw[PUPPER_R_STAND_STILL_JOINT_VELOCITY] = sv * (cn < c.stand_still_command_threshold ? 1.f : 0.f);
    rw[PUPPER_R_ABDUCTION_ANGLE] = abd;
    rw[PUPPER_R_FEET_AIR_TIME] = airt * (cn > 0.05f ? 1.f : 0.f);
    rw[PUPPER_R_FOOT_SLIP] = slip;
    rw[PUPPER_R_TERMINATION] = (done && step0 < c.early_termination_step_threshold) ? 1.f : 0.f;
    rw[PUPPER_R_KNEE_COLLISION] = so.knee_hits;
    rw[PUPPER_R_BODY_COLLISION] = so.torso_hits;
  }
  float total = 0.f;
#pragma unroll
  for (int i = 0; i < PUPPER_NREWARD; i++) { rw[i] = __fmul_rn(rw[i], c.reward_scales[i]); total = __fadd_rn(total, rw[i]); }
  float reward = fminf(fmaxf(total * dtenv, 0.f), 10000.f);
  float fdone = done ? 1.f : 0.f;
  const float total_dist = brax_norm(so.torso_pos);

  // ---- S10 bookkeeping, S11 resampling ------------------------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const int u = 3 * k + j;
    if (valid) p.st.last_act[(size_t)u * stride + e] = act[j];
    if (valid) p.st.last_vel[(size_t)u * stride + e] = L.vl[j];
  }
  if (valid) p.st.feet_air_time[(size_t)k * stride + e] = filt_mm ? 0.f : air;
  const uint32_t cbits = (__ballot_sync(qm, contact) >> qbase) & 0xFu;
  int step = step0 + 1;
  const bool resample = step > c.resample_velocity_step;
  if (__any_sync(qm, resample)) {
    float cmd[3] = {0.f, 0.f, 0.f}, dz[3] = {0.f, 0.f, 0.f};
    if (k == 0 && resample) { sample_command(c, cmd_rng, cmd, ext, ext_s); sample_body_orientation(c, cmd_rng, dz, ext, ext_s); }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float cv = __shfl_sync(qm, cmd[i], qbase), dv = __shfl_sync(qm, dz[i], qbase);
      if (k == i && resample) { if (valid) p.st.command[(size_t)i * stride + e] = cv; if (valid) p.st.desired_world_z[(size_t)i * stride + e] = dv; }
    }
  }
  if (done || resample) step = 0;
  if (k == 0) {
    if (valid) p.st.last_contact[e] = cbits;
    if (valid) p.st.step[e] = step;
    if (valid) p.st.kick[e] = kick0; if (valid) p.st.kick[stride + e] = kick1;
  }

  // ---- fused brax EpisodeWrapper + AutoResetWrapper ---------------------------------------------------
  if (p.has_ep) {
    const int es_ = p.ep.stride;
    const float prev_done = p.ep.episode_done[e];
    const float keep = prev_done != 0.f ? 0.f : 1.f;
    int steps = p.ep.steps[e];
    if (prev_done != 0.f) steps = 0;
    steps += c.action_repeat;
    const bool trunc = steps >= c.episode_length;
    const float truncation = (trunc && !done) ? 1.f : 0.f;
    if (trunc) fdone = 1.f;
    const float sum_reward = (p.ep.sum_reward[e] + reward) * keep;
    const float length = (p.ep.length[e] + (float)c.action_repeat) * keep;
    float ep_metric[5];  // this lane's metric rows i = k (mod 4)
#pragma unroll
    for (int t = 0; t < 5; t++) {  // loads first, so they overlap
      int i = k + 4 * t;
      ep_metric[t] = i < PUPPER_NMETRIC ? p.ep.sum_metrics[(size_t)i * es_ + e] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 5; t++) {
      int i = k + 4 * t;
      if (i < PUPPER_NMETRIC) {
        float mv = total_dist;
#pragma unroll
        for (int q = 0; q < PUPPER_NREWARD; q++) if (q + 1 == i) mv = rw[q];
        float v = (ep_metric[t] + mv) * keep;
        if (valid) p.ep.sum_metrics[(size_t)i * es_ + e] = v;
        ep_metric[t] = v;
      }
    }
    __syncwarp(qm);
    if (k == 0) {
      if (valid) p.ep.steps[e] = steps; if (valid) p.ep.truncation[e] = truncation; if (valid) p.ep.sum_reward[e] = sum_reward; if (valid) p.ep.length[e] = length;
      if (valid) p.ep.episode_done[e] = fdone;
    }
    if (p.ep.totals) {
      if (fdone != 0.f) {  // completed episode: add its sums to the device accumulator (all-reduced by the host side)
        if (k == 0) { if (valid) atomicAdd(p.ep.totals + 0, 1.f); if (valid) atomicAdd(p.ep.totals + 1, sum_reward); if (valid) atomicAdd(p.ep.totals + 2, length); if (valid) atomicAdd(p.ep.totals + 22, done ? 1.f : 0.f); }
#pragma unroll
        for (int t = 0; t < 5; t++) if (k + 4 * t < PUPPER_NMETRIC) if (valid) atomicAdd(p.ep.totals + 3 + k + 4 * t, ep_metric[t]);
      }
    }
    if (fdone != 0.f) {  // AutoResetWrapper: restore the first pipeline state and first obs
      // (loads batched ahead of the stores so they overlap)
      float tq[5], tv[5], tw[5];
#pragma unroll
      for (int t = 0; t < 5; t++) {
        const int i = k + 4 * t;
        tq[t] = i < PUPPER_NQ ? p.ep.first_qpos[(size_t)i * es_ + e] : 0.f;
        tv[t] = i < PUPPER_NV ? p.ep.first_qvel[(size_t)i * es_ + e] : 0.f;
        tw[t] = i < PUPPER_NV ? p.ep.first_warmstart[(size_t)i * es_ + e] : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 5; t++) {
        const int i = k + 4 * t;
        if (valid && i < PUPPER_NQ) p.st.qpos[(size_t)i * stride + e] = tq[t];
        if (valid && i < PUPPER_NV) { p.st.qvel[(size_t)i * stride + e] = tv[t]; p.st.qacc_warmstart[(size_t)i * stride + e] = tw[t]; }
      }
      for (int h = 0; h < H; h++) {
        float to[PUPPER_OBS_DIM / 4];
        const size_t base = (size_t)e * H * PUPPER_OBS_DIM + (size_t)h * PUPPER_OBS_DIM + k;
#pragma unroll
        for (int t = 0; t < PUPPER_OBS_DIM / 4; t++) to[t] = p.ep.first_obs[base + 4 * t];
        if (valid) {
#pragma unroll
          for (int t = 0; t < PUPPER_OBS_DIM / 4; t++) p.st.obs[base + 4 * t] = to[t];
        }
      }
    }
  }

  // ---- S12 outputs ------------------------------------------------------------------------------------------------------
  if (k == 0) { if (valid) at_step(io.reward, p.n_envs)[e] = reward; if (valid) at_step(io.done, p.n_envs)[e] = fdone; if (valid) p.out.metrics[(size_t)e * PUPPER_NMETRIC] = total_dist; }
#pragma unroll
  for (int q = 0; q < PUPPER_NREWARD; q++) if ((q & 3) == k) if (valid) p.out.metrics[(size_t)e * PUPPER_NMETRIC + 1 + q] = rw[q];

  // ---- optional second copy of the observation rows (PupperStepOut.obs_copy; mapped pinned host memory) --------------------
  // The warp's 8 envs are 8 consecutive rows: one contiguous block, copied in 16-byte pieces (512 bytes per warp instruction)
  // after everything that writes obs (get_obs, the auto-reset restore).
  if (p.out.obs_copy) {
    __syncwarp(qm);
    const int w4 = H * (PUPPER_OBS_DIM / 4);  // 16-byte pieces per row
    const int e_w = block * kEnvsPerBlock + (tid >> 5) * 8;
    const int total = min(8, p.n_envs - e_w) * w4;
    const float4 *src = reinterpret_cast<const float4 *>(p.st.obs + (size_t)e_w * H * PUPPER_OBS_DIM);
    float4 *dst = reinterpret_cast<float4 *>(p.out.obs_copy + (size_t)e_w * H * PUPPER_OBS_DIM);
#pragma unroll 1
    for (int i = lane; i < total; i += 32) dst[i] = src[i];
  }
}

template <bool RESET, bool DBG>
__global__ void PUPPER_LB env_kernel(const KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BlockShared &sh = *reinterpret_cast<BlockShared *>(smem_raw);
  env_body<RESET, DBG, false>(p, sh, StepIO{p.action, p.out.reward, p.out.done, 0}, (int)blockIdx.x, (int)threadIdx.x);
}

}  // namespace pupper

// FP32 roofline probe: 8 independent FMA chains per thread (2 flop each), no memory traffic in the loop.
__global__ void __launch_bounds__(256) ffma_probe_kernel(int iters, float *sink) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f, c = 1e-3f;
#pragma unroll 4
  for (int i = 0; i < iters; i++) {
    a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
    a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
  }
  float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 123456.789f) sink[0] = r;  // never true; keeps the loop alive
}

// =====================================================================================================================
// C ABI
// =====================================================================================================================
struct PupperModel {
  int smem_bytes;  // dynamic shared memory per CTA (sizeof(BlockShared) + optional PUPPER_EXTRA_SMEM padding, an occupancy experiment knob)
  int device;
  pupper::ConstBlock *d_consts;  // model description + env configuration + derived constants, as the kernel stages them
  PupperModelDesc h_desc;
  PupperEnvCfg h_cfg;
  int last_launches;
  bool rollout_ready;  // shared-memory attribute of the rollout kernels set on this model's device (first pupper_rollout call)
  bool rollout_chain;  // chained grid (one launch per unroll) or one launch per step
  int *group_step;     // device: per-env-group step counters of the chained launch
  int group_step_len;
};

static thread_local char g_cuda_err[256] = "";

// host twin of kbi() at pos = 0 (friction-loss rows)
static void host_kbi(const float *solref, const float *solimp, float timestep, float &k, float &b, float &imp) {
  float timeconst = fmaxf(solref[0], 2.f * timestep), dampratio = solref[1];
  float dmin = fminf(fmaxf(solimp[0], 1e-4f), 0.9999f), dmax = fminf(fmaxf(solimp[1], 1e-4f), 0.9999f);
  k = 1.f / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  b = 2.f / (dmax * timeconst);
  if (solref[0] <= 0.f) k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0.f) b = -solref[1] / dmax;
  imp = dmin;  // x = 0 < mid: imp = dmin + 0 * (dmax - dmin)
}

static int cuda_fail(cudaError_t e, const char *what) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return PUPPER_ECUDA;
}

extern "C" {

int pupper_abi_version(void) { return PUPPER_ABI_VERSION; }

const char *pupper_strerror(int code) {
  switch (code) {
    case PUPPER_OK: return "ok";
    case PUPPER_EINVAL: return "invalid argument";
    case PUPPER_EUNSUPPORTED: return "model or configuration outside the supported family";
    case PUPPER_ECUDA: return "CUDA runtime error (see pupper_last_cuda_error)";
    case PUPPER_ENOMEM: return "out of memory";
    case PUPPER_EVERSION: return "ABI version mismatch";
    default: return "unknown error";
  }
}

const char *pupper_last_cuda_error(void) { return g_cuda_err; }

int pupper_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(PupperModelDesc);
    case 1: return (int)sizeof(PupperEnvCfg);
    case 2: return (int)sizeof(PupperState);
    case 3: return (int)sizeof(PupperDR);
    case 4: return (int)sizeof(PupperStepOut);
    case 5: return (int)sizeof(PupperEpisode);
    case 6: return (int)sizeof(PupperRand);
    default: return -1;
  }
}

int pupper_model_create(const PupperModelDesc *desc, const PupperEnvCfg *cfg, int device, PupperModel **out) {
  if (!desc || !cfg || !out) return PUPPER_EINVAL;
  if (desc->abi_version != PUPPER_ABI_VERSION || cfg->abi_version != PUPPER_ABI_VERSION) return PUPPER_EVERSION;
  if (desc->iterations != 1 || cfg->threefry_partitionable != 1) return PUPPER_EUNSUPPORTED;
  if (desc->max_geom_pairs < 1 || desc->max_geom_pairs > 4) return PUPPER_EUNSUPPORTED;  // one narrow phase per lane
  if (desc->max_contact_points < 1 || desc->max_contact_points > pupper::kMaxCon) return PUPPER_EUNSUPPORTED;
  if (desc->nbox < 0 || desc->nbox > PUPPER_MAX_BOX) return PUPPER_EUNSUPPORTED;
  if (cfg->observation_history < 1 || cfg->n_frames < 1) return PUPPER_EINVAL;
  if (cfg->n_latency < 1 || cfg->n_latency > PUPPER_MAX_LAT || cfg->n_imu_latency < 1 || cfg->n_imu_latency > PUPPER_MAX_LAT) return PUPPER_EINVAL;
  for (int i = 1; i < PUPPER_NU; i++)  // the DR contract carries one kp / kd per env
    if (desc->act_gain[i] != desc->act_gain[0] || desc->act_bias2[i] != desc->act_bias2[0] || desc->act_bias1[i] != -desc->act_gain[i]) return PUPPER_EUNSUPPORTED;
  for (int b = 1; b < PUPPER_NBODY; b++) {
    int expect = b == 1 ? 0 : ((b - 2) % 3 == 0 ? 1 : b - 1);
    if (desc->body_parent[b] != expect) return PUPPER_EUNSUPPORTED;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  PupperModel *m = new (std::nothrow) PupperModel();
  if (!m) return PUPPER_ENOMEM;
  m->device = device;
  m->h_desc = *desc;
  m->h_cfg = *cfg;
  m->last_launches = 0;
  m->rollout_ready = false;
  m->rollout_chain = true;
  m->group_step = nullptr;
  m->group_step_len = 0;
  // constants that do not depend on the state: friction-loss rows have pos = 0, so their impedance is fixed
  pupper::DerivedConsts dc;
  memset(&dc, 0, sizeof(dc));
  {
    float kk, bb, imp;
    host_kbi(desc->dof_solref, desc->dof_solimp, desc->timestep, kk, bb, imp);
    dc.fric_b = bb;
    for (int d = 0; d < PUPPER_NV; d++) {
      bool on = desc->frictionloss_rows && d >= 6 && desc->dof_frictionloss[d] > 0.f;
      float R = fmaxf(desc->dof_invweight0[d] * (1.f - imp) / imp, 1e-15f);
      dc.fric_loss[d] = on ? desc->dof_frictionloss[d] : 0.f;
      dc.fric_D[d] = 1.f / R;
      dc.fric_rf[d] = dc.fric_loss[d] / dc.fric_D[d];
    }
    {
      int np = 0;
      for (int a = 0; a < PUPPER_NSPHERE; a++)
        for (int b = a + 1; b < PUPPER_NSPHERE; b++)
          if (a / 2 != b / 2) dc.ss_pair[np++] = a | (b << 8);
    }
    for (int b = 0; b < desc->nbox; b++)
      dc.box_rbound[b] = sqrtf(desc->box_size[b][0] * desc->box_size[b][0] + desc->box_size[b][1] * desc->box_size[b][1] + desc->box_size[b][2] * desc->box_size[b][2]);
  }
  m->d_consts = nullptr;
  {
    pupper::ConstBlock *hb = new (std::nothrow) pupper::ConstBlock();
    if (!hb) { delete m; return PUPPER_ENOMEM; }
    memset(hb, 0, sizeof(*hb));
    hb->m = *desc; hb->c = *cfg; hb->d = dc;
    e = cudaMalloc(&m->d_consts, sizeof(pupper::ConstBlock));
    if (e == cudaSuccess) e = cudaMemcpy(m->d_consts, hb, sizeof(pupper::ConstBlock), cudaMemcpyHostToDevice);
    delete hb;
  }
  int smem = (int)sizeof(pupper::BlockShared);
  if (const char *pad = getenv("PUPPER_EXTRA_SMEM")) smem += atoi(pad);
  m->smem_bytes = smem;
  if (e == cudaSuccess) e = cudaFuncSetAttribute(pupper::env_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(pupper::env_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(pupper::env_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(pupper::env_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { cudaFree(m->d_consts); delete m; return cuda_fail(e, "pupper_model_create"); }
  *out = m;
  return PUPPER_OK;
}

int pupper_model_destroy(PupperModel *m) {
  if (!m) return PUPPER_EINVAL;
  cudaSetDevice(m->device);
  cudaFree(m->d_consts);
  if (m->group_step) cudaFree(m->group_step);
  delete m;
  return PUPPER_OK;
}

static int check_common(const PupperModel *model, int n_envs, const PupperDR *dr, const PupperState *st, const PupperStepOut *out,
                        const PupperEpisode *ep) {
  if (!model || !st || !out || n_envs <= 0) return PUPPER_EINVAL;
  if (st->stride < n_envs) return PUPPER_EINVAL;
  if (!st->qpos || !st->qvel || !st->qacc_warmstart || !st->rng || !st->last_act || !st->action_buffer || !st->imu_buffer ||
      !st->last_vel || !st->command || !st->desired_world_z || !st->last_contact || !st->feet_air_time || !st->step || !st->kick || !st->obs)
    return PUPPER_EINVAL;
  if (!out->reward || !out->done || !out->metrics) return PUPPER_EINVAL;
  if ((reinterpret_cast<uintptr_t>(out->obs_copy) & 15) != 0 || (reinterpret_cast<uintptr_t>(st->obs) & 15) != 0) return PUPPER_EINVAL;
  if (dr && (dr->stride < n_envs || !dr->friction || !dr->kp || !dr->kd || !dr->base_ipos || !dr->body_inertia || !dr->body_mass)) return PUPPER_EINVAL;
  if (ep && (ep->stride < n_envs || !ep->first_qpos || !ep->first_qvel || !ep->first_warmstart || !ep->first_obs || !ep->steps ||
             !ep->truncation || !ep->sum_reward || !ep->length || !ep->sum_metrics || !ep->episode_done))
    return PUPPER_EINVAL;
  return PUPPER_OK;
}

static pupper::KParams make_params(const PupperModel *model, int n_envs, const PupperDR *dr, const PupperState *st, const float *action,
                                   const uint32_t *keys, const PupperStepOut *out, const PupperEpisode *ep, const PupperRand *rand) {
  pupper::KParams p;
  memset(&p, 0, sizeof(p));
  p.consts = model->d_consts;
  p.n_envs = n_envs;
  p.st = *st;
  if (dr) { p.dr = *dr; p.has_dr = 1; }
  p.action = action;
  p.keys = keys;
  p.out = *out;
  if (ep) { p.ep = *ep; p.has_ep = 1; }
  if (rand) { p.rand = *rand; p.has_rand = 1; }
  return p;
}

static bool wants_debug(const PupperStepOut *o) {
  return o->dbg_x_pos || o->dbg_qfrc_actuator || o->dbg_contact_dist || o->dbg_site_xpos || o->dbg_qacc || o->dbg_solver || o->dbg_efc;
}
// the x / xd taps come as a group, and so do the two contact taps
static bool debug_ok(const PupperStepOut *o) {
  const bool any_x = o->dbg_x_pos || o->dbg_x_rot || o->dbg_xd_vel || o->dbg_xd_ang;
  const bool all_x = o->dbg_x_pos && o->dbg_x_rot && o->dbg_xd_vel && o->dbg_xd_ang;
  const bool con_ok = (o->dbg_contact_dist != nullptr) == (o->dbg_contact_geom != nullptr);
  return (!any_x || all_x) && con_ok;
}

int pupper_reset(const PupperModel *model, int n_envs, const uint32_t *keys, const PupperDR *dr, PupperState *state, PupperStepOut *out,
                 PupperEpisode *episode, const PupperRand *ext_rand, pupper_stream_t stream) {
  int rc = check_common(model, n_envs, dr, state, out, episode);
  if (rc != PUPPER_OK) return rc;
  if (!keys || !debug_ok(out)) return PUPPER_EINVAL;
  if (ext_rand && (!ext_rand->u || ext_rand->stride < n_envs)) return PUPPER_EINVAL;
  pupper::KParams p = make_params(model, n_envs, dr, state, nullptr, keys, out, episode, ext_rand);
  const int grid = (n_envs + pupper::kEnvsPerBlock - 1) / pupper::kEnvsPerBlock;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (wants_debug(out) || ext_rand) pupper::env_kernel<true, true><<<grid, pupper::kBlock, model->smem_bytes, s>>>(p);
  else pupper::env_kernel<true, false><<<grid, pupper::kBlock, model->smem_bytes, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pupper_reset launch");
  const_cast<PupperModel *>(model)->last_launches = 1;
  return PUPPER_OK;
}

int pupper_step(const PupperModel *model, int n_envs, const PupperDR *dr, PupperState *state, const float *action, const PupperRand *ext_rand,
                PupperStepOut *out, PupperEpisode *episode, pupper_stream_t stream) {
  int rc = check_common(model, n_envs, dr, state, out, episode);
  if (rc != PUPPER_OK) return rc;
  if (!action || !debug_ok(out)) return PUPPER_EINVAL;
  if (ext_rand && (!ext_rand->u || ext_rand->stride < n_envs)) return PUPPER_EINVAL;
  pupper::KParams p = make_params(model, n_envs, dr, state, action, nullptr, out, episode, ext_rand);
  const int grid = (n_envs + pupper::kEnvsPerBlock - 1) / pupper::kEnvsPerBlock;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (wants_debug(out) || ext_rand) pupper::env_kernel<false, true><<<grid, pupper::kBlock, model->smem_bytes, s>>>(p);
  else pupper::env_kernel<false, false><<<grid, pupper::kBlock, model->smem_bytes, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pupper_step launch");
  const_cast<PupperModel *>(model)->last_launches = 1;
  return PUPPER_OK;
}

// Enqueues the FP32 probe: blocks x 256 threads x iters x 8 FMAs (flop = blocks*256*iters*16). Time it with events.
int pupper_probe_ffma(int blocks, int iters, float *device_sink, pupper_stream_t stream) {
  if (blocks <= 0 || iters <= 0 || !device_sink) return PUPPER_EINVAL;
  ffma_probe_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(iters, device_sink);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? PUPPER_OK : cuda_fail(e, "pupper_probe_ffma launch");
}

int pupper_last_launch_count(const PupperModel *model) { return model ? model->last_launches : PUPPER_EINVAL; }

int pupper_state_rows(const PupperEnvCfg *cfg, int32_t *rows_out) {
  if (!cfg || !rows_out) return PUPPER_EINVAL;
  const int32_t rows[14] = {PUPPER_NQ, PUPPER_NV, PUPPER_NV, 2, PUPPER_NU, PUPPER_NU * cfg->n_latency, 6 * cfg->n_imu_latency, PUPPER_NU, 3, 3, 1, 4, 1, 2};
  for (int i = 0; i < 14; i++) rows_out[i] = rows[i];
  return PUPPER_OK;
}

}  // extern "C"

// policy-MLP forward pass (include/pupper_policy.h): same library, same error helpers
#include "pupper_policy.cuh"
// persistent rollout kernel: policy phase + env phase, T steps per launch (include/pupper_policy.h pupper_rollout)
#include "pupper_rollout.cuh"

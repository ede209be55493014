// pupper_rollout.cuh -- ONE launch per unroll of rollout collection (BASELINE configs[4]).
//
// Grid = (CTAs of the batch) x (T steps of the unroll).  CTA (c, t) carries the 32 envs of group c through step t:
//
//   wait           for CTA (c, t - 1): one acquire-load spin on a per-group step counter (CTAs are dispatched in increasing
//                  linear index, so all CTAs of step t - 1 were dispatched before any CTA of step t: the CTA that is waited
//                  for is running or done, never behind the waiter -- the decoupled look-back argument);
//   policy phase   the policy MLP on the group's 32 observation rows (two m16 row blocks), m16n8k8 TF32 MMAs, the four warps
//                  split the output columns so that each weight fragment is read once per CTA -- straight from L2 into
//                  registers (the packed B-fragment tables of pupper_policy.cuh, one coalesced 8-byte load per lane and
//                  tile, 32 loads per lane in flight), activations in a shared-memory strip that aliases the solver scratch;
//   env phase      env_body: the fused env step of pupper_env.cu, unchanged;
//   release        the group's step counter (release store after a device-scope fence).
//
// Envs are independent and the policy reads only its own env's observation, so a group never waits for another group: there
// is no per-step barrier across the batch (a slow group -- a leg-leg contact on the rare solver path -- delays only itself),
// no host round trip, no separate policy / copy launches; the observation and action rows go from one phase to the next
// through L2, and the trajectory (obs, action, reward, done of every step) is written once, by the kernel, in the [T][n][...]
// layout the trainer consumes.  Constants and DR leaves are staged BEFORE the wait, so they overlap the previous step's tail.
//
// Why not a loop over t inside a persistent CTA (the first version of this file): with the policy phase (64 accumulator
// registers, a 64-register fragment ring) and the env step (which has no register to spare) in one loop, ptxas spills the env
// step -- 713 local loads in its body instead of the single-step kernel's 41, whether the policy is inlined or called, with
// every loop-invariant value laundered, the step counter in shared memory, either phase order (the spill count follows the
// policy code's register count: 556 at 199 registers, 180 at 151, 58 for a trivial callee); measured +35 % env-phase time.
// Straight-line, the same two phases compile to the single-step kernel's code (55 local loads).
// Included at the end of pupper_env.cu (after pupper_policy.cuh).
#pragma once

namespace pupper {

constexpr int kRoNT = kPolMaxNT / 4;  // n-tiles per warp (output columns split over the CTA's 4 warps)
static_assert(kBlock == 128, "the rollout kernel splits the policy's output columns over 4 warps");
// the activation strip lives in the solver scratch, which is dead between two env steps
constexpr int kRoScratchFloats = (int)((sizeof(BlockShared) - offsetof(BlockShared, rows)) / sizeof(float));

#ifdef PUPPER_RO_TRACE  // tools/jobs: cycles per phase, summed over CTAs and steps (thread 0 of every CTA)
__device__ unsigned long long g_ro_trace[16];
#define RO_MARK(slot) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_ro_trace[slot], (unsigned long long)(t_ - ro_t0)); ro_t0 = t_; } } while (0)
#define RO_T0_DECL long long ro_t0 = clock64()
#define RO_T0_ARG , long long &ro_t0
#define RO_T0_PASS , ro_t0
#else
#define RO_MARK(slot) ((void)0)
#define RO_T0_DECL ((void)0)
#define RO_T0_ARG
#define RO_T0_PASS
#endif

#ifndef RO_POLICY_INLINE
#define RO_POLICY_INLINE __noinline__
#endif

struct RolloutParams {
  int t0;           // first step of this launch (gridDim.y steps follow)
  int *group_step;  // [CTAs + 1] steps finished per env group (chained launch; NULL: one launch per step), last word: wait time-outs
  float *obs;     // [T][n][in_dim]   obs[t] = what the policy saw at step t
  float *action;  // [T][n][12]
  float *reward;  // [T][n]
  float *done;    // [T][n]
};

// k-steps of one layer for this warp's n-tiles sp, sp + 4, ... and BOTH row blocks; B fragments come from global memory
// (L2 resident, ~1000 cycles away) through a ring of k-steps of registers sized so that 32 fragment loads per lane (8 KB per
// warp) are in flight behind the MMAs whatever the tile count; the A fragments of the next k-step are read from the strip
// while the current k-step's MMAs issue.  NTW > 0: exactly NTW tiles; NTW == 0: up to kRoNT tiles, each guarded by
// nt < nt_n (warp-uniform).
template <int PREC, int NTW>
__device__ __forceinline__ void rollout_ksteps(float (&acc)[kRoNT][2][4], const float *strip, int stride, const float2 *wf, int ksteps,
                                               int nt_n, int sp, int g, int t) {
  constexpr int N = NTW > 0 ? NTW : kRoNT;
  constexpr int D = 32 / N > 16 ? 16 : 32 / N;
  float2 b[D][N];
  auto load = [&](float2 (&dst)[N], int ks) {
    const float2 *wk = wf + (size_t)ks * nt_n * 32;
#pragma unroll
    for (int i = 0; i < N; i++) dst[i] = (NTW > 0 || sp + 4 * i < nt_n) ? __ldg(wk + (sp + 4 * i) * 32) : make_float2(0.f, 0.f);
  };
#pragma unroll
  for (int u = 0; u < D; u++) if (u < ksteps) load(b[u], u);
  const float *r0 = strip + g * stride + t;  // rows g, g + 8 of row block 0; row block 1 is 16 rows further
  float a_cur[2][4], a_nxt[2][4];
  auto load_a = [&](float (&dst)[2][4], int ks) {
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
      const float *a0 = r0 + (mt * 16) * stride + ks * 8, *a1 = a0 + 8 * stride;
      dst[mt][0] = a0[0]; dst[mt][1] = a1[0]; dst[mt][2] = a0[4]; dst[mt][3] = a1[4];
    }
  };
  load_a(a_cur, 0);
#pragma unroll 1
  for (int ks0 = 0; ks0 < ksteps; ks0 += D) {
#pragma unroll
    for (int u = 0; u < D; u++) {
      const int ks = ks0 + u;
      if (ks < ksteps) {
        load_a(a_nxt, ks + 1 < ksteps ? ks + 1 : ks);
        uint32_t a_hi[2][4], a_lo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int i = 0; i < 4; i++) {
            a_hi[mt][i] = PREC == 3 ? tf32_head(a_cur[mt][i]) : __float_as_uint(a_cur[mt][i]);
            a_lo[mt][i] = PREC == 3 ? __float_as_uint(a_cur[mt][i] - __uint_as_float(a_hi[mt][i])) : 0u;
          }
#pragma unroll
        for (int i = 0; i < N; i++) {
          if (NTW > 0 || sp + 4 * i < nt_n) {
            const float2 bb = b[u][i];
            const uint32_t b0 = PREC == 3 ? tf32_head(bb.x) : __float_as_uint(bb.x), b1 = PREC == 3 ? tf32_head(bb.y) : __float_as_uint(bb.y);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
              if (PREC == 3) {
                const uint32_t c0 = __float_as_uint(bb.x - __uint_as_float(b0)), c1 = __float_as_uint(bb.y - __uint_as_float(b1));
                mma_tf32(acc[i][mt], a_lo[mt], b0, b1);
                mma_tf32(acc[i][mt], a_hi[mt], c0, c1);
              }
              mma_tf32(acc[i][mt], a_hi[mt], b0, b1);
            }
          }
        }
        if (ks + D < ksteps) load(b[u], ks + D);
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int i = 0; i < 4; i++) a_cur[mt][i] = a_nxt[mt][i];
      }
    }
  }
}

// Activation in place over the values this lane has just stored (tiles i < ntm, both row blocks, rows g and g + 8): a rolled
// loop, four independent float2 chains per trip.  (Rolled: the policy phase runs once per step on cold code, and straight-line
// transcendentals for 64 accumulators would be fetched at ~6 cycles per instruction.)
template <int ACT>
__device__ __forceinline__ void rollout_apply(float *strip, int stride, int ntm, int sp, int g, int t) {
#pragma unroll 1
  for (int idx0 = 0; idx0 < ntm * 4; idx0 += 4) {
    float2 *ptr[4];
    float2 v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int idx = idx0 + q, i = idx >> 2, row = ((idx >> 1) & 1) * 16 + g + (idx & 1) * 8;
      ptr[q] = reinterpret_cast<float2 *>(strip + row * stride + (sp + 4 * i) * 8 + 2 * t);
      v[q] = *ptr[q];
    }
#pragma unroll
    for (int q = 0; q < 4; q++) { v[q].x = policy_act<ACT>(v[q].x); v[q].y = policy_act<ACT>(v[q].y); }
#pragma unroll
    for (int q = 0; q < 4; q++) *ptr[q] = v[q];
  }
}

// The policy MLP on the CTA's 32 rows.  obs: this CTA's first observation row in global memory (rows `in_dim` apart, written
// by this CTA's env phase or by an earlier launch: plain loads); traj_obs / action: this CTA's first row of the step's
// trajectory slices.  nrows = valid rows of this CTA (the rest compute on zeros and store nothing).  Out of line on purpose:
// its register allocation (64 accumulators + a 64-register fragment ring) stays out of the env step's, whose spills it would
// otherwise add to (measured: +35 % env-phase time when inlined).
template <int PREC>
__device__ RO_POLICY_INLINE void rollout_policy(const PolicyParams &pol, float *strip, const float *obs, float *traj_obs, float *action, int nrows RO_T0_ARG) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sp = warp, g = lane >> 2, t = lane & 3;
  const int stride = pol.stride;
  // ---- observation rows -> strip (zero padded to the layer's k-steps) and -> the trajectory ----------------------------
  // in_dim = 36 H is a multiple of 4 and the CTA's rows are contiguous: 16-byte pieces, all loads of a thread in flight at once
  {
    const int kp0 = pol.layer[0].kp, in_dim = pol.in_dim, q4 = in_dim >> 2, total4 = nrows * q4;
    const float4 *src = reinterpret_cast<const float4 *>(obs);
    float4 *dst = reinterpret_cast<float4 *>(traj_obs);
    constexpr int U = 6;  // 6 x 128 threads x 16 B = 12 KB per pass (H = 2: 9 KB, one pass)
#pragma unroll 1
    for (int base = 0; base < kEnvsPerBlock * q4; base += U * kBlock) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int i = base + u * kBlock + threadIdx.x;
        v[u] = i < total4 ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int i = base + u * kBlock + threadIdx.x;
        if (i < kEnvsPerBlock * q4) {
          const int r = i / q4, c4 = i - r * q4;
          *reinterpret_cast<float4 *>(strip + r * stride + 4 * c4) = v[u];
          if (i < total4) dst[i] = v[u];
        }
      }
    }
    const int pad = kp0 - in_dim;  // 0 or 4
    for (int i = threadIdx.x; i < kEnvsPerBlock * pad; i += kBlock) strip[(i / pad) * stride + in_dim + (i % pad)] = 0.f;
  }
  __syncthreads();
  RO_MARK(0);
  for (int l = 0; l < pol.n_layers; l++) {
    const PolicyLayer &L = pol.layer[l];
    const int ksteps = L.kp >> 3, nt_n = L.np >> 3;
    const int ntm = nt_n > sp ? (nt_n - sp + 3) / 4 : 0;  // n-tiles owned by this warp (warp-uniform)
    float acc[kRoNT][2][4];
#pragma unroll
    for (int i = 0; i < kRoNT; i++)
#pragma unroll
      for (int mt = 0; mt < 2; mt++) { acc[i][mt][0] = acc[i][mt][1] = acc[i][mt][2] = acc[i][mt][3] = 0.f; }
    const float2 *wf = L.wfrag + lane;
    if (ntm == kRoNT) rollout_ksteps<PREC, kRoNT>(acc, strip, stride, wf, ksteps, nt_n, sp, g, t);
    else if (ntm == kRoNT / 2) rollout_ksteps<PREC, kRoNT / 2>(acc, strip, stride, wf, ksteps, nt_n, sp, g, t);
    else if (ntm == 1) rollout_ksteps<PREC, 1>(acc, strip, stride, wf, ksteps, nt_n, sp, g, t);
    else if (ntm > 0) rollout_ksteps<PREC, 0>(acc, strip, stride, wf, ksteps, nt_n, sp, g, t);
    __syncthreads();  // every warp is done reading this layer's inputs: the strip can be overwritten in place
    RO_MARK(1 + l);
#pragma unroll
    for (int i = 0; i < kRoNT; i++) {
      const int nt = sp + 4 * i;
      if (nt < nt_n) {  // warp-uniform
        const int col = nt * 8 + 2 * t;
        const float2 bb = __ldg(reinterpret_cast<const float2 *>(L.bias + col));
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          *reinterpret_cast<float2 *>(strip + (mt * 16 + g) * stride + col) = make_float2(acc[i][mt][0] + bb.x, acc[i][mt][1] + bb.y);
          *reinterpret_cast<float2 *>(strip + (mt * 16 + g + 8) * stride + col) = make_float2(acc[i][mt][2] + bb.x, acc[i][mt][3] + bb.y);
        }
      }
    }
    switch (L.act) {  // one compact rolled loop per activation kind
      case PUPPER_ACT_RELU: rollout_apply<PUPPER_ACT_RELU>(strip, stride, ntm, sp, g, t); break;
      case PUPPER_ACT_SIGMOID: rollout_apply<PUPPER_ACT_SIGMOID>(strip, stride, ntm, sp, g, t); break;
      case PUPPER_ACT_ELU: rollout_apply<PUPPER_ACT_ELU>(strip, stride, ntm, sp, g, t); break;
      case PUPPER_ACT_TANH: rollout_apply<PUPPER_ACT_TANH>(strip, stride, ntm, sp, g, t); break;
      case PUPPER_ACT_SWISH: rollout_apply<PUPPER_ACT_SWISH>(strip, stride, ntm, sp, g, t); break;
      case PUPPER_ACT_GELU: rollout_apply<PUPPER_ACT_GELU>(strip, stride, ntm, sp, g, t); break;
      case PUPPER_ACT_LEAKY_RELU: rollout_apply<PUPPER_ACT_LEAKY_RELU>(strip, stride, ntm, sp, g, t); break;
      default: break;
    }
    __syncthreads();
    RO_MARK(8);
  }
  // ---- last layer's outputs -> action rows (contiguous in global memory for the CTA's rows) -----------------------------
  {
    const int n_out = pol.layer[pol.n_layers - 1].n_out;
#pragma unroll 1
    for (int i = threadIdx.x; i < nrows * n_out; i += kBlock) {
      const int r = i / n_out, c = i - r * n_out;
      action[i] = strip[r * stride + c];
    }
  }
}

__device__ __forceinline__ int ld_acquire_gpu(const int *ptr) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int *ptr, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory"); }

template <int PREC>
__global__ void __launch_bounds__(kBlock, PUPPER_MIN_BLOCKS) rollout_kernel(const KParams p, const __grid_constant__ PolicyParams pol, const RolloutParams ro) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BlockShared &sh = *reinterpret_cast<BlockShared *>(smem_raw);
  RO_T0_DECL;
  // ---- model constants, per-env DR leaves, cleared contact slots (what env_body stages when it runs alone) --------------
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.consts);
    uint4 *dst = reinterpret_cast<uint4 *>(static_cast<ConstBlock *>(&sh));
    constexpr int n16 = (int)(sizeof(ConstBlock) / 16);
#pragma unroll 1
    for (int i = threadIdx.x; i < n16; i += kBlock) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int e0 = blockIdx.x * kEnvsPerBlock;
  {
    const int k = threadIdx.x & 3, el = threadIdx.x >> 2;
    const int e = min(e0 + el, p.n_envs - 1);
    const PupperModelDesc &m = sh.m;
    EnvShared &es = sh.env[el];
    float *ef = reinterpret_cast<float *>(&es);
#pragma unroll 1
    for (int i = k; i < 58; i += 4) {
      float v;
      if (p.has_dr) {
        const int ds = p.dr.stride;
        const float *src = i < 13 ? p.dr.body_mass + (size_t)i * ds
                         : i < 52 ? p.dr.body_inertia + (size_t)(i - 13) * ds
                         : i < 55 ? p.dr.base_ipos + (size_t)(i - 52) * ds
                         : i == 55 ? p.dr.friction : (i == 56 ? p.dr.kp : p.dr.kd);
        v = __ldg(src + e);
      } else {
        v = i < 13 ? m.body_mass[1 + i] : i < 52 ? m.body_inertia[1 + (i - 13) / 3][(i - 13) % 3]
          : i < 55 ? m.body_ipos[1][i - 52] : i == 55 ? -1.f : (i == 56 ? m.act_gain[0] : -m.act_bias2[0]);
      }
      ef[i] = v;
    }
    float *cz = reinterpret_cast<float *>(es.con);
    for (int i = k; i < (int)(sizeof(es.con) / 4); i += 4) cz[i] = 0.f;
    if (k == 0) es.ncon = 0;
  }
  const int t = ro.t0 + (int)blockIdx.y;
  // ---- chained launch: this group's previous step must be finished (its state, obs and episode rows released) ------------
  if (ro.group_step && blockIdx.y > 0) {
    if (threadIdx.x == 0) {
      int spins = 0;
      while (ld_acquire_gpu(ro.group_step + blockIdx.x) < (int)blockIdx.y) {
        __nanosleep(100);
        if (++spins > (1 << 18)) { atomicAdd(ro.group_step + gridDim.x, 1); break; }  // ~0.3 s; never in practice: do not hang the device (pupper_rollout_timeouts counts it)
      }
    }
  }
  __syncthreads();
  RO_MARK(11);
  {
    const int in_dim = pol.in_dim, n_act = pol.layer[pol.n_layers - 1].n_out;
    const size_t row = (size_t)t * p.n_envs + e0;
    rollout_policy<PREC>(pol, sh.rows, p.st.obs + (size_t)e0 * in_dim, ro.obs + row * in_dim, ro.action + row * n_act,
                         min(kEnvsPerBlock, p.n_envs - e0) RO_T0_PASS);
  }
  __syncthreads();  // the actions are visible to the CTA; the strip is dead from here (the solver scratch takes it back)
  RO_MARK(9);
  env_body<false, false, true>(p, sh, StepIO{ro.action, ro.reward, ro.done, t}, (int)blockIdx.x, (int)threadIdx.x);
  const bool last = blockIdx.y + 1 == gridDim.y;
  if (last) {  // the single-step outputs of the runtime keep meaning "the last step" (each lane copies what it has just written)
    const int e = e0 + ((int)threadIdx.x >> 2);
    if ((threadIdx.x & 3) == 0 && e < p.n_envs) {
      p.out.reward[e] = ro.reward[(size_t)t * p.n_envs + e];
      p.out.done[e] = ro.done[(size_t)t * p.n_envs + e];
    }
  }
  RO_MARK(10);
  if (ro.group_step) {
    __threadfence();
    __syncthreads();
    // the last step of the launch leaves the counter at 0 for the next launch (stream order makes that launch start after this one)
    if (threadIdx.x == 0) st_release_gpu(ro.group_step + blockIdx.x, last ? 0 : (int)blockIdx.y + 1);
  }
}

}  // namespace pupper

extern "C" {

#ifdef PUPPER_RO_TRACE
int pupper_rollout_trace(unsigned long long *host16, int clear) {
  cudaError_t e = cudaMemcpyFromSymbol(host16, pupper::g_ro_trace, sizeof(pupper::g_ro_trace));
  if (e == cudaSuccess && clear) { unsigned long long z[16] = {0}; e = cudaMemcpyToSymbol(pupper::g_ro_trace, z, sizeof(z)); }
  return e == cudaSuccess ? 0 : -3;
}
#endif

int pupper_rollout(const PupperModel *model, const PupperPolicy *policy, int n_envs, int unroll_length, const PupperDR *dr, PupperState *state,
                   PupperStepOut *out, PupperEpisode *episode, float *traj_obs, float *traj_action, float *traj_reward, float *traj_done,
                   pupper_stream_t stream) {
  int rc = check_common(model, n_envs, dr, state, out, episode);
  if (rc != PUPPER_OK) return rc;
  if (!policy || unroll_length < 1 || unroll_length > 65535 || !traj_obs || !traj_action || !traj_reward || !traj_done) return PUPPER_EINVAL;
  if (wants_debug(out)) return PUPPER_EINVAL;  // the debug taps belong to single-step calls
  const pupper::PolicyParams &P = policy->params;
  if (policy->device != model->device) return PUPPER_EINVAL;
  if (P.in_dim != model->h_cfg.observation_history * PUPPER_OBS_DIM || P.layer[P.n_layers - 1].n_out != PUPPER_NU) return PUPPER_EINVAL;
  if (pupper::kEnvsPerBlock * P.stride > pupper::kRoScratchFloats) return PUPPER_EUNSUPPORTED;  // activation strip vs solver scratch
  for (int l = 0; l < P.n_layers; l++)
    if (P.layer[l].np > 8 * 4 * pupper::kRoNT) return PUPPER_EUNSUPPORTED;
  PupperModel *mm = const_cast<PupperModel *>(model);
  const int grid = (n_envs + pupper::kEnvsPerBlock - 1) / pupper::kEnvsPerBlock;
  cudaError_t e = cudaSetDevice(model->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  if (!mm->rollout_ready) {
    e = cudaFuncSetAttribute(pupper::rollout_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, model->smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pupper::rollout_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, model->smem_bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(rollout_kernel)");
    mm->rollout_ready = true;
    mm->rollout_chain = getenv("PUPPER_ROLLOUT_PER_STEP") == nullptr;  // diagnostic switch: one launch per step instead of the chained grid
  }
  if (mm->rollout_chain && mm->group_step_len < grid + 1) {  // per-group step counters (+ the time-out word), zero between launches
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(static_cast<cudaStream_t>(stream), &cs);
    if (cs != cudaStreamCaptureStatusNone) return PUPPER_EINVAL;  // first call for this batch size allocates: make it before capturing
    if (mm->group_step) cudaFree(mm->group_step);
    mm->group_step = nullptr; mm->group_step_len = 0;
    e = cudaMalloc(&mm->group_step, (size_t)(grid + 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(mm->group_step, 0, (size_t)(grid + 1) * sizeof(int));
    if (e != cudaSuccess) return cuda_fail(e, "pupper_rollout step counters");
    mm->group_step_len = grid + 1;
  }
  pupper::KParams p = make_params(model, n_envs, dr, state, nullptr, nullptr, out, episode, nullptr);
  pupper::PolicyParams pp = P;
  pp.n = n_envs; pp.obs = nullptr; pp.action = nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool tf32 = policy->precision == PUPPER_POLICY_TF32;
  const int launches = mm->rollout_chain ? 1 : unroll_length;
  for (int l = 0; l < launches; l++) {
    pupper::RolloutParams ro{l, mm->rollout_chain ? mm->group_step : nullptr, traj_obs, traj_action, traj_reward, traj_done};
    const dim3 g(grid, mm->rollout_chain ? unroll_length : 1);
    if (tf32) pupper::rollout_kernel<1><<<g, pupper::kBlock, model->smem_bytes, s>>>(p, pp, ro);
    else pupper::rollout_kernel<3><<<g, pupper::kBlock, model->smem_bytes, s>>>(p, pp, ro);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pupper_rollout launch");
  mm->last_launches = launches;
  return PUPPER_OK;
}

/* Diagnostics of the chained launch: how many waits on a previous step timed out since the counters were allocated (0 unless the
 * device's CTA dispatch order ever broke the chain; synchronises the device). */
int pupper_rollout_timeouts(const PupperModel *model) {
  if (!model) return PUPPER_EINVAL;
  if (!model->group_step) return 0;
  int v = 0;
  cudaError_t e = cudaSetDevice(model->device);
  if (e == cudaSuccess) e = cudaMemcpy(&v, model->group_step + model->group_step_len - 1, sizeof(int), cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? v : cuda_fail(e, "pupper_rollout_timeouts");
}

}  // extern "C"

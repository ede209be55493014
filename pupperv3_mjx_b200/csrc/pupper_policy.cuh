// pupper_policy.cuh -- fused forward pass of the policy MLP (include/pupper_policy.h), one launch per call.
//
// A group of NSPLIT warps owns 16 rows (envs) of the batch and carries them through ALL layers: the activations of the
// rows live in a shared-memory strip that only this group touches, warp `sp` of the group accumulates the output
// n-tiles sp, sp + NSPLIT, ... of the layer in registers, and writes them back in place over the inputs once the layer
// is finished (the CTA barriers of the weight-chunk pipeline order those writes against the other warps' reads).
// Splitting the columns over warps is what fills the machine at rollout batch sizes: 8192 rows are only 512 row
// blocks, i.e. 3.5 warps per SM when one warp carries a whole row block (NSPLIT = 1; measured 81 us per call), and
// every phase of the kernel is latency bound at that occupancy.  The weights stream through
// two 32 KB shared-memory buffers in chunks of k-steps (cp.async, the next chunk in flight while the current one is
// multiplied), shared by the CTA's four warps.  The products run on the tensor cores as m16n8k8 TF32 MMAs with float32
// accumulation; with PREC = 3 every operand is split into a TF32 head and a tail and the three significant partial
// products are accumulated (small terms first), which restores float32-level accuracy.  Weights are re-packed on the
// host into MMA B-fragment order, so a chunk is one contiguous block in global memory (L2 resident: ~340 KB for the
// reference's 72-256-128-128-128-12 policy) and a lane reads its fragment of one (k-step, n-tile) with one 8-byte load.
// Included at the end of pupper_env.cu (one translation unit, shared error helpers).
#pragma once
#include <vector>

#include "../../include/pupper_policy.h"

namespace pupper {

constexpr int kPolRowsPerWarp = 16;
constexpr int kPolWarps = 4;
constexpr int kPolRows = kPolRowsPerWarp * kPolWarps;  // rows per CTA (kPolWarps row blocks, NSPLIT warps each)
constexpr int kPolMaxNT = PUPPER_POLICY_MAX_OUT / 8;   // n-tiles of the widest layer
#ifndef PUPPER_POLICY_NSPLIT
#define PUPPER_POLICY_NSPLIT 4
#endif

struct PolicyLayer {
  const float2 *wfrag;  // [K/8][N/8][32 lanes] B fragments (b0, b1) of row-major W[in, out], zero padded
  const float *bias;    // [N padded to 8]
  int kp, np;           // padded sizes (multiples of 8)
  int n_out;            // true output width
  int act;
};
struct PolicyParams {
  PolicyLayer layer[PUPPER_POLICY_MAX_LAYERS];
  int n_layers;
  int in_dim;   // true input width
  int stride;   // floats per activation row in shared memory (== 4 mod 32: conflict-free A-fragment loads)
  int n;
  const float *obs;
  float *action;
  float *obs_record;  // optional [n][in_dim]: copy of the observation rows read (the rollout's trajectory slice)
};

// TF32 head of a float by truncation (the tensor core ignores the low 13 mantissa bits of its operands anyway, so the
// single-product mode feeds raw float bits).  cvt.rna.tf32.f32 is not used: on sm_100a it expands to a ~5-instruction
// sequence per value, which dominated this kernel's instruction count.  With the head truncated, x - head is exact.
__device__ __forceinline__ uint32_t tf32_head(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int ACT>
__device__ __forceinline__ float policy_act(float x) {
  if (ACT == PUPPER_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == PUPPER_ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-x));  // ex2.approx + rcp.approx (no IEEE fix-up path)
  if (ACT == PUPPER_ACT_ELU) return x > 0.f ? x : expm1f(x);
  if (ACT == PUPPER_ACT_TANH) return tanhf(x);
  if (ACT == PUPPER_ACT_SWISH) return __fdividef(x, 1.f + __expf(-x));  // ex2.approx + rcp.approx: ~1e-6 relative, far inside float32 policy noise
  if (ACT == PUPPER_ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  if (ACT == PUPPER_ACT_LEAKY_RELU) return x > 0.f ? x : 0.01f * x;
  return x;
}

// Activation over this warp's n-tiles (sp, sp + NSPLIT, ...) of the 16-row strip, in place, or to global memory for the
// last layer.  A rolled loop on purpose (see the epilogue comment in the kernel).
template <int ACT, int NSPLIT>
__device__ __forceinline__ void policy_apply(float *rows, int stride, int np, int n_out, bool last, int row0, int n, float *action, int lane, int sp) {
  const int nt_n = np >> 3;
  const int ntm = nt_n > sp ? (nt_n - sp + NSPLIT - 1) / NSPLIT : 0;  // tiles owned by this warp
  const int wcols = ntm * 8, total = kPolRowsPerWarp * wcols;
#pragma unroll 4
  for (int i = lane; i < total; i += 32) {
    const int r = i / wcols, cc = i - r * wcols;
    const int c = (sp + NSPLIT * (cc >> 3)) * 8 + (cc & 7);
    float *rr = rows + r * stride;
    const float v = policy_act<ACT>(rr[c]);
    const int row = row0 + r;
    if (!last) rr[c] = v;
    else if (row < n && c < n_out) action[(size_t)row * n_out + c] = v;
  }
}

constexpr int kPolChunkFloats = 8192;  // one staged weight chunk: 32 KB = (k-steps) x (n-tiles) x 64 floats

// Stage chunk `c` of layer `l` (its k-steps [c*kc, ...)) into `dst` with 16-byte async copies by the whole CTA.
template <int THREADS>
__device__ __forceinline__ void policy_issue_chunk(const PolicyParams &p, int l, int c, float *dst) {
  const PolicyLayer &L = p.layer[l];
  const int nt_n = L.np >> 3, ksteps = L.kp >> 3;
  const int kc = max(1, kPolChunkFloats / (nt_n * 64));
  const int ks0 = c * kc, nks = min(kc, ksteps - ks0);
  const float4 *src = reinterpret_cast<const float4 *>(L.wfrag) + (size_t)ks0 * nt_n * 16;
  const int units = nks * nt_n * 16;
  const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
  for (int i = threadIdx.x; i < units; i += THREADS)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u * i), "l"(src + i));
}

// Bias + activation on the accumulator registers of this warp's n-tiles, stored once: to the strip (input of the next
// layer) or, for the last layer, to global memory.  Lane (g, t) holds rows g, g + 8 and columns 8 nt + 2 t, + 1.
template <int ACT, int NSPLIT, int NT>
__device__ __forceinline__ void policy_store(const float (&acc)[NT][4], float *rows, int stride, const PolicyLayer &L, int nt_n, bool last,
                                             int row0, int n, float *action, int g, int t, int sp) {
#pragma unroll
  for (int i = 0; i < NT; i++) {
    const int nt = sp + NSPLIT * i;
    if (nt < nt_n) {  // warp-uniform
      const int col = nt * 8 + 2 * t;
      const float2 bb = __ldg(reinterpret_cast<const float2 *>(L.bias + col));
      const float v0 = policy_act<ACT>(acc[i][0] + bb.x), v1 = policy_act<ACT>(acc[i][1] + bb.y);
      const float v2 = policy_act<ACT>(acc[i][2] + bb.x), v3 = policy_act<ACT>(acc[i][3] + bb.y);
      if (!last) {
        *reinterpret_cast<float2 *>(rows + g * stride + col) = make_float2(v0, v1);
        *reinterpret_cast<float2 *>(rows + (g + 8) * stride + col) = make_float2(v2, v3);
      } else {
        const int r0 = row0 + g, r1 = r0 + 8;
        if (r0 < n && col < L.n_out) action[(size_t)r0 * L.n_out + col] = v0;
        if (r0 < n && col + 1 < L.n_out) action[(size_t)r0 * L.n_out + col + 1] = v1;
        if (r1 < n && col < L.n_out) action[(size_t)r1 * L.n_out + col] = v2;
        if (r1 < n && col + 1 < L.n_out) action[(size_t)r1 * L.n_out + col + 1] = v3;
      }
    }
  }
}

// The k-steps [ks0, ks0 + nks) of one staged weight chunk for this warp's n-tiles sp, sp + NSPLIT, ...:
// NTW > 0: exactly NTW tiles, no guards; NTW == 0: up to NT tiles, each guarded by nt < nt_n.
template <int PREC, int NSPLIT, int NT, int NTW>
__device__ __forceinline__ void policy_ksteps(float (&acc)[NT][4], const float *ra0, const float *ra1, const float2 *wf, int ks0, int nks,
                                              int nt_n, int sp) {
  constexpr int N = NTW > 0 ? NTW : NT;
#pragma unroll 1
  for (int ks = 0; ks < nks; ks++) {
    // A fragment of the group's 16 rows x 8 input columns
    const float *a0 = ra0 + (ks0 + ks) * 8, *a1 = ra1 + (ks0 + ks) * 8;
    const float a_f[4] = {a0[0], a1[0], a0[4], a1[4]};
    uint32_t a_hi[4], a_lo[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      a_hi[i] = PREC == 3 ? tf32_head(a_f[i]) : __float_as_uint(a_f[i]);
      a_lo[i] = PREC == 3 ? __float_as_uint(a_f[i] - __uint_as_float(a_hi[i])) : 0u;
    }
    const float2 *wk = wf + ks * nt_n * 32;
    float2 b[N];
#pragma unroll
    for (int i = 0; i < N; i++) b[i] = (NTW > 0 || sp + NSPLIT * i < nt_n) ? wk[(sp + NSPLIT * i) * 32] : make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < N; i++) {
      if (NTW > 0 || sp + NSPLIT * i < nt_n) {  // warp-uniform
        const uint32_t b0 = PREC == 3 ? tf32_head(b[i].x) : __float_as_uint(b[i].x), b1 = PREC == 3 ? tf32_head(b[i].y) : __float_as_uint(b[i].y);
        if (PREC == 3) {
          const uint32_t c0 = __float_as_uint(b[i].x - __uint_as_float(b0)), c1 = __float_as_uint(b[i].y - __uint_as_float(b1));
          mma_tf32(acc[i], a_lo, b0, b1);
          mma_tf32(acc[i], a_hi, c0, c1);
        }
        mma_tf32(acc[i], a_hi, b0, b1);
      }
    }
  }
}

template <int PREC, int NSPLIT>
__global__ void __launch_bounds__(32 * kPolWarps * NSPLIT) policy_kernel(const PolicyParams p) {
  constexpr int THREADS = 32 * kPolWarps * NSPLIT;
  constexpr int NT = (kPolMaxNT + NSPLIT - 1) / NSPLIT;  // n-tiles per warp, held in registers
  extern __shared__ __align__(16) float pol_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = warp / NSPLIT, sp = warp - rb * NSPLIT;  // row block of this warp, column split within it
  const int g = lane >> 2, t = lane & 3;  // MMA fragment coordinates: group id, thread in group
  float *rows = pol_smem + (size_t)rb * kPolRowsPerWarp * p.stride;  // the group's 16 activation rows
  float *wbase = pol_smem + (size_t)kPolRows * p.stride;  // two weight chunk buffers
  const int row0 = blockIdx.x * kPolRows + rb * kPolRowsPerWarp;
  const float *ra0 = rows + g * p.stride + t, *ra1 = rows + (g + 8) * p.stride + t;  // A-fragment rows of this lane

  // first weight chunk on its way while the input rows are loaded
  policy_issue_chunk<THREADS>(p, 0, 0, wbase);
  asm volatile("cp.async.commit_group;");
  // ---- layer-0 input: obs rows -> shared memory (zero padded to kp, zero rows past the batch) -----------------------
  // All loads of a pass are issued before the first store, so they overlap (16 rows x kp values per warp).
  {
    const int kp0 = p.layer[0].kp;
    const int total = kPolRowsPerWarp * kp0;
    const int gl = sp * 32 + lane;  // thread index within the group
#pragma unroll 1
    for (int base = 0; base < total; base += 32 * NSPLIT * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int i = base + u * 32 * NSPLIT + gl, r = i / kp0, c = i - r * kp0;
        const int row = row0 + r;
        v[u] = (i < total && row < p.n && c < p.in_dim) ? __ldg(p.obs + (size_t)row * p.in_dim + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int i = base + u * 32 * NSPLIT + gl, r = i / kp0, c = i - r * kp0;
        if (i < total) rows[r * p.stride + c] = v[u];
        if (p.obs_record && i < total && row0 + r < p.n && c < p.in_dim) p.obs_record[(size_t)(row0 + r) * p.in_dim + c] = v[u];
      }
    }
  }
  // (the first chunk's CTA barrier below orders these stores against the other warps' fragment loads)

  int buf = 0;
  for (int l = 0; l < p.n_layers; l++) {
    const PolicyLayer &L = p.layer[l];
    const int ksteps = L.kp >> 3, nt_n = L.np >> 3;
    const int kc = max(1, kPolChunkFloats / (nt_n * 64));
    const int nchunk = (ksteps + kc - 1) / kc;
    const int ntm = nt_n > sp ? (nt_n - sp + NSPLIT - 1) / NSPLIT : 0;  // n-tiles owned by this warp (warp-uniform)
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; i++) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
#pragma unroll 1
    for (int c = 0; c < nchunk; c++) {
      // prefetch the next chunk (of this layer or of the next one) into the other buffer, then wait for the current one
      const bool more_here = c + 1 < nchunk, more = more_here || l + 1 < p.n_layers;
      if (more) policy_issue_chunk<THREADS>(p, more_here ? l : l + 1, more_here ? c + 1 : 0, wbase + (buf ^ 1) * kPolChunkFloats);
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 1;");
      __syncthreads();
      const float2 *wf = reinterpret_cast<const float2 *>(wbase + buf * kPolChunkFloats) + lane;
      const int ks0 = c * kc, nks = min(kc, ksteps - ks0);
      // k-steps of this chunk with the warp's tile count as a compile-time constant: a predicated-off HMMA still
      // occupies the tensor pipe (ncu: half of the pipe's busy cycles went to them), so narrow layers get their own body
      if (ntm == NT) policy_ksteps<PREC, NSPLIT, NT, NT>(acc, ra0, ra1, wf, ks0, nks, nt_n, sp);
      else if (NT >= 2 && ntm == NT / 2) policy_ksteps<PREC, NSPLIT, NT, (NT >= 2 ? NT / 2 : 1)>(acc, ra0, ra1, wf, ks0, nks, nt_n, sp);
      else if (NT >= 4 && ntm == NT / 4) policy_ksteps<PREC, NSPLIT, NT, (NT >= 4 ? NT / 4 : 1)>(acc, ra0, ra1, wf, ks0, nks, nt_n, sp);
      else if (ntm == 1) policy_ksteps<PREC, NSPLIT, NT, 1>(acc, ra0, ra1, wf, ks0, nks, nt_n, sp);
      else if (ntm > 0) policy_ksteps<PREC, NSPLIT, NT, 0>(acc, ra0, ra1, wf, ks0, nks, nt_n, sp);  // other counts: guarded tiles
      __syncthreads();  // the buffer just consumed is the target of the next iteration's prefetch
      buf ^= 1;
    }
    // Epilogue.  (The last chunk's trailing CTA barrier means every warp is done reading this layer's inputs, so the strip
    // can be overwritten; each warp writes only the columns of its own n-tiles, the next layer's first chunk barrier
    // publishes them.)  With few tiles per warp (NT <= 8) bias + activation are applied to the accumulator registers and
    // the result is stored once; with many, the accumulators go to the strip first and a ROLLED loop applies the
    // activation there (unrolling a transcendental per accumulator register would cost ~100 KB of straight-line code).
    const bool last = l == p.n_layers - 1;
    if (NT <= 8) {
      switch (L.act) {
        case PUPPER_ACT_RELU: policy_store<PUPPER_ACT_RELU, NSPLIT, NT>(acc, rows, p.stride, L, nt_n, last, row0, p.n, p.action, g, t, sp); break;
        case PUPPER_ACT_SIGMOID: policy_store<PUPPER_ACT_SIGMOID, NSPLIT, NT>(acc, rows, p.stride, L, nt_n, last, row0, p.n, p.action, g, t, sp); break;
        case PUPPER_ACT_ELU: policy_store<PUPPER_ACT_ELU, NSPLIT, NT>(acc, rows, p.stride, L, nt_n, last, row0, p.n, p.action, g, t, sp); break;
        case PUPPER_ACT_TANH: policy_store<PUPPER_ACT_TANH, NSPLIT, NT>(acc, rows, p.stride, L, nt_n, last, row0, p.n, p.action, g, t, sp); break;
        case PUPPER_ACT_SWISH: policy_store<PUPPER_ACT_SWISH, NSPLIT, NT>(acc, rows, p.stride, L, nt_n, last, row0, p.n, p.action, g, t, sp); break;
        case PUPPER_ACT_GELU: policy_store<PUPPER_ACT_GELU, NSPLIT, NT>(acc, rows, p.stride, L, nt_n, last, row0, p.n, p.action, g, t, sp); break;
        case PUPPER_ACT_LEAKY_RELU: policy_store<PUPPER_ACT_LEAKY_RELU, NSPLIT, NT>(acc, rows, p.stride, L, nt_n, last, row0, p.n, p.action, g, t, sp); break;
        default: policy_store<PUPPER_ACT_LINEAR, NSPLIT, NT>(acc, rows, p.stride, L, nt_n, last, row0, p.n, p.action, g, t, sp); break;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NT; i++) {
        const int nt = sp + NSPLIT * i;
        if (nt < nt_n) {
          const int col = nt * 8 + 2 * t;
          const float2 bb = __ldg(reinterpret_cast<const float2 *>(L.bias + col));
          *reinterpret_cast<float2 *>(rows + g * p.stride + col) = make_float2(acc[i][0] + bb.x, acc[i][1] + bb.y);
          *reinterpret_cast<float2 *>(rows + (g + 8) * p.stride + col) = make_float2(acc[i][2] + bb.x, acc[i][3] + bb.y);
        }
      }
      __syncwarp();
      // columns past n_out hold act(0): the next layer's padded weight rows are zero, any finite value works
      switch (L.act) {  // one compact rolled loop per activation kind
        case PUPPER_ACT_RELU: policy_apply<PUPPER_ACT_RELU, NSPLIT>(rows, p.stride, L.np, L.n_out, last, row0, p.n, p.action, lane, sp); break;
        case PUPPER_ACT_SIGMOID: policy_apply<PUPPER_ACT_SIGMOID, NSPLIT>(rows, p.stride, L.np, L.n_out, last, row0, p.n, p.action, lane, sp); break;
        case PUPPER_ACT_ELU: policy_apply<PUPPER_ACT_ELU, NSPLIT>(rows, p.stride, L.np, L.n_out, last, row0, p.n, p.action, lane, sp); break;
        case PUPPER_ACT_TANH: policy_apply<PUPPER_ACT_TANH, NSPLIT>(rows, p.stride, L.np, L.n_out, last, row0, p.n, p.action, lane, sp); break;
        case PUPPER_ACT_SWISH: policy_apply<PUPPER_ACT_SWISH, NSPLIT>(rows, p.stride, L.np, L.n_out, last, row0, p.n, p.action, lane, sp); break;
        case PUPPER_ACT_GELU: policy_apply<PUPPER_ACT_GELU, NSPLIT>(rows, p.stride, L.np, L.n_out, last, row0, p.n, p.action, lane, sp); break;
        case PUPPER_ACT_LEAKY_RELU: policy_apply<PUPPER_ACT_LEAKY_RELU, NSPLIT>(rows, p.stride, L.np, L.n_out, last, row0, p.n, p.action, lane, sp); break;
        default: policy_apply<PUPPER_ACT_LINEAR, NSPLIT>(rows, p.stride, L.np, L.n_out, last, row0, p.n, p.action, lane, sp); break;
      }
    }
    __syncwarp();
  }
}

}  // namespace pupper

#include "pupper_policy_tc.cuh"

struct PupperPolicy {
  pupper::PolicyParams params;  // device pointers filled in, n / obs / action set per call
  int device, precision, smem_bytes;
  std::vector<void *> allocs;
  bool use_tc = false;          // TF32 mode on the tcgen05 kernel (every layer width <= 256)
  bool tc_v2 = false;           // PUPPER_POLICY_TC2 at create time: the warp-specialised tcgen05 kernel (measured equal at 8192 rows, +3 % at 65536)
  pupper::TcParams tc;
};

extern "C" {

int pupper_policy_destroy(PupperPolicy *policy) {
  if (!policy) return PUPPER_EINVAL;
  for (void *a : policy->allocs) cudaFree(a);
  delete policy;
  return PUPPER_OK;
}

int pupper_policy_create(int n_layers, const int32_t *in_dims, const int32_t *out_dims, const int32_t *activations,
                         const float *const *weights, const float *const *biases, int device, int precision,
                         PupperPolicy **out) {
  if (!in_dims || !out_dims || !activations || !weights || !biases || !out) return PUPPER_EINVAL;
  if (n_layers < 1 || n_layers > PUPPER_POLICY_MAX_LAYERS) return PUPPER_EUNSUPPORTED;
  if (precision != PUPPER_POLICY_TF32 && precision != PUPPER_POLICY_3XTF32) return PUPPER_EINVAL;
  int wmax = 0;
  for (int l = 0; l < n_layers; l++) {
    if (!weights[l] || !biases[l] || in_dims[l] < 1 || out_dims[l] < 1) return PUPPER_EINVAL;
    if (in_dims[l] > PUPPER_POLICY_MAX_WIDTH || out_dims[l] > PUPPER_POLICY_MAX_OUT) return PUPPER_EUNSUPPORTED;
    if (l > 0 && in_dims[l] != out_dims[l - 1]) return PUPPER_EINVAL;
    if (activations[l] < PUPPER_ACT_LINEAR || activations[l] > PUPPER_ACT_LEAKY_RELU) return PUPPER_EUNSUPPORTED;
    wmax = std::max(wmax, std::max((in_dims[l] + 7) / 8 * 8, (out_dims[l] + 7) / 8 * 8));
  }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  PupperPolicy *pol = new (std::nothrow) PupperPolicy();
  if (!pol) return PUPPER_ENOMEM;
  pol->device = device;
  pol->precision = precision;
  pupper::PolicyParams &P = pol->params;
  memset(&P, 0, sizeof(P));
  P.n_layers = n_layers;
  P.in_dim = in_dims[0];
  P.stride = (wmax + 31) / 32 * 32 + 4;
  pol->smem_bytes = (pupper::kPolRows * P.stride + 2 * pupper::kPolChunkFloats) * (int)sizeof(float);
  for (int l = 0; l < n_layers; l++) {
    const int K = in_dims[l], N = out_dims[l], kp = (K + 7) / 8 * 8, np = (N + 7) / 8 * 8;
    std::vector<float> frag((size_t)kp * np), bias(np, 0.f);
    const float *W = weights[l];
    for (int ks = 0; ks < kp / 8; ks++)
      for (int nt = 0; nt < np / 8; nt++)
        for (int lane = 0; lane < 32; lane++) {
          const int k0 = ks * 8 + (lane & 3), k1 = k0 + 4, n = nt * 8 + (lane >> 2);
          const size_t o = (((size_t)ks * (np / 8) + nt) * 32 + lane) * 2;
          frag[o] = (k0 < K && n < N) ? W[(size_t)k0 * N + n] : 0.f;
          frag[o + 1] = (k1 < K && n < N) ? W[(size_t)k1 * N + n] : 0.f;
        }
    for (int n = 0; n < N; n++) bias[n] = biases[l][n];
    void *dw = nullptr, *db = nullptr;
    e = cudaMalloc(&dw, frag.size() * sizeof(float));
    if (e == cudaSuccess) { pol->allocs.push_back(dw); e = cudaMalloc(&db, bias.size() * sizeof(float)); }
    if (e == cudaSuccess) { pol->allocs.push_back(db); e = cudaMemcpy(dw, frag.data(), frag.size() * sizeof(float), cudaMemcpyHostToDevice); }
    if (e == cudaSuccess) e = cudaMemcpy(db, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { pupper_policy_destroy(pol); return cuda_fail(e, "policy weight upload"); }
    P.layer[l] = pupper::PolicyLayer{reinterpret_cast<const float2 *>(dw), reinterpret_cast<const float *>(db), kp, np, N, activations[l]};
  }
  // ---- tcgen05 path (TF32 mode, widths <= 256): weights packed per K-chunk as W^T in the K-major canonical UMMA layout ----
  {
    bool fits = precision == PUPPER_POLICY_TF32 && !getenv("PUPPER_POLICY_LEGACY");
    for (int l = 0; l < n_layers && fits; l++) fits = in_dims[l] <= pupper::kTcMaxW && out_dims[l] <= pupper::kTcMaxW;
    pupper::TcParams &T = pol->tc;
    memset(&T, 0, sizeof(T));
    int nchunk = 0;
    for (int l = 0; l < n_layers && fits; l++) {
      const int K = in_dims[l], N = out_dims[l], kp8 = (K + 7) / 8 * 8, np16 = (N + 15) / 16 * 16;
      for (int n = 0; n < N; n++) T.bias[l][n] = biases[l][n];
      T.layer[l] = pupper::TcLayer{kp8, np16, N, activations[l]};
      const int kc_max = std::max(8, pupper::kTcBBytes / (np16 * 4) / 8 * 8);
      for (int k0 = 0; k0 < kp8; k0 += kc_max) {
        if (nchunk >= pupper::kTcMaxChunks) { fits = false; break; }
        const int kc = std::min(kc_max, kp8 - k0);
        std::vector<float> pack((size_t)np16 * kc, 0.f);
        const int sbo = (kc / 4) * 32;  // floats between 8-row groups
        for (int n = 0; n < N; n++)
          for (int k = k0; k < std::min(K, k0 + kc); k++) {
            const int kk = k - k0;
            pack[(size_t)(n / 8) * sbo + (kk / 4) * 32 + (n % 8) * 4 + (kk % 4)] = weights[l][(size_t)k * N + n];
          }
        void *dw = nullptr;
        e = cudaMalloc(&dw, pack.size() * sizeof(float));
        if (e == cudaSuccess) { pol->allocs.push_back(dw); e = cudaMemcpy(dw, pack.data(), pack.size() * sizeof(float), cudaMemcpyHostToDevice); }
        if (e != cudaSuccess) { pupper_policy_destroy(pol); return cuda_fail(e, "policy weight upload (tcgen05 layout)"); }
        T.chunk[nchunk++] = pupper::TcChunk{reinterpret_cast<const float *>(dw), np16 * kc * 4, l, k0, kc, k0 + kc >= kp8 ? 1 : 0};
      }
    }
    if (fits) {
      T.n_chunks = nchunk; T.n_layers = n_layers; T.in_dim = in_dims[0];
      e = cudaFuncSetAttribute(pupper::policy_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pupper::kTcSmemBytes);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(pupper::policy_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pupper::kTcSmemBytes);
      pol->tc_v2 = getenv("PUPPER_POLICY_TC2") != nullptr;
      if (e != cudaSuccess) { pupper_policy_destroy(pol); return cuda_fail(e, "cudaFuncSetAttribute(policy_tc_kernel)"); }
      pol->use_tc = true;
    }
  }
  e = cudaFuncSetAttribute(pupper::policy_kernel<1, PUPPER_POLICY_NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, pol->smem_bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(pupper::policy_kernel<3, PUPPER_POLICY_NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, pol->smem_bytes);
  if (e != cudaSuccess) { pupper_policy_destroy(pol); return cuda_fail(e, "cudaFuncSetAttribute(policy_kernel)"); }
  *out = pol;
  return PUPPER_OK;
}

#ifdef PUPPER_TC_TRACE
int pupper_policy_tc_trace(long long *host256) {
  return cudaMemcpyFromSymbol(host256, pupper::g_tc_trace, 256 * sizeof(long long)) == cudaSuccess ? 0 : -3;
}
#endif

int pupper_policy_forward(const PupperPolicy *policy, int n, const float *obs, float *action, pupper_stream_t stream) {
  return pupper_policy_forward_record(policy, n, obs, action, nullptr, stream);
}

int pupper_policy_forward_record(const PupperPolicy *policy, int n, const float *obs, float *action, float *obs_record, pupper_stream_t stream) {
  if (!policy || !obs || !action || n <= 0 || obs_record == obs) return PUPPER_EINVAL;
  pupper::PolicyParams p = policy->params;
  p.n = n; p.obs = obs; p.action = action; p.obs_record = obs_record;
  const int grid = (n + pupper::kPolRows - 1) / pupper::kPolRows;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (policy->use_tc) {
    pupper::TcParams t = policy->tc;
    t.n = n; t.obs = obs; t.action = action; t.obs_record = obs_record;
    const int grid = (n + pupper::kTcRows - 1) / pupper::kTcRows;
    if (policy->tc_v2) pupper::policy_tc2_kernel<<<grid, pupper::kTc2Threads, pupper::kTcSmemBytes, s>>>(t);
    else pupper::policy_tc_kernel<<<grid, pupper::kTcThreads, pupper::kTcSmemBytes, s>>>(t);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "policy_tc_kernel launch");
    return PUPPER_OK;
  }
  constexpr int threads = 32 * pupper::kPolWarps * PUPPER_POLICY_NSPLIT;
  if (policy->precision == PUPPER_POLICY_TF32) pupper::policy_kernel<1, PUPPER_POLICY_NSPLIT><<<grid, threads, policy->smem_bytes, s>>>(p);
  else pupper::policy_kernel<3, PUPPER_POLICY_NSPLIT><<<grid, threads, policy->smem_bytes, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "policy_kernel launch");
  return PUPPER_OK;
}

}  // extern "C"

// pupper_policy_tc.cuh -- the policy MLP forward on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// accumulators in tensor memory), used for PUPPER_POLICY_TF32 when every layer fits (widths <= 256).
//
// One CTA owns 128 rows (envs) and carries them through all layers:
//   * A operand = the 128 x K activations of the current layer, resident in shared memory in the K-major canonical
//     (no-swizzle) UMMA layout: element (r, k) at (r % 8) * 16 + (k / 4) * 128 + (r / 8) * 8192 bytes
//     (8 rows x 16 bytes core matrices; leading-dimension byte offset 128, stride byte offset 8192 = a 256-wide row block);
//   * B operand = the layer's weights as W^T [N, K], K-major, in the same canonical layout, packed on the host per
//     K-chunk so that a chunk is one contiguous block of global memory (L2 resident) that ONE thread moves with a bulk
//     async copy (cp.async.bulk, completion on an mbarrier) into one of three 32 KB buffers, two chunks ahead of the
//     MMAs -- copies, MMAs and the epilogue overlap;
//   * D = 128 lanes x N columns of float32 in tensor memory (256 columns allocated);
//   * one thread issues the layer's K/8 tcgen05.mma instructions (M = 128, N = layer width padded to 16, K = 8 each) and
//     commits them to an mbarrier; the 16 warps then read the accumulators back with tcgen05.ld (warp w: lane quadrant
//     w % 4, every fourth 16-column group), add the bias, apply the activation and write the result straight into the A
//     tile of the next layer (or to global memory after the last layer).
// TF32 operands are the raw float32 bits (the tensor core reads the top 19), as in the mma.sync kernel's TF32 mode.
#pragma once
#include <vector>

#include "../../include/pupper_policy.h"

namespace pupper {

constexpr int kTcRows = 128;             // rows per CTA = MMA M
#ifndef PUPPER_TC_THREADS
#define PUPPER_TC_THREADS 512
#endif
constexpr int kTcThreads = PUPPER_TC_THREADS;
constexpr int kTcColSplit = kTcThreads / 128;  // warps per lane quadrant
constexpr int kTcMaxW = 256;             // widest layer input / output
constexpr int kTcABytes = kTcRows * kTcMaxW * 4;   // 128 KB activation tile
constexpr int kTcBBytes = 32 * 1024;     // one weight chunk buffer
constexpr int kTcBufs = 3;               // ... of three: copies run two chunks ahead
#ifndef PUPPER_TC_PIECES
#define PUPPER_TC_PIECES 4
#endif
constexpr int kTcPieces = PUPPER_TC_PIECES;  // bulk-copy requests per chunk
constexpr int kTcSboA = (kTcMaxW / 4) * 128;       // bytes between 8-row groups of the A tile
constexpr int kTcMaxChunks = 32;

struct TcChunk {
  const float *src;   // packed chunk in global memory
  int bytes;          // np16 * kc * 4
  int layer, k0, kc;  // K range [k0, k0 + kc) of the layer (multiples of 8)
  int last;           // last chunk of its layer
};
struct TcLayer {
  int kp8, np16, n_out, act;
};
struct TcParams {
  TcChunk chunk[kTcMaxChunks];
  TcLayer layer[PUPPER_POLICY_MAX_LAYERS];
  float bias[PUPPER_POLICY_MAX_LAYERS][kTcMaxW];  // zero padded; kernel parameters live in the constant bank (uniform reads)
  int n_chunks, n_layers, in_dim, n;
  const float *obs;
  float *action;
  float *obs_record;  // optional [n][in_dim]: the staged observation rows are also written here (the rollout's trajectory slice)
};

#ifdef PUPPER_TC_TRACE  // timeline of CTA 0 (clock64 stamps; slot layout in tools/tc_trace.py), exported through pupper_policy_tc_trace
__device__ long long g_tc_trace[256];
#define TC_STAMP(slot) do { if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 64)) g_tc_trace[(slot) + (threadIdx.x ? 128 : 0)] = clock64(); } while (0)
#define TC_STAMP2(slot, second) do { if (blockIdx.x == 0) g_tc_trace[(slot) + ((second) ? 128 : 0)] = clock64(); } while (0)
#else
#define TC_STAMP(slot) ((void)0)
#define TC_STAMP2(slot, second) ((void)0)
#endif
__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: start address, leading byte offset (between the two 16-byte K halves of one MMA), stride byte
// offset (between 8-row groups); descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::tf32, D = F32, A/B = TF32, both K-major, M = 128
__device__ __forceinline__ uint32_t tc_instr_desc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tTC_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC_DONE_%=;\n\tbra TC_WAIT_%=;\n\tTC_DONE_%=:\n\t}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// Activations of the TF32 kernel.  The sigmoid family goes through the hardware tanh (tanh.approx.f32: ONE special-function
// operation, absolute error ~5e-4 of a value bounded by 1 -- the size of the TF32 rounding the operands of the next layer get
// anyway) instead of ex2 + rcp: the epilogue is bound by the special-function unit (81,920 activations per 128-row CTA), so
// this halves its cost.  Everything else falls through to the float32-accurate forms of pupper_policy.cuh.
__device__ __forceinline__ float tc_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int ACT>
__device__ __forceinline__ float tc_act(float x) {
  if (ACT == PUPPER_ACT_SWISH) { const float h = 0.5f * x; return fmaf(h, tc_tanh(h), h); }  // x sigmoid(x) = h + h tanh(h), h = x / 2
  if (ACT == PUPPER_ACT_SIGMOID) return fmaf(0.5f, tc_tanh(0.5f * x), 0.5f);
  if (ACT == PUPPER_ACT_TANH) return tc_tanh(x);
  return policy_act<ACT>(x);
}

// Epilogue of one layer for this warp: quadrant q (rows 32 q + lane), 16-column groups h, h + 2, ...
// `ready0` != 0 (warp-specialised kernel): after a group's 16 columns are in the A tile the warp arrives on the group's
// mbarrier (ready0 + 8 g), so the MMA warp can start the next layer's k-steps on those columns at once.
template <int ACT>
__device__ __forceinline__ void tc_epilogue(const TcParams &p, const TcLayer &L, const float *bias, bool last, uint32_t tmem, unsigned char *smA,
                                            int q, int h, int lane, int row0, uint32_t ready0 = 0u) {
  const int r = 32 * q + lane;
  unsigned char *arow = smA + (r & 7) * 16 + (r >> 3) * kTcSboA;
  const int row = row0 + r;
  for (int g = h; g < (L.np16 >> 4); g += kTcColSplit) {
    float v[16];
    tc_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(16 * g), v);
    const int c0 = 16 * g;
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = tc_act<ACT>(v[i] + bias[c0 + i]);
    if (!last) {
#pragma unroll
      for (int i = 0; i < 4; i++)
        *reinterpret_cast<float4 *>(arow + ((c0 >> 2) + i) * 128) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      if (ready0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // A-tile stores -> visible to the tensor core's proxy
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");    // this warp's tensor-memory reads are complete
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ready0 + 8u * g) : "memory");
      }
    } else if (row < p.n) {
#pragma unroll
      for (int i = 0; i < 16; i++)
        if (c0 + i < L.n_out) p.action[(size_t)row * L.n_out + c0 + i] = v[i];
    }
  }
}

__global__ void __launch_bounds__(kTcThreads, 1) policy_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  unsigned char *smA = tc_smem;
  unsigned char *smB = tc_smem + kTcABytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tc_smem + kTcABytes + kTcBufs * kTcBBytes);  // full[3], empty[3], layer done
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kTcBufs + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTcRows;
  auto full_bar = [&](int b) { return tc_smem_u32(&bars[b]); };
  auto empty_bar = [&](int b) { return tc_smem_u32(&bars[kTcBufs + b]); };
  // One phase per layer, waited on by every thread in order.  (The buffers' "empty" barriers cannot serve here: a thread
  // that skips their intermediate phases cannot tell phase k from phase k + 2 by parity.)
  const uint32_t done_bar = tc_smem_u32(&bars[2 * kTcBufs]);

  TC_STAMP(0);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(tc_smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weight chunk c -> buffer c % 3 as one bulk async copy; its bytes complete the buffer's "full" mbarrier
  auto issue_chunk = [&](int c) {
    const TcChunk &ch = p.chunk[c];
    const uint32_t dst = tc_smem_u32(smB + (c % kTcBufs) * kTcBBytes), bar = full_bar(c % kTcBufs);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)ch.bytes) : "memory");
    // several requests per chunk: the copy engine overlaps them (one 32 KB request alone took ~3.3 k cycles)
    const uint32_t piece = (uint32_t)ch.bytes / kTcPieces;  // chunk sizes are multiples of 512 bytes
    const unsigned char *src = reinterpret_cast<const unsigned char *>(ch.src);
#pragma unroll
    for (int i = 0; i < kTcPieces; i++)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + i * piece),
                   "l"(src + (size_t)i * piece), "r"(piece), "r"(bar)
                   : "memory");
  };
  if (tid == 32) {  // the copy thread
    for (int b = 0; b < 2 * kTcBufs + 1; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&bars[b])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    issue_chunk(0);
    if (p.n_chunks > 1) issue_chunk(1);
  }
  // layer-0 input: obs rows -> A tile (zero padded to kp8, zero rows past the batch); consecutive threads read
  // consecutive 16-byte pieces of a row (coalesced).  All loads of a pass are issued before the first store.
  {
    const int kq = p.layer[0].kp8 >> 2;  // 16-byte chunks per row
    const bool vec = (p.in_dim & 3) == 0 && (reinterpret_cast<uintptr_t>(p.obs) & 15) == 0;
    constexpr int U = 5;  // 128 rows x 72 floats = 2304 pieces = one pass of 5 per thread
    for (int base = 0; base < kTcRows * kq; base += U * kTcThreads) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int id = base + u * kTcThreads + tid;
        const int r = id / kq, c4 = id - r * kq, row = row0 + r, k = 4 * c4;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id < kTcRows * kq && row < p.n) {
          const float *o = p.obs + (size_t)row * p.in_dim + k;
          if (vec && k + 3 < p.in_dim) v[u] = __ldg(reinterpret_cast<const float4 *>(o));
          else {
            v[u].x = k < p.in_dim ? __ldg(o) : 0.f;
            v[u].y = k + 1 < p.in_dim ? __ldg(o + 1) : 0.f;
            v[u].z = k + 2 < p.in_dim ? __ldg(o + 2) : 0.f;
            v[u].w = k + 3 < p.in_dim ? __ldg(o + 3) : 0.f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int id = base + u * kTcThreads + tid;
        const int r = id / kq, c4 = id - r * kq;
        if (id < kTcRows * kq) *reinterpret_cast<float4 *>(smA + (r & 7) * 16 + c4 * 128 + (r >> 3) * kTcSboA) = v[u];
        if (p.obs_record && id < kTcRows * kq && row0 + r < p.n) {
          float *o = p.obs_record + (size_t)(row0 + r) * p.in_dim + 4 * c4;
          if (vec && 4 * c4 + 3 < p.in_dim && (reinterpret_cast<uintptr_t>(p.obs_record) & 15) == 0) *reinterpret_cast<float4 *>(o) = v[u];
          else {
            if (4 * c4 < p.in_dim) o[0] = v[u].x;
            if (4 * c4 + 1 < p.in_dim) o[1] = v[u].y;
            if (4 * c4 + 2 < p.in_dim) o[2] = v[u].z;
            if (4 * c4 + 3 < p.in_dim) o[3] = v[u].w;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();  // tensor-memory address and mbarrier inits visible
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t a_base = tc_smem_u32(smA);
  TC_STAMP(1);

  for (int c = 0; c < p.n_chunks; c++) {
    const TcChunk &ch = p.chunk[c];
    const int b = c % kTcBufs;
    // copy thread: chunk c + 2 goes where chunk c - 1 was; wait until the MMAs that read it have been committed
    if (tid == 32 && c + 2 < p.n_chunks) {
      if (c >= 1) tc_mbar_wait(empty_bar((c - 1) % kTcBufs), (uint32_t)(((c - 1) / kTcBufs) & 1));
      issue_chunk(c + 2);
    }
    if (ch.k0 == 0) {
      // first chunk of a layer: the A tile was just written through the generic proxy (staging loop / previous
      // epilogue) and the previous accumulators were just read from tensor memory
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
    }
    const TcLayer &L = p.layer[ch.layer];
    if (tid == 0) {  // the MMA thread
      tc_mbar_wait(full_bar(b), (uint32_t)((c / kTcBufs) & 1));
      TC_STAMP(8 + 4 * c);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = tc_instr_desc(L.np16);
      // descriptors of k-step 0; each further k-step (8 columns = two 16-byte core-matrix columns) advances both start
      // addresses by 256 bytes, i.e. the 14-bit address field (bytes >> 4) by 16 -- no carry out of the field below 256 KB
      uint64_t da = tc_smem_desc(a_base + (uint32_t)(ch.k0 >> 2) * 128u, 128u, (uint32_t)kTcSboA);
      uint64_t db = tc_smem_desc(tc_smem_u32(smB + b * kTcBBytes), 128u, (uint32_t)(ch.kc >> 2) * 128u);
      const int nk = ch.kc >> 3;
      tc_mma(tmem, da, db, idesc, ch.k0 > 0 ? 1u : 0u);
#pragma unroll 4
      for (int j = 1; j < nk; j++) {
        da += 16; db += 16;
        tc_mma(tmem, da, db, idesc, 1u);
      }
      tc_commit(empty_bar(b));
      if (ch.last) tc_commit(done_bar);
      TC_STAMP(9 + 4 * c);
    }
    if (ch.last) {
      tc_mbar_wait(done_bar, (uint32_t)(ch.layer & 1));  // the layer's MMAs are complete
      TC_STAMP(10 + 4 * c);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool last = ch.layer == p.n_layers - 1;
      const int q = warp & 3, h = warp >> 2;
      const float *bias = p.bias[ch.layer];
      switch (L.act) {
        case PUPPER_ACT_RELU: tc_epilogue<PUPPER_ACT_RELU>(p, L, bias, last, tmem, smA, q, h, lane, row0); break;
        case PUPPER_ACT_SIGMOID: tc_epilogue<PUPPER_ACT_SIGMOID>(p, L, bias, last, tmem, smA, q, h, lane, row0); break;
        case PUPPER_ACT_ELU: tc_epilogue<PUPPER_ACT_ELU>(p, L, bias, last, tmem, smA, q, h, lane, row0); break;
        case PUPPER_ACT_TANH: tc_epilogue<PUPPER_ACT_TANH>(p, L, bias, last, tmem, smA, q, h, lane, row0); break;
        case PUPPER_ACT_SWISH: tc_epilogue<PUPPER_ACT_SWISH>(p, L, bias, last, tmem, smA, q, h, lane, row0); break;
        case PUPPER_ACT_GELU: tc_epilogue<PUPPER_ACT_GELU>(p, L, bias, last, tmem, smA, q, h, lane, row0); break;
        case PUPPER_ACT_LEAKY_RELU: tc_epilogue<PUPPER_ACT_LEAKY_RELU>(p, L, bias, last, tmem, smA, q, h, lane, row0); break;
        default: tc_epilogue<PUPPER_ACT_LINEAR>(p, L, bias, last, tmem, smA, q, h, lane, row0); break;
      }
      // (the fences + CTA barrier at the first chunk of the next layer order these tensor-memory reads and A-tile
      //  writes before that layer's MMAs)
      TC_STAMP(11 + 4 * c);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ---- warp-specialised variant ---------------------------------------------------------------------------------------
// 16 epilogue warps + one MMA warp + one copy warp.  The accumulators are double buffered in tensor memory (512 columns:
// layer l uses columns 256 (l & 1) ...), and the epilogue hands the next layer's A tile over in 16-column groups, each
// with its own mbarrier (4 arrivals: the group's four quadrant warps), so the MMA warp issues the k-steps of layer l + 1
// while the epilogue of layer l is still producing the later columns.
// Measured (tools/tc_trace2.py): the overlap happens, but it does not shorten the chain -- TF32 MMAs at N = 128 read their
// operands from shared memory at ~122 of the 128 B/clk, so the epilogue's A-tile stores and the concurrent MMAs slow each
// other down (a 16-column group takes 2.2 k cycles instead of 1.3 k): 22.7 us at 8192 rows either way, 82 vs 84.5 us at
// 65,536 rows.  Opt-in (PUPPER_POLICY_TC2=1 at create time) until the MMAs need less shared-memory traffic (A operand from
// tensor memory, or a swizzled layout if the no-swizzle reads conflict).
constexpr int kTc2Threads = kTcThreads + 64;
constexpr int kTcGroups = kTcMaxW / 16;
__global__ void __launch_bounds__(kTc2Threads, 1) policy_tc2_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  unsigned char *smA = tc_smem;
  unsigned char *smB = tc_smem + kTcABytes;
  // full[3], empty[3], done[2], staged, ready[16]
  uint64_t *bars = reinterpret_cast<uint64_t *>(tc_smem + kTcABytes + kTcBufs * kTcBBytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kTcBufs + 3 + kTcGroups);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTcRows;
  auto full_bar = [&](int b) { return tc_smem_u32(&bars[b]); };
  auto empty_bar = [&](int b) { return tc_smem_u32(&bars[kTcBufs + b]); };
  auto done_bar = [&](int l) { return tc_smem_u32(&bars[2 * kTcBufs + (l & 1)]); };
  const uint32_t staged_bar = tc_smem_u32(&bars[2 * kTcBufs + 2]);
  const uint32_t ready0 = tc_smem_u32(&bars[2 * kTcBufs + 3]);
  constexpr int kMmaWarp = kTcThreads / 32, kCopyWarp = kMmaWarp + 1;

  if (tid == kMmaWarp * 32) TC_STAMP2(0, false);
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == kCopyWarp * 32) {
    for (int b = 0; b < 2 * kTcBufs + 2; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&bars[b])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(staged_bar), "r"(kTcThreads / 32) : "memory");
    for (int g = 0; g < kTcGroups; g++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(ready0 + 8u * g) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();  // tensor-memory address and mbarrier inits visible
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == kCopyWarp) {
    if (lane == 0) {
      for (int c = 0; c < p.n_chunks; c++) {
        const TcChunk &ch = p.chunk[c];
        const int b = c % kTcBufs;
        if (c >= kTcBufs) tc_mbar_wait(empty_bar(b), (uint32_t)(((c / kTcBufs) - 1) & 1));  // the MMAs that read this buffer are done
        const uint32_t dst = tc_smem_u32(smB + b * kTcBBytes), bar = full_bar(b);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)ch.bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(ch.src),
                     "r"((uint32_t)ch.bytes), "r"(bar)
                     : "memory");
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      const uint32_t a_base = tc_smem_u32(smA);
      uint32_t ready_phase = 0u;  // bit g: parity of the next completion of group g's barrier
      for (int c = 0; c < p.n_chunks; c++) {
        const TcChunk &ch = p.chunk[c];
        const TcLayer &L = p.layer[ch.layer];
        const int b = c % kTcBufs;
        if (ch.layer == 0) {
          if (ch.k0 == 0) tc_mbar_wait(staged_bar, 0u);
        } else {
          for (int g = ch.k0 >> 4; g <= ((ch.k0 + ch.kc - 1) >> 4); g++) {  // this chunk's columns of the A tile are in place
            if (((ch.k0 >> 4) == g && (ch.k0 & 15) != 0)) continue;        // group already waited for by the previous chunk
            tc_mbar_wait(ready0 + 8u * g, (ready_phase >> g) & 1u);
            ready_phase ^= 1u << g;
          }
        }
        TC_STAMP2(10 + 4 * c, false);  // A columns ready
        tc_mbar_wait(full_bar(b), (uint32_t)((c / kTcBufs) & 1));
        TC_STAMP2(8 + 4 * c, false);   // weights seen
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = tc_instr_desc(L.np16);
        const uint32_t d_tmem = tmem + (uint32_t)(ch.layer & 1) * 256u;
        uint64_t da = tc_smem_desc(a_base + (uint32_t)(ch.k0 >> 2) * 128u, 128u, (uint32_t)kTcSboA);
        uint64_t db = tc_smem_desc(tc_smem_u32(smB + b * kTcBBytes), 128u, (uint32_t)(ch.kc >> 2) * 128u);
        const int nk = ch.kc >> 3;
        tc_mma(d_tmem, da, db, idesc, ch.k0 > 0 ? 1u : 0u);
#pragma unroll 4
        for (int j = 1; j < nk; j++) {
          da += 16; db += 16;
          tc_mma(d_tmem, da, db, idesc, 1u);
        }
        tc_commit(empty_bar(b));
        if (ch.last) tc_commit(done_bar(ch.layer));
        TC_STAMP2(9 + 4 * c, false);   // MMAs issued
      }
    }
  } else {
    // ---- epilogue warps: stage the layer-0 input, then one epilogue per layer ------------------------------------------
    {
      const int kq = p.layer[0].kp8 >> 2;  // 16-byte chunks per row
      const bool vec = (p.in_dim & 3) == 0 && (reinterpret_cast<uintptr_t>(p.obs) & 15) == 0;
      constexpr int U = 5;
      for (int base = 0; base < kTcRows * kq; base += U * kTcThreads) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int id = base + u * kTcThreads + tid;
          const int r = id / kq, c4 = id - r * kq, row = row0 + r, k = 4 * c4;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (id < kTcRows * kq && row < p.n) {
            const float *o = p.obs + (size_t)row * p.in_dim + k;
            if (vec && k + 3 < p.in_dim) v[u] = __ldg(reinterpret_cast<const float4 *>(o));
            else {
              v[u].x = k < p.in_dim ? __ldg(o) : 0.f;
              v[u].y = k + 1 < p.in_dim ? __ldg(o + 1) : 0.f;
              v[u].z = k + 2 < p.in_dim ? __ldg(o + 2) : 0.f;
              v[u].w = k + 3 < p.in_dim ? __ldg(o + 3) : 0.f;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int id = base + u * kTcThreads + tid;
          const int r = id / kq, c4 = id - r * kq;
          if (id < kTcRows * kq) *reinterpret_cast<float4 *>(smA + (r & 7) * 16 + c4 * 128 + (r >> 3) * kTcSboA) = v[u];
          if (p.obs_record && id < kTcRows * kq && row0 + r < p.n) {
            float *o = p.obs_record + (size_t)(row0 + r) * p.in_dim + 4 * c4;
            if (vec && 4 * c4 + 3 < p.in_dim && (reinterpret_cast<uintptr_t>(p.obs_record) & 15) == 0) *reinterpret_cast<float4 *>(o) = v[u];
            else {
              if (4 * c4 < p.in_dim) o[0] = v[u].x;
              if (4 * c4 + 1 < p.in_dim) o[1] = v[u].y;
              if (4 * c4 + 2 < p.in_dim) o[2] = v[u].z;
              if (4 * c4 + 3 < p.in_dim) o[3] = v[u].w;
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(staged_bar) : "memory");
    }
    const int q = warp & 3, h = warp >> 2;
    for (int l = 0; l < p.n_layers; l++) {
      const TcLayer &L = p.layer[l];
      tc_mbar_wait(done_bar(l), (uint32_t)((l >> 1) & 1));  // the layer's MMAs are complete
      if (tid == 64) TC_STAMP2(8 + 4 * l, true);   // layer done seen
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool last = l == p.n_layers - 1;
      const float *bias = p.bias[l];
      const uint32_t d_tmem = tmem + (uint32_t)(l & 1) * 256u;
      switch (L.act) {
        case PUPPER_ACT_RELU: tc_epilogue<PUPPER_ACT_RELU>(p, L, bias, last, d_tmem, smA, q, h, lane, row0, ready0); break;
        case PUPPER_ACT_SIGMOID: tc_epilogue<PUPPER_ACT_SIGMOID>(p, L, bias, last, d_tmem, smA, q, h, lane, row0, ready0); break;
        case PUPPER_ACT_ELU: tc_epilogue<PUPPER_ACT_ELU>(p, L, bias, last, d_tmem, smA, q, h, lane, row0, ready0); break;
        case PUPPER_ACT_TANH: tc_epilogue<PUPPER_ACT_TANH>(p, L, bias, last, d_tmem, smA, q, h, lane, row0, ready0); break;
        case PUPPER_ACT_SWISH: tc_epilogue<PUPPER_ACT_SWISH>(p, L, bias, last, d_tmem, smA, q, h, lane, row0, ready0); break;
        case PUPPER_ACT_GELU: tc_epilogue<PUPPER_ACT_GELU>(p, L, bias, last, d_tmem, smA, q, h, lane, row0, ready0); break;
        case PUPPER_ACT_LEAKY_RELU: tc_epilogue<PUPPER_ACT_LEAKY_RELU>(p, L, bias, last, d_tmem, smA, q, h, lane, row0, ready0); break;
        default: tc_epilogue<PUPPER_ACT_LINEAR>(p, L, bias, last, d_tmem, smA, q, h, lane, row0, ready0); break;
      }
      if (tid == 64) TC_STAMP2(9 + 4 * l, true);   // epilogue done
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

constexpr int kTcSmemBytes = kTcABytes + kTcBufs * kTcBBytes + 256;

// Host side: can this MLP run on the tcgen05 kernel, and its chunk table.
struct TcPlan {
  bool ok = false;
  TcParams params;
};

}  // namespace pupper

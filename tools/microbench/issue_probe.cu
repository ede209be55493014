// issue_probe.cu -- micro-benchmarks that size the env-step kernel's two limits on B200 (sm_100a):
//   (1) issue cost / latency of packed FFMA2 against scalar FFMA, alone and mixed with non-FMA work;
//   (2) instruction-fetch cost of long straight-line code at 4/8/16 resident warps per SM, with and without
//       CTA barriers that keep warps in step.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_probe issue_probe.cu ; run: ./issue_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

#define REP4(x) x x x x
#define REP16(x) REP4(REP4(x))
#define REP64(x) REP16(REP4(x))
#define REP256(x) REP64(REP4(x))
#define REP1024(x) REP256(REP4(x))

// ---------------- (1) throughput / latency -------------------------------------------------------------
template <int MODE>
__global__ void tput_kernel(float *out, int iters, float b, float c, long long *cyc) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
  float2 *a2 = reinterpret_cast<float2 *>(a);
  float2 b2 = make_float2(b, b * 1.0001f), c2 = make_float2(c, c * 0.999f);
  float s0 = 0.f, s1 = 1.f, s2 = 2.f, s3 = 3.f;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {  // 16 independent FFMA
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], b, c);
    } else if (MODE == 1) {  // 8 independent FFMA2 (same flops as MODE 0)
#pragma unroll
      for (int i = 0; i < 8; i++) a2[i] = __ffma2_rn(a2[i], b2, c2);
    } else if (MODE == 2) {  // 16 FFMA + 8 integer-side ops
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], b, c);
#pragma unroll
      for (int i = 0; i < 2; i++) { s0 = s0 > b ? s1 : s2; s1 = s1 > c ? s2 : s3; s2 = s2 > b ? s3 : s0; s3 = s3 > c ? s0 : s1; }
    } else if (MODE == 3) {  // 8 FFMA2 + 8 selects
#pragma unroll
      for (int i = 0; i < 8; i++) a2[i] = __ffma2_rn(a2[i], b2, c2);
#pragma unroll
      for (int i = 0; i < 2; i++) { s0 = s0 > b ? s1 : s2; s1 = s1 > c ? s2 : s3; s2 = s2 > b ? s3 : s0; s3 = s3 > c ? s0 : s1; }
    } else if (MODE == 4) {  // latency: dependent FFMA chain
#pragma unroll
      for (int i = 0; i < 16; i++) a[0] = fmaf(a[0], b, c);
    } else if (MODE == 5) {  // latency: dependent FFMA2 chain
#pragma unroll
      for (int i = 0; i < 16; i++) a2[0] = __ffma2_rn(a2[0], b2, c2);
    } else if (MODE == 6) {  // 8 FFMA2 + 8 FFMA interleaved
#pragma unroll
      for (int i = 0; i < 4; i++) { a2[i] = __ffma2_rn(a2[i], b2, c2); a[8 + 2 * i] = fmaf(a[8 + 2 * i], b, c); a[9 + 2 * i] = fmaf(a[9 + 2 * i], b, c); }
    }
  }
  long long t1 = clock64();
  float s = s0 + s1 + s2 + s3;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---------------- (2) straight-line code streaming -----------------------------------------------------
// one "group" = 8 instructions on 8 independent chains (no dependency stalls inside a group at 2 warps/SMSP)
#define GROUP a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c); a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
#define GROUP2 p0 = __ffma2_rn(p0, b2, c2); p1 = __ffma2_rn(p1, b2, c2); p2 = __ffma2_rn(p2, b2, c2); p3 = __ffma2_rn(p3, b2, c2);

#define SYNCPT if (SYNC) __syncthreads();
#define CHUNK256 if (PACKED) { REP16(GROUP2) REP16(GROUP2) REP16(GROUP2) REP16(GROUP2) } else { REP16(GROUP) REP16(GROUP) } SYNCPT
#define BODY2K CHUNK256 CHUNK256 CHUNK256 CHUNK256 CHUNK256 CHUNK256 CHUNK256 CHUNK256
template <int SIZE, bool SYNC, bool PACKED>   // SIZE x 2048 straight-line instructions per outer iteration (FFMA, or FFMA2 when PACKED)
__global__ void __launch_bounds__(128) stream_kernel(float *out, int iters, float b, float c, long long *cyc) {
  extern __shared__ float pad[];
  float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f, a4 = 4.f, a5 = 5.f, a6 = 6.f, a7 = 7.f;
  float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
  float2 b2 = make_float2(b, b * 1.0001f), c2 = make_float2(c, c * 0.999f);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (SIZE >= 1) { BODY2K }
    if (SIZE >= 2) { BODY2K }
    if (SIZE >= 3) { BODY2K }
    if (SIZE >= 4) { BODY2K }
    if (SIZE >= 5) { BODY2K }
    if (SIZE >= 6) { BODY2K }
    if (SIZE >= 7) { BODY2K }
    if (SIZE >= 8) { BODY2K }
    if (SIZE >= 9) { BODY2K }
    if (SIZE >= 10) { BODY2K }
    if (SIZE >= 11) { BODY2K }
    if (SIZE >= 12) { BODY2K }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y + pad[threadIdx.x & 1];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double max_cycles(long long *d, int n) {
  std::vector<long long> h(n);
  CK(cudaMemcpy(h.data(), d, n * sizeof(long long), cudaMemcpyDeviceToHost));
  long long m = 0; for (auto v : h) m = v > m ? v : m;
  return (double)m;
}

int main() {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  float *out; long long *cyc;
  CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&cyc, 8192 * sizeof(long long)));
  printf("SMs %d\n", nsm);
  // (1) per-SMSP issue rates: W warps per SM in one CTA per SM
  const char *names[] = {"16xFFMA", "8xFFMA2", "16xFFMA+8xFSEL", "8xFFMA2+8xFSEL", "dep FFMA x16", "dep FFMA2 x16", "4xFFMA2+8xFFMA"};
  const int ninstr[] = {16, 8, 24, 16, 16, 16, 12};
  for (int mode = 0; mode < 7; mode++)
    for (int wps : {4, 8, 16, 32}) {
      int iters = 20000;
      auto launch = [&](int it) {
        switch (mode) {
          case 0: tput_kernel<0><<<nsm, wps * 32>>>(out, it, 1.0001f, 0.5f, cyc); break;
          case 1: tput_kernel<1><<<nsm, wps * 32>>>(out, it, 1.0001f, 0.5f, cyc); break;
          case 2: tput_kernel<2><<<nsm, wps * 32>>>(out, it, 1.0001f, 0.5f, cyc); break;
          case 3: tput_kernel<3><<<nsm, wps * 32>>>(out, it, 1.0001f, 0.5f, cyc); break;
          case 4: tput_kernel<4><<<nsm, wps * 32>>>(out, it, 1.0001f, 0.5f, cyc); break;
          case 5: tput_kernel<5><<<nsm, wps * 32>>>(out, it, 1.0001f, 0.5f, cyc); break;
          case 6: tput_kernel<6><<<nsm, wps * 32>>>(out, it, 1.0001f, 0.5f, cyc); break;
        }
      };
      launch(100); CK(cudaDeviceSynchronize());
      launch(iters); CK(cudaDeviceSynchronize());
      double c = max_cycles(cyc, nsm);
      double per_warp = (double)iters * ninstr[mode];
      printf("tput %-18s warps/SM %2d : %.3f cycles per warp-instr per warp ; %.3f warp-instr/clk/SMSP\n", names[mode], wps, c / per_warp,
             per_warp * wps / 4.0 / c);
    }
  // (2) streaming: 2 CTAs x 128 threads per SM (smem padding 100 KB forces 2 CTAs/SM), or 1 CTA (4 warps) / 4 CTAs (16 warps)
  auto run_stream = [&](auto kern, const char *name, int kgroups, bool packed, int ctas_per_sm) {
    size_t smem = ctas_per_sm == 1 ? 200 * 1024 : (ctas_per_sm == 2 ? 100 * 1024 : 50 * 1024);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long per_iter = (long long)kgroups * 2048;
    int iters = (int)(8000000LL / per_iter); if (iters < 2) iters = 2;
    kern<<<nsm * ctas_per_sm, 128, smem>>>(out, 2, 1.0001f, 0.5f, cyc); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    kern<<<nsm * ctas_per_sm, 128, smem>>>(out, iters, 1.0001f, 0.5f, cyc);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double c = max_cycles(cyc, nsm * ctas_per_sm);
    double instr = (double)iters * per_iter;
    printf("stream %-28s body %6lld instr (%5.0f KB) warps/SM %2d : %.3f cyc/instr/warp ; %.3f warp-instr/clk/SMSP ; %.3f ms\n", name, per_iter,
           per_iter * 16.0 / 1024, ctas_per_sm * 4, c / instr, instr * ctas_per_sm / c, ms);
  };
#define RUN(KG, SY, PK) for (int cps : {1, 2, 4}) run_stream(stream_kernel<KG, SY, PK>, #KG "x2k sync=" #SY " packed=" #PK, KG, PK, cps)
  if (getenv("PROBE_FINE")) {  // finer sweep of the body size: where the fetch rate steps
    RUN(1, false, false); RUN(2, false, false); RUN(3, false, false); RUN(4, false, false); RUN(5, false, false); RUN(6, false, false);
    RUN(7, false, false); RUN(8, false, false); RUN(9, false, false); RUN(10, false, false); RUN(11, false, false); RUN(12, false, false);
    return 0;
  }
  RUN(1, false, false); RUN(2, false, false); RUN(4, false, false); RUN(8, false, false);
  RUN(4, true, false); RUN(8, true, false);
  RUN(1, false, true); RUN(4, false, true); RUN(8, false, true);
  return 0;
}

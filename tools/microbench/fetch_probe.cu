// fetch_probe.cu -- instruction-fetch behaviour of B200 (sm_100a) for long straight-line loop bodies when the resident warps are
// NOT in step (each warp enters the loop body at a different place), as in the env-step kernel.  Per body size (32..192 KB) and
// resident warps per SM (4, 8, 12, 16): warp-instructions per clock per SMSP with aligned and with staggered entry points.
// The instruction mix is two dependent FFMA chains per warp (a lone warp issues one instruction per ~2 cycles from cache-resident
// code), so fetch stalls add to dependency stalls the way they do in the real kernel.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fetch_probe fetch_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define REP4(x) x x x x
#define REP16(x) REP4(REP4(x))
#define REP64(x) REP16(REP4(x))
#define PAIR a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c);
#define CHUNK256 REP64(PAIR PAIR)          /* 256 instructions = 4 KB */
#define CASE(i) case i: CHUNK256
#define CASES8(b) CASE(b) CASE(b + 1) CASE(b + 2) CASE(b + 3) CASE(b + 4) CASE(b + 5) CASE(b + 6) CASE(b + 7)

template <int KB32>  // body = KB32 * 32 KB = KB32 * 8 chunks
__global__ void __launch_bounds__(128) body_kernel(float *out, int iters, float b, float c, int stagger, long long *cyc) {
  extern __shared__ float pad[];
  float a0 = threadIdx.x, a1 = 1.f;
  const int nchunk = KB32 * 8;
  const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
  int entry = stagger ? (gw * stagger) % nchunk : 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    switch (entry) {
      CASES8(0)
      if (KB32 >= 2) { CASES8(8) }
      if (KB32 >= 3) { CASES8(16) }
      if (KB32 >= 4) { CASES8(24) }
      if (KB32 >= 5) { CASES8(32) }
      if (KB32 >= 6) { CASES8(40) }
    }
    entry = 0;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + pad[threadIdx.x & 1];
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
  auto run = [&](auto kern, int kb32, int ctas_per_sm, int stagger) {
    size_t smem = ctas_per_sm == 1 ? 200 * 1024 : (ctas_per_sm == 2 ? 100 * 1024 : (ctas_per_sm == 3 ? 70 * 1024 : 50 * 1024));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long per_iter = (long long)kb32 * 2048;
    int iters = (int)(6000000LL / per_iter); if (iters < 4) iters = 4;
    kern<<<nsm * ctas_per_sm, 128, smem>>>(out, 2, 1.0001f, 0.5f, stagger, cyc); CK(cudaDeviceSynchronize());
    kern<<<nsm * ctas_per_sm, 128, smem>>>(out, iters, 1.0001f, 0.5f, stagger, cyc); CK(cudaDeviceSynchronize());
    double c = max_cycles(cyc, nsm * ctas_per_sm);
    double instr = (double)iters * per_iter;  // upper bound for staggered warps (the first pass is partial)
    printf("body %3d KB  warps/SM %2d  stagger %2d : %.3f cyc/instr/warp ; %.3f warp-instr/clk/SMSP\n", kb32 * 32, ctas_per_sm * 4, stagger,
           c / instr, instr * ctas_per_sm / c);
  };
#define RUNALL(K) for (int cps : {1, 2, 3, 4}) for (int st : {0, 5}) run(body_kernel<K>, K, cps, st)
  RUNALL(1); RUNALL(2); RUNALL(3); RUNALL(4); RUNALL(5); RUNALL(6);
  return 0;
}

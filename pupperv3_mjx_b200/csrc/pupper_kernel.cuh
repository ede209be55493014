// pupper_kernel.cuh -- the batched PupperV3Env reset/step kernel for sm_100a.
//
// Mapping: FOUR LANES PER ENV ("quad"), lane k of a quad owns leg k (3 bodies, 3 hinge dofs, the knee
// and foot collision spheres, the foot site, 3 action/observation channels); the floating base is
// replicated in the 4 lanes and combined with 2-step xor-shuffle butterflies (bitwise identical in
// all 4 lanes).  8 envs per warp, state in SoA so that a warp's 8 envs x 4 legs touch 4 32-byte
// sectors per row.  Per-env DR leaves, sphere centres and the active-contact list live in shared
// memory; everything else stays in registers.  The 18x18 joint-space matrices (M and the Newton
// Hessian) are never formed densely: the kinematic tree gives them an arrow structure
// [base 6x6 | 4 x (3x6 coupling, 3x3 leg)] that is factorised leaves-first (per-lane 3x3 Cholesky,
// Schur complement reduced over the quad, replicated 6x6 Cholesky).  Only a leg-leg sphere contact
// couples two legs; that rare case gathers the system to lane 0 of the quad and solves it densely.
//
// What is computed is the reference's PupperV3Env.step (pupperv3_mjx/environment.py:348-483) with the
// physics it delegates to brax/mjx.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <type_traits>

#include "../../include/pupper_env.h"
#include "pupper_math.cuh"

namespace pupper {

#ifndef PUPPER_BLOCK
#define PUPPER_BLOCK 128
#endif
#ifndef PUPPER_PHASE_SYNC
#define PUPPER_PHASE_SYNC 1
#endif
// A/B switches of individual optimisations (tools/jobs/ab.sh builds the variants and times them on one box)
#ifndef PUPPER_FAST_ROT
#define PUPPER_FAST_ROT 1  // kinematics: the 15-instruction form of the quaternion rotation (pupper_math.cuh rotate_fast)
#endif
#if PUPPER_FAST_ROT
#define KROT rotate_fast
#else
#define KROT rotate
#endif
#ifndef PUPPER_ZFOLD
#define PUPPER_ZFOLD 1   // structural zeros folded by hand (the compiler may not drop 0*x terms)
#endif

// CTA barriers at phase boundaries keep a CTA's warps within one instruction-cache window of each other, so a line fetched
// by one warp serves the others (tools/microbench/fetch_probe.cu: warps of an SM that run a 96 KB loop body in step issue 4x
// the instructions per clock of warps spread over it).  A/B on B200 with the round-2 kernel: +2.1 % at 65,536 envs, -0.4 % at 4096; twice as many barriers: -0.3 %.
#if PUPPER_PHASE_SYNC
#define PHASE_SYNC() __syncthreads()
#else
#define PHASE_SYNC() ((void)0)
#endif
#ifndef PUPPER_SYNC_MASK
#define PUPPER_SYNC_MASK 0x0aa  // which of the 9 phase barriers of forward() are compiled in (A/B on one box, 3 rounds: 0x0aa +0.7 % at 65,536 envs and +1.3 % at 4096 over all nine; 0x000 is -2..4 %)
#endif
#define PHASE_SYNC_AT(i) do { if ((PUPPER_SYNC_MASK >> (i)) & 1) PHASE_SYNC(); } while (0)
constexpr int kBlock = PUPPER_BLOCK;   // threads per CTA
constexpr int kEnvsPerBlock = kBlock / 4;
constexpr int kMaxCon = 5;            // contact slots per env (max_contact_points <= 5)
constexpr float kMinVal = 1e-15f, kMinImp = 1e-4f, kMaxImp = 0.9999f;
constexpr float kInf = 3.0e38f;

struct ConstBlock;
struct KParams {
  const ConstBlock *consts;  // device copy of the model description, env configuration and derived constants (one block)
  int n_envs;
  PupperState st;
  PupperDR dr;
  int has_dr;
  const float *action;
  const uint32_t *keys;  // reset only
  PupperStepOut out;
  PupperEpisode ep;
  int has_ep;
  PupperRand rand;  // external randoms (debug / parity instantiation only)
  int has_rand;
};

struct ContactSlot {  // one ACTIVE contact (dist < 0) of an env, shared by the quad
  float r[3];       // contact point relative to the subtree COM
  float frame[9];   // rows: normal (geom1 -> geom2), tangent 1, tangent 2
  float mu, D, b, kimp;  // friction, efc_D of its 4 pyramid rows, damping gain, k*imp*dist
  float dist;
  int code1, code2;  // leg*4 + depth (1: link2 / knee sphere, 2: link3 / foot sphere) or -1 for the world
  int s1, s2;        // sphere indices (or -1) for the collision rewards
  int ty, box;       // pair type (0 plane-sphere, 1 sphere-box, 2 sphere-sphere) and box index
};

struct EnvShared {
  float mass[13], inertia[39], ipos[3], friction, kp, kd;  // DR leaves (or nominal values); must stay first (staging loop)
  // forward-pass quantities the env level reads after the last substep (the "stale" mix); parked
  // here so that nothing of them occupies registers across the solver
  float st_torso[16];    // pos3 rot4 ang3 vel3 com3
  float st_leg[4][15];   // per leg: foot_site3 lower_pos3 lower_ang3 lower_vel3 frc3
  float st_hits[2];      // knee / torso collision counters
  // env-level values that must survive the physics loop
  float lv_act[12], lv_kick[2];
  uint32_t lv_cmd_rng[2];
  float sph[8][3];
  ContactSlot con[kMaxCon];
  int ncon;
};

struct DerivedConsts {  // computed once on the host in pupper_model_create
  float fric_loss[PUPPER_NV];  // friction loss of instantiated friction rows (0: no row)
  float fric_D[PUPPER_NV];     // efc_D of the friction-loss rows (pos = 0 -> constant)
  float fric_rf[PUPPER_NV];    // R * frictionloss: half-width of the quadratic zone
  float fric_b;                // velocity gain of aref
  float box_rbound[PUPPER_MAX_BOX];
  int ss_pair[24];             // leg-leg sphere pairs in MJX order: a | b << 8  ({(a,b): a<b, a/2 != b/2}, lexicographic)
};

// Everything the kernel reads that is constant for a model: one device block, copied to the head of each CTA's shared memory
// in 16-byte pieces (pupper_model_create builds it).
struct alignas(16) ConstBlock {
  PupperModelDesc m;
  PupperEnvCfg c;
  DerivedConsts d;
};
static_assert(sizeof(ConstBlock) % 16 == 0, "ConstBlock is copied in 16-byte pieces");

struct BlockShared : ConstBlock {
  EnvShared env[kEnvsPerBlock];
  float rows[3 * kMaxCon * kBlock];  // contact-edge row scalars: [buffer][contact][thread]
  float4 lsf[6 * kBlock];            // line search, per friction-loss row j: [2j] (Jaref, jv, R f, qc), [2j+1] linear-zone corrections
  float4 lsq[kMaxCon * kBlock];      // line search, per contact-edge row: (Jaref, jv, qa, qb) ...
  float lsc[kMaxCon * kBlock];       // ... and qc   ([contact][thread]; zero for slots past the env's contacts)
  float mat[45 * kBlock];            // per-lane copy of the mass matrix blocks: [element][thread]
};

__device__ __forceinline__ float qsum(float v, unsigned qm) {
  v += __shfl_xor_sync(qm, v, 1);
  v += __shfl_xor_sync(qm, v, 2);
  return v;
}
__device__ __forceinline__ V3 qsum3(V3 v, unsigned qm) { return V3{qsum(v.x, qm), qsum(v.y, qm), qsum(v.z, qm)}; }
#if PUPPER_F2
__device__ __forceinline__ S6 qsum6(S6 v, unsigned qm) {  // same sums, the additions on register pairs
  P6 p = pack6(v);
#pragma unroll
  for (int sft = 1; sft <= 2; sft <<= 1) {
    p.p0 = __fadd2_rn(p.p0, make_float2(__shfl_xor_sync(qm, p.p0.x, sft), __shfl_xor_sync(qm, p.p0.y, sft)));
    p.p1 = __fadd2_rn(p.p1, make_float2(__shfl_xor_sync(qm, p.p1.x, sft), __shfl_xor_sync(qm, p.p1.y, sft)));
    p.p2 = __fadd2_rn(p.p2, make_float2(__shfl_xor_sync(qm, p.p2.x, sft), __shfl_xor_sync(qm, p.p2.y, sft)));
  }
  return unpack6(p);
}
#else
__device__ __forceinline__ S6 qsum6(S6 v, unsigned qm) { return S6{qsum3(v.a, qm), qsum3(v.l, qm)}; }
#endif
__device__ __forceinline__ float qbcast(float v, int src, unsigned qm, int qbase) { return __shfl_sync(qm, v, qbase + src); }

// Arrow-structured symmetric matrix over (base 6 | leg 3), one leg per lane.
struct TreeMat {
  float B[21];     // base block, packed lower triangle, row-major: (i,j) at i*(i+1)/2+j   (replicated)
  float C[3][6];   // leg-base coupling
  float D[6];      // leg block, packed lower triangle
};
__device__ __forceinline__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

#if PUPPER_F2
// Packed FP32x2 views of a 6-row (three register pairs) with a broadcast scalar: r = s * a (+ b), element-wise IEEE.
struct R6 { float2 p0, p1, p2; };
__device__ __forceinline__ R6 row6(const float a[6]) { return R6{make_float2(a[0], a[1]), make_float2(a[2], a[3]), make_float2(a[4], a[5])}; }
__device__ __forceinline__ void unrow6(R6 r, float a[6]) { a[0] = r.p0.x; a[1] = r.p0.y; a[2] = r.p1.x; a[3] = r.p1.y; a[4] = r.p2.x; a[5] = r.p2.y; }
__device__ __forceinline__ R6 mul6(float s, R6 a) { const float2 ss = make_float2(s, s); return R6{__fmul2_rn(ss, a.p0), __fmul2_rn(ss, a.p1), __fmul2_rn(ss, a.p2)}; }
__device__ __forceinline__ R6 fmar6(float s, R6 a, R6 b) {
  const float2 ss = make_float2(s, s);
  return R6{__ffma2_rn(ss, a.p0, b.p0), __ffma2_rn(ss, a.p1, b.p1), __ffma2_rn(ss, a.p2, b.p2)};
}
__device__ __forceinline__ float dotr6(R6 a, R6 b) {
  const float2 t = __ffma2_rn(a.p2, b.p2, __ffma2_rn(a.p1, b.p1, __fmul2_rn(a.p0, b.p0)));
  return t.x + t.y;
}
__device__ __forceinline__ float2 qsum2(float2 v, unsigned qm) {
  v = __fadd2_rn(v, make_float2(__shfl_xor_sync(qm, v.x, 1), __shfl_xor_sync(qm, v.y, 1)));
  return __fadd2_rn(v, make_float2(__shfl_xor_sync(qm, v.x, 2), __shfl_xor_sync(qm, v.y, 2)));
}
#endif

// y = A x for the arrow matrix (A.B must already hold the full base block)
__device__ __forceinline__ void tree_matvec(const TreeMat &A, const float xb[6], const float xl[3], float yb[6], float yl[3], unsigned qm) {
#if PUPPER_F2
  const R6 X = row6(xb), C0 = row6(A.C[0]), C1 = row6(A.C[1]), C2 = row6(A.C[2]);
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float s = dotr6(j == 0 ? C0 : (j == 1 ? C1 : C2), X);
#pragma unroll
    for (int i = 0; i < 3; i++) s = fmaf(A.D[i >= j ? tri(i, j) : tri(j, i)], xl[i], s);
    yl[j] = s;
  }
  R6 t = fmar6(xl[2], C2, fmar6(xl[1], C1, mul6(xl[0], C0)));
  t = R6{qsum2(t.p0, qm), qsum2(t.p1, qm), qsum2(t.p2, qm)};
  float tb[6];
  unrow6(t, tb);
#pragma unroll
  for (int d = 0; d < 6; d++) {
    float s = tb[d];
#pragma unroll
    for (int i = 0; i < 6; i++) s = fmaf(A.B[i >= d ? tri(i, d) : tri(d, i)], xb[i], s);
    yb[d] = s;
  }
#else
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) s = fmaf(A.D[i >= j ? tri(i, j) : tri(j, i)], xl[i], s);
#pragma unroll
    for (int d = 0; d < 6; d++) s = fmaf(A.C[j][d], xb[d], s);
    yl[j] = s;
  }
#pragma unroll
  for (int d = 0; d < 6; d++) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) s = fmaf(A.C[j][d], xl[j], s);
    s = qsum(s, qm);
#pragma unroll
    for (int i = 0; i < 6; i++) s = fmaf(A.B[i >= d ? tri(i, d) : tri(d, i)], xb[i], s);
    yb[d] = s;
  }
#endif
}

// Same product with the matrix read from shared memory (layout: element i of [B(21) | C(18) | D(6)] at sm[i*kBlock])
__device__ __forceinline__ void tree_matvec_smem(const float *sm, const float xb[6], const float xl[3], float yb[6], float yl[3], unsigned qm) {
  const float *B = sm, *C = sm + 21 * kBlock, *D = sm + 39 * kBlock;
#if PUPPER_F2
  float c[3][6];
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int d = 0; d < 6; d++) c[j][d] = C[(j * 6 + d) * kBlock];
  const R6 X = row6(xb), C0 = row6(c[0]), C1 = row6(c[1]), C2 = row6(c[2]);
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float s = dotr6(j == 0 ? C0 : (j == 1 ? C1 : C2), X);
#pragma unroll
    for (int i = 0; i < 3; i++) s = fmaf(D[(i >= j ? tri(i, j) : tri(j, i)) * kBlock], xl[i], s);
    yl[j] = s;
  }
  R6 t = fmar6(xl[2], C2, fmar6(xl[1], C1, mul6(xl[0], C0)));
  t = R6{qsum2(t.p0, qm), qsum2(t.p1, qm), qsum2(t.p2, qm)};
  float tb[6];
  unrow6(t, tb);
#pragma unroll
  for (int d = 0; d < 6; d++) {
    float s = tb[d];
#pragma unroll
    for (int i = 0; i < 6; i++) s = fmaf(B[(i >= d ? tri(i, d) : tri(d, i)) * kBlock], xb[i], s);
    yb[d] = s;
  }
#else
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) s = fmaf(D[(i >= j ? tri(i, j) : tri(j, i)) * kBlock], xl[i], s);
#pragma unroll
    for (int d = 0; d < 6; d++) s = fmaf(C[(j * 6 + d) * kBlock], xb[d], s);
    yl[j] = s;
  }
#pragma unroll
  for (int d = 0; d < 6; d++) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) s = fmaf(C[(j * 6 + d) * kBlock], xl[j], s);
    s = qsum(s, qm);
#pragma unroll
    for (int i = 0; i < 6; i++) s = fmaf(B[(i >= d ? tri(i, d) : tri(d, i)) * kBlock], xb[i], s);
    yb[d] = s;
  }
#endif
}

// In-place leaves-first Cholesky.  On exit: D = L_k (3x3 lower), C = Y_k = L_k^-1 C_k,
// B = chol(B + sum_k (Badd_k - Y_k^T Y_k)) (lower, replicated).  Badd may be null.
// The DIAGONAL entries of D and B hold the reciprocals 1 / l_ii (the solves multiply instead of dividing).
__device__ __forceinline__ void tree_factor(TreeMat &A, const float *Badd, unsigned qm) {
  float i00 = rsqrtf(A.D[0]);
  float l10 = A.D[1] * i00, l20 = A.D[3] * i00;
  float i11 = rsqrtf(A.D[2] - l10 * l10);
  float l21 = (A.D[4] - l20 * l10) * i11;
  float i22 = rsqrtf(A.D[5] - l20 * l20 - l21 * l21);
  A.D[0] = i00; A.D[1] = l10; A.D[2] = i11; A.D[3] = l20; A.D[4] = l21; A.D[5] = i22;  // diagonals hold 1 / l_ii
#if PUPPER_F2
  const R6 Y0 = mul6(i00, row6(A.C[0]));
  const R6 Y1 = mul6(i11, fmar6(-l10, Y0, row6(A.C[1])));
  const R6 Y2 = mul6(i22, fmar6(-l21, Y1, fmar6(-l20, Y0, row6(A.C[2]))));
  unrow6(Y0, A.C[0]); unrow6(Y1, A.C[1]); unrow6(Y2, A.C[2]);
  // Schur complement, row i of the lower triangle as register pairs (j, j+1); the half of a pair beyond the diagonal is unused
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const float n0 = -A.C[0][i], n1 = -A.C[1][i], n2 = -A.C[2][i];
#pragma unroll
    for (int jp = 0; jp <= i; jp += 2) {
      const bool both = jp + 1 <= i;
      float2 s = Badd ? make_float2(Badd[tri(i, jp)], both ? Badd[tri(i, jp + 1)] : 0.f) : make_float2(0.f, 0.f);
      const float2 y0 = jp == 0 ? Y0.p0 : (jp == 2 ? Y0.p1 : Y0.p2), y1 = jp == 0 ? Y1.p0 : (jp == 2 ? Y1.p1 : Y1.p2),
                   y2 = jp == 0 ? Y2.p0 : (jp == 2 ? Y2.p1 : Y2.p2);
      s = __ffma2_rn(make_float2(n0, n0), y0, s);
      s = __ffma2_rn(make_float2(n1, n1), y1, s);
      s = __ffma2_rn(make_float2(n2, n2), y2, s);
      if (both) {
        s = qsum2(s, qm);
        A.B[tri(i, jp)] += s.x; A.B[tri(i, jp + 1)] += s.y;
      } else {
        A.B[tri(i, jp)] += qsum(s.x, qm);
      }
    }
  }
#else
#pragma unroll
  for (int d = 0; d < 6; d++) {
    float y0 = A.C[0][d] * i00;
    float y1 = (A.C[1][d] - l10 * y0) * i11;
    float y2 = (A.C[2][d] - l20 * y0 - l21 * y1) * i22;
    A.C[0][d] = y0; A.C[1][d] = y1; A.C[2][d] = y2;
  }
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      float s = Badd ? Badd[tri(i, j)] : 0.f;
      s = fmaf(-A.C[0][i], A.C[0][j], s);
      s = fmaf(-A.C[1][i], A.C[1][j], s);
      s = fmaf(-A.C[2][i], A.C[2][j], s);
      A.B[tri(i, j)] += qsum(s, qm);
    }
#endif
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      float s = A.B[tri(i, j)];
#pragma unroll
      for (int p = 0; p < j; p++) s = fmaf(-A.B[tri(i, p)], A.B[tri(j, p)], s);
      A.B[tri(i, j)] = (i == j) ? rsqrtf(s) : s * A.B[tri(j, j)];  // diagonals hold 1 / l_ii
    }
}

// x = A^-1 g with the factor produced by tree_factor (reciprocal diagonals)
__device__ __forceinline__ void tree_solve(const TreeMat &F, const float gb[6], const float gl[3], float xb[6], float xl[3], unsigned qm) {
  float z0 = gl[0] * F.D[0];
  float z1 = (gl[1] - F.D[1] * z0) * F.D[2];
  float z2 = (gl[2] - F.D[3] * z0 - F.D[4] * z1) * F.D[5];
  float y[6];
#if PUPPER_F2
  const R6 C0 = row6(F.C[0]), C1 = row6(F.C[1]), C2 = row6(F.C[2]);
  {
    R6 t = fmar6(z0, C0, fmar6(z1, C1, mul6(z2, C2)));
    t = R6{qsum2(t.p0, qm), qsum2(t.p1, qm), qsum2(t.p2, qm)};
    const R6 G = row6(gb);
    unrow6(R6{__fadd2_rn(G.p0, make_float2(-t.p0.x, -t.p0.y)), __fadd2_rn(G.p1, make_float2(-t.p1.x, -t.p1.y)),
              __fadd2_rn(G.p2, make_float2(-t.p2.x, -t.p2.y))}, y);
  }
#else
#pragma unroll
  for (int d = 0; d < 6; d++) {
    float s = fmaf(F.C[0][d], z0, fmaf(F.C[1][d], z1, F.C[2][d] * z2));
    y[d] = gb[d] - qsum(s, qm);
  }
#endif
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float s = y[i];
#pragma unroll
    for (int p = 0; p < i; p++) s = fmaf(-F.B[tri(i, p)], y[p], s);
    y[i] = s * F.B[tri(i, i)];
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    float s = y[i];
#pragma unroll
    for (int p = i + 1; p < 6; p++) s = fmaf(-F.B[tri(p, i)], xb[p], s);
    xb[i] = s * F.B[tri(i, i)];
  }
#if PUPPER_F2
  const R6 X = row6(xb);
  const float w0 = z0 - dotr6(C0, X), w1 = z1 - dotr6(C1, X), w2 = z2 - dotr6(C2, X);
#else
  float w0 = z0, w1 = z1, w2 = z2;
#pragma unroll
  for (int d = 0; d < 6; d++) {
    w0 = fmaf(-F.C[0][d], xb[d], w0);
    w1 = fmaf(-F.C[1][d], xb[d], w1);
    w2 = fmaf(-F.C[2][d], xb[d], w2);
  }
#endif
  xl[2] = w2 * F.D[5];
  xl[1] = (w1 - F.D[4] * xl[2]) * F.D[2];
  xl[0] = (w0 - F.D[1] * xl[1] - F.D[3] * xl[2]) * F.D[0];
}

__device__ __noinline__ float impedance_curve_generic(float x, float mid, float power) {
  const float ia = (1.f / powf(mid, power - 1.f)) * powf(x, power);
  const float ib = 1.f - (1.f / powf(1.f - mid, power - 1.f)) * powf(1.f - x, power);
  return x < mid ? ia : ib;
}

// solref/solimp -> (k, b, imp) at constraint violation `pos`
__device__ __forceinline__ void kbi(const float *solref, const float *solimp, float timestep, float pos, float &k, float &b, float &imp) {
  float timeconst = fmaxf(solref[0], 2.f * timestep), dampratio = solref[1];
  float dmin = fminf(fmaxf(solimp[0], kMinImp), kMaxImp), dmax = fminf(fmaxf(solimp[1], kMinImp), kMaxImp);
  float width = fmaxf(kMinVal, solimp[2]), mid = fminf(fmaxf(solimp[3], kMinImp), kMaxImp), power = fmaxf(1.f, solimp[4]);
  k = 1.f / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  b = 2.f / (dmax * timeconst);
  if (solref[0] <= 0.f) k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0.f) b = -solref[1] / dmax;
  float x = fabsf(pos) / width;
  float y;
  if (power == 2.f) {  // the model's setting; the general power law lives out of line to keep the hot path compact
    const float ia = (1.f / mid) * (x * x);
    const float ib = 1.f - (1.f / (1.f - mid)) * ((1.f - x) * (1.f - x));
    y = x < mid ? ia : ib;
  } else {
    y = impedance_curve_generic(x, mid, power);
  }
  float im = dmin + y * (dmax - dmin);
  im = fminf(fmaxf(im, dmin), dmax);
  if (x > 1.f) im = dmax;
  imp = im;
}

__device__ __forceinline__ void make_frame(V3 n, float fr[9]) {
  V3 a = n;
  normalize3(a);
  V3 b = fabsf(a.y) < 0.5f ? V3{0.f, 1.f, 0.f} : V3{0.f, 0.f, 1.f};
  float ab = dot(a, b);
  b = b - ab * a;
  normalize3(b);
  V3 c = cross(a, b);
  fr[0] = a.x; fr[1] = a.y; fr[2] = a.z; fr[3] = b.x; fr[4] = b.y; fr[5] = b.z; fr[6] = c.x; fr[7] = c.y; fr[8] = c.z;
}

// sphere vs box (box = 6-face polytope).  Returns dist; pos / n in world frame.
__device__ __forceinline__ float sphere_box(V3 cw, float radius, const float *bpos, const float *bmat, const float *size, V3 &pos_out, V3 &n_out) {
  V3 d = cw - V3{bpos[0], bpos[1], bpos[2]};
  V3 c = V3{bmat[0] * d.x + bmat[3] * d.y + bmat[6] * d.z, bmat[1] * d.x + bmat[4] * d.y + bmat[7] * d.z,
            bmat[2] * d.x + bmat[5] * d.y + bmat[8] * d.z};
  // faces in MJX order: -y, -z, +x, +y, +z, -x ; support_f = (c - n r - v0).n
  const float sx = size[0], sy = size[1], sz = size[2];
  float sup[6] = {-(c.y) - radius - sy, -(c.z) - radius - sz, c.x - radius - sx, c.y - radius - sy, c.z - radius - sz, -(c.x) - radius - sx};
  int best = 0;
  float bs = 0.f;
#pragma unroll
  for (int f = 0; f < 6; f++) {
    float s = sup[f] >= 0.f ? -1e12f : sup[f];
    if (f == 0 || s > bs) { best = f; bs = s; }
  }
  // face axis / sign and its two in-plane axes; vertices listed in the MJX winding
  int ax = (best == 2 || best == 5) ? 0 : ((best == 0 || best == 3) ? 1 : 2);
  float sgn = (best == 2 || best == 3 || best == 4) ? 1.f : -1.f;
  float half[3] = {sx, sy, sz};
  float cc[3] = {c.x, c.y, c.z};
  float pt[3] = {c.x, c.y, c.z};
  pt[ax] = sgn * half[ax];  // projection of the centre on the face plane
  // face vertex loops (vertex id v = 4*ix + 2*iy + iz)
  const int FACE[6][4] = {{0, 4, 5, 1}, {0, 2, 6, 4}, {6, 7, 5, 4}, {2, 3, 7, 6}, {1, 5, 7, 3}, {0, 1, 3, 2}};
  float fn[3] = {0.f, 0.f, 0.f};
  fn[ax] = sgn;
  float fv[4][3];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int v = FACE[best][i];
    fv[i][0] = (v & 4) ? sx : -sx;
    fv[i][1] = (v & 2) ? sy : -sy;
    fv[i][2] = (v & 1) ? sz : -sz;
  }
  bool inside = true;
  int idx = 0;
  float be = 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float *p0 = fv[(i + 3) & 3], *p1 = fv[i];
    V3 e = V3{p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    V3 en = cross(e, V3{fn[0], fn[1], fn[2]});
    float ed = dot(V3{pt[0] - p0[0], pt[1] - p0[1], pt[2] - p0[2]}, en);
    if (!(ed <= 0.f)) inside = false;
    bool degenerate = (en.x == 0.f && en.y == 0.f && en.z == 0.f);
    float v = (degenerate || ed < 0.f) ? 1e12f : ed;
    if (i == 0 || v < be) { be = v; idx = i; }
  }
  if (!inside) {
    const float *a = fv[(idx + 3) & 3], *b = fv[idx];
    V3 ab = V3{b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    V3 pa = V3{pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
    float t = dot(pa, ab) / (dot(ab, ab) + 1e-6f);
    t = fminf(fmaxf(t, 0.f), 1.f);
    pt[0] = a[0] + t * ab.x; pt[1] = a[1] + t * ab.y; pt[2] = a[2] + t * ab.z;
  }
  V3 n = V3{pt[0] - cc[0], pt[1] - cc[1], pt[2] - cc[2]};
  float dn = normalize3(n);
  V3 p = V3{(pt[0] + (cc[0] + n.x * radius)) * 0.5f, (pt[1] + (cc[1] + n.y * radius)) * 0.5f, (pt[2] + (cc[2] + n.z * radius)) * 0.5f};
  n_out = V3{bmat[0] * n.x + bmat[1] * n.y + bmat[2] * n.z, bmat[3] * n.x + bmat[4] * n.y + bmat[5] * n.z, bmat[6] * n.x + bmat[7] * n.y + bmat[8] * n.z};
  pos_out = V3{bmat[0] * p.x + bmat[1] * p.y + bmat[2] * p.z + bpos[0], bmat[3] * p.x + bmat[4] * p.y + bmat[5] * p.z + bpos[1],
               bmat[6] * p.x + bmat[7] * p.y + bmat[8] * p.z + bpos[2]};
  return dn - radius;
}

// ---------------------------------------------------------------------------------------------------
// per-lane physics state
// ---------------------------------------------------------------------------------------------------
struct LaneState {
  float qb[7];      // base position + quaternion (replicated)
  float ql[3];      // this leg's joint angles
  float vb[6], vl[3];
  float wb[6], wl[3];  // qacc_warmstart
  float ctrl[3];
};

// quantities of the forward pass that the env level reads after the last substep ("stale" mix)
struct StaleOut {
  V3 torso_pos;
  Q4 torso_rot;
  V3 torso_ang, torso_vel;   // xd of the torso (body-origin world velocity)
  V3 com;
  V3 foot_site;              // this leg's foot site position
  V3 lower_pos;              // xpos of this leg's lower-leg body
  V3 lower_ang, lower_vel;   // xd of the lower leg
  float frc[3];              // qfrc_actuator of this leg
  float knee_hits, torso_hits;  // collision reward counters (whole env, replicated)
};

__device__ __forceinline__ StaleOut load_stale(const struct EnvShared &es, int k);
struct DbgOut {  // only filled when DBG
  V3 pos[3]; Q4 rot[3]; V3 ang[3], vel[3];
  float qacc_b[6], qacc_l[3];
  int solver[8];  // PupperStepOut.dbg_solver (replicated over the quad)
  // PupperStepOut.dbg_efc, this lane's rows: friction-loss / limit rows of its 3 hinges, its pyramid edge of every slot
  float efc_fD[3], efc_fA[3], efc_lD[3], efc_lA[3], efc_cD[kMaxCon], efc_cA[kMaxCon];
};

__device__ __forceinline__ StaleOut load_stale(const EnvShared &es, int k) {
  StaleOut so;
  const float *t = es.st_torso, *q = es.st_leg[k];
  so.torso_pos = V3{t[0], t[1], t[2]}; so.torso_rot = Q4{t[3], t[4], t[5], t[6]};
  so.torso_ang = V3{t[7], t[8], t[9]}; so.torso_vel = V3{t[10], t[11], t[12]}; so.com = V3{t[13], t[14], t[15]};
  so.foot_site = V3{q[0], q[1], q[2]}; so.lower_pos = V3{q[3], q[4], q[5]}; so.lower_ang = V3{q[6], q[7], q[8]};
  so.lower_vel = V3{q[9], q[10], q[11]}; so.frc[0] = q[12]; so.frc[1] = q[13]; so.frc[2] = q[14];
  so.knee_hits = es.st_hits[0]; so.torso_hits = es.st_hits[1];
  return so;
}

// Per-lane participation in contact c: bits 0-1 depth as body1 (-), bits 2-3 depth as body2 (+).
__device__ __forceinline__ int participation(const ContactSlot &s, int k) {
  int p = 0;
  if (s.code1 >= 0 && (s.code1 >> 2) == k) p |= (s.code1 & 3);
  if (s.code2 >= 0 && (s.code2 >> 2) == k) p |= (s.code2 & 3) << 2;
  return p;
}

// friction-loss row zone at residual x: 1 quadratic, 2 linear (x <= -R f), 3 linear (x >= R f)
__device__ __forceinline__ int fzone(float x, float rf) { return x <= -rf ? 2 : (x >= rf ? 3 : 1); }

struct LSPoint {
  float alpha, cost, d0, d1;
};
__device__ __forceinline__ LSPoint ls_select(bool c, const LSPoint &a, const LSPoint &b) {  // c ? a : b, field by field
  return LSPoint{c ? a.alpha : b.alpha, c ? a.cost : b.cost, c ? a.d0 : b.d0, c ? a.d1 : b.d1};
}
__device__ __forceinline__ bool in_bracket(const LSPoint &x, const LSPoint &y) {
  return (((x.d0 < y.d0) & (y.d0 < 0.f)) | ((x.d0 > y.d0) & (y.d0 > 0.f))) != 0;  // bitwise: no short-circuit branches
}

// Spatial velocity of this leg's link2 (depth 1) under each of N generalized vectors v = [vb (6, replicated) | vl (this
// leg's 3)]; link3 adds vl[2]*cd[2].
template <int N>
__device__ __forceinline__ void link2_velocity(const S6 cd[3], const V3 ba[3], const V3 bo[3], const float (&vb)[N][6], const float (&vl)[N][3], S6 (&W1)[N]) {
#pragma unroll
  for (int n = 0; n < N; n++) {
    S6 W0;
    W0.a = vb[n][3] * ba[0] + vb[n][4] * ba[1] + vb[n][5] * ba[2];
    W0.l = V3{vb[n][0], vb[n][1], vb[n][2]} + vb[n][3] * bo[0] + vb[n][4] * bo[1] + vb[n][5] * bo[2];
    W1[n] = fma6(vl[n][1], cd[1], fma6(vl[n][0], cd[0], W0));
  }
}

// Constraint-row scalars J.v of the world-vs-leg contacts, computed by the lane whose leg touches ("own" contacts): that
// lane has the chain's cdofs, so it evaluates the contact-point velocity once and derives all FOUR pyramid-edge rows
// (Jn +- mu Jt1, Jn +- mu Jt2) from it; the trip count is the warp-wide maximum of contacts per LEG (1-2), not of contacts
// per env.  The rows land in shared memory in the edge-per-lane layout the rest of the solver reads:
// row of edge e of contact c at out[c*kBlock + e], with `out` pointing at lane 0 of this quad.
// COST (N == 3, vectors = qvel, warm start, qacc_smooth): rows 1 and 2 are stored as Jaref = J.x - aref (buffers B, C)
// and the constraint cost at both start points is accumulated in cw / cs; row 0 only feeds aref.
// !COST (N == 1): the row is stored in buffer A.
template <int N, bool COST, bool TAP = false>
__device__ __forceinline__ void contact_rows_own(const EnvShared &es, const S6 cd[3], const V3 ba[3], const V3 bo[3], const float (&vb)[N][6],
                                                 const float (&vl)[N][3], int own_list, int own_count, int nown_w, int part_all, float *out,
                                                 float &cw, float &cs) {
  S6 W1[N];
  link2_velocity<N>(cd, ba, bo, vb, vl, W1);
#pragma unroll 1
  for (int i = 0; i < nown_w; i++) {  // warp-uniform trip count; lanes with fewer contacts sit the iteration out
    if (i < own_count) {
      const int c = (own_list >> (3 * i)) & 7;
      const ContactSlot &s = es.con[c];
      const V3 r = V3{s.r[0], s.r[1], s.r[2]};
      const int pc = (part_all >> (4 * c)) & 15;
      const int d1 = pc & 3, d2 = (pc >> 2) & 3, dep = d1 | d2;
      const float sg = d2 ? 1.f : -1.f;  // + as body2, - as body1
      const float mu = s.mu;
      float row[N][4];
#pragma unroll
      for (int n = 0; n < N; n++) {
        const S6 W = fma6(dep == 2 ? vl[n][2] : 0.f, cd[2], W1[n]);
        V3 pv = W.l + cross(W.a, r);  // velocity of the contact point as carried by the touching link
        pv = V3{sg * pv.x, sg * pv.y, sg * pv.z};
        const float jn = s.frame[0] * pv.x + s.frame[1] * pv.y + s.frame[2] * pv.z;
        const float jt1 = s.frame[3] * pv.x + s.frame[4] * pv.y + s.frame[5] * pv.z;
        const float jt2 = s.frame[6] * pv.x + s.frame[7] * pv.y + s.frame[8] * pv.z;
        row[n][0] = fmaf(mu, jt1, jn); row[n][1] = fmaf(-mu, jt1, jn); row[n][2] = fmaf(mu, jt2, jn); row[n][3] = fmaf(-mu, jt2, jn);
      }
      float *o = out + c * kBlock;
      if (COST) {
        const float b = s.b, kimp = s.kimp, hD = 0.5f * s.D;
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float aref = -b * row[0][e] - kimp;
          const float xw = row[N > 1 ? 1 : 0][e] - aref, xs = row[N > 2 ? 2 : 0][e] - aref;
          if (TAP) o[e] = aref;  // debug instantiation: buffer A (unused until the line search) carries aref to the dbg_efc tap
          o[kMaxCon * kBlock + e] = xw; o[2 * kMaxCon * kBlock + e] = xs;  // Jaref at the warm start / at qacc_smooth
          const float mw = fminf(xw, 0.f), ms = fminf(xs, 0.f);
          cw = fmaf(hD * mw, mw, cw);
          cs = fmaf(hD * ms, ms, cs);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; e++) o[e] = row[0][e];
      }
    }
  }
}

// Same rows for LEG-LEG contacts (rare): two legs carry the contact point, so every lane evaluates ITS pyramid edge of
// the slot and the two legs' contributions are summed over the quad.  Walks the slots of `slots_w` (warp-uniform: slots
// holding a leg-leg contact in some env of the warp) and writes those that are leg-leg in this env (`slots`):
// out[(n*kMaxCon + c)*kBlock] with `out` pointing at this lane's column.
template <int N>
__device__ __forceinline__ void contact_rows_legleg(const EnvShared &es, const S6 cd[3], const V3 ba[3], const V3 bo[3], const float (&vb)[N][6],
                                                    const float (&vl)[N][3], float esgn, bool et2, unsigned qm, int part_all, int slots, int slots_w,
                                                    float *out) {
  S6 W1[N];
  link2_velocity<N>(cd, ba, bo, vb, vl, W1);
#pragma unroll 1
  for (int rem = slots_w; rem; rem &= rem - 1) {
    const int c = __ffs(rem) - 1;
    const ContactSlot &s = es.con[c];
    const V3 r = V3{s.r[0], s.r[1], s.r[2]};
    const bool mine = (slots >> c) & 1;
    const int pc = mine ? (part_all >> (4 * c)) & 15 : 0;
    const int d1 = pc & 3, d2 = (pc >> 2) & 3, dep = d1 | d2;
    const float sg = (d2 ? 1.f : 0.f) - (d1 ? 1.f : 0.f);  // + as body2, - as body1, 0 if this leg is not involved
    const float t0 = et2 ? s.frame[6] : s.frame[3], t1 = et2 ? s.frame[7] : s.frame[4], t2 = et2 ? s.frame[8] : s.frame[5];
    const float em = esgn * s.mu;
#pragma unroll
    for (int n = 0; n < N; n++) {
      const S6 W = fma6(dep == 2 ? vl[n][2] : 0.f, cd[2], W1[n]);
      V3 pv = W.l + cross(W.a, r);
      pv = qsum3(V3{sg * pv.x, sg * pv.y, sg * pv.z}, qm);
      const float jn = s.frame[0] * pv.x + s.frame[1] * pv.y + s.frame[2] * pv.z;
      const float jt = t0 * pv.x + t1 * pv.y + t2 * pv.z;
      if (mine) out[(n * kMaxCon + c) * kBlock] = fmaf(em, jt, jn);
    }
  }
}

// Rare path: a leg-leg sphere contact couples two legs, so the Hessian is no longer arrow-shaped.
// Gather the system on lane 0 of the quad and solve it densely (H holds M + diagonal row terms).
struct DenseIO {  // memory-backed copy made only on the rare path, so the hot path's arrays stay in registers
  TreeMat H;
  float gb[6], gl[3];
  S6 cd[3];
  V3 ba[3], bo[3];
  float hb[6], hl[3];
};
__device__ __noinline__ void dense_newton_direction(const EnvShared &es, int ncon, int k, bool need, int qbase, DenseIO &io, const float *rowJ) {
  const TreeMat &H = io.H;
  const float *gb = io.gb, *gl = io.gl;
  const S6 *cd = io.cd;
  const V3 *ba = io.ba, *bo = io.bo;
  float *hb = io.hb, *hl = io.hl;
  const unsigned qm = 0xffffffffu;
  float Hd[18 * 18], Ld[18 * 18], xd[18], yd[18];
  float cdall[12][6];
  for (int i = 0; i < 18 * 18; i++) Hd[i] = 0.f;
  for (int i = 0; i < 18; i++) xd[i] = 0.f;
  for (int kk = 0; kk < 4; kk++) {
    for (int j = 0; j < 3; j++) {
      float c6[6] = {cd[j].a.x, cd[j].a.y, cd[j].a.z, cd[j].l.x, cd[j].l.y, cd[j].l.z};
      for (int i = 0; i < 6; i++) cdall[3 * kk + j][i] = __shfl_sync(qm, c6[i], qbase + kk);
      for (int d = 0; d < 6; d++) { float v = __shfl_sync(qm, H.C[j][d], qbase + kk); Hd[(6 + 3 * kk + j) * 18 + d] = v; Hd[d * 18 + 6 + 3 * kk + j] = v; }
      for (int jj = 0; jj <= j; jj++) {
        float v = __shfl_sync(qm, H.D[tri(j, jj)], qbase + kk);
        Hd[(6 + 3 * kk + j) * 18 + 6 + 3 * kk + jj] = v; Hd[(6 + 3 * kk + jj) * 18 + 6 + 3 * kk + j] = v;
      }
      yd[6 + 3 * kk + j] = __shfl_sync(qm, gl[j], qbase + kk);
    }
  }
  for (int i = 0; i < 6; i++) { yd[i] = gb[i]; for (int j = 0; j <= i; j++) { Hd[i * 18 + j] = H.B[tri(i, j)]; Hd[j * 18 + i] = H.B[tri(i, j)]; } }
  float dall[kMaxCon][4];  // active-row weights of every contact's 4 pyramid edges
  for (int c = 0; c < kMaxCon; c++) {
    float de = 0.f;
    if (c < ncon && rowJ[c * kBlock] < 0.f) de = es.con[c].D;
    for (int e = 0; e < 4; e++) dall[c][e] = __shfl_sync(qm, de, qbase + e);
  }
  if (k == 0 && need) {
    for (int c = 0; c < ncon; c++) {
      const ContactSlot &s = es.con[c];
      V3 r = V3{s.r[0], s.r[1], s.r[2]};
      float Jc[3][18];
      for (int d = 0; d < 18; d++) {
        V3 col = V3{0.f, 0.f, 0.f};
        for (int side = 0; side < 2; side++) {
          int code = side ? s.code2 : s.code1;
          if (code < 0) continue;
          float sg = side ? 1.f : -1.f;
          int leg = code >> 2, dep = code & 3;
          V3 cc = V3{0.f, 0.f, 0.f};
          if (d < 3) cc = V3{d == 0 ? 1.f : 0.f, d == 1 ? 1.f : 0.f, d == 2 ? 1.f : 0.f};
          else if (d < 6) cc = bo[d - 3] + cross(ba[d - 3], r);
          else if ((d - 6) / 3 == leg && (d - 6) % 3 <= dep) {
            const float *q = cdall[d - 6];
            cc = V3{q[3], q[4], q[5]} + cross(V3{q[0], q[1], q[2]}, r);
          }
          col = col + sg * cc;
        }
        for (int i = 0; i < 3; i++) Jc[i][d] = s.frame[3 * i] * col.x + s.frame[3 * i + 1] * col.y + s.frame[3 * i + 2] * col.z;
      }
      for (int e = 0; e < 4; e++) {
        float de = dall[c][e];
        if (de == 0.f) continue;
        float sg = (e & 1) ? -s.mu : s.mu;
        const float *Jt = Jc[1 + (e >> 1)];
        for (int i = 0; i < 18; i++) {
          float ji = (Jc[0][i] + sg * Jt[i]) * de;
          if (ji == 0.f) continue;
          for (int j = 0; j < 18; j++) Hd[i * 18 + j] += ji * (Jc[0][j] + sg * Jt[j]);
        }
      }
    }
    for (int i = 0; i < 18; i++)
      for (int j = 0; j <= i; j++) {
        float s = Hd[i * 18 + j];
        for (int p = 0; p < j; p++) s -= Ld[i * 18 + p] * Ld[j * 18 + p];
        Ld[i * 18 + j] = (i == j) ? sqrtf(s) : s / Ld[j * 18 + j];
      }
    for (int i = 0; i < 18; i++) {
      float s = yd[i];
      for (int p = 0; p < i; p++) s -= Ld[i * 18 + p] * xd[p];
      xd[i] = s / Ld[i * 18 + i];
    }
    for (int i = 17; i >= 0; i--) {
      float s = xd[i];
      for (int p = i + 1; p < 18; p++) s -= Ld[p * 18 + i] * xd[p];
      xd[i] = s / Ld[i * 18 + i];
    }
  }
  for (int d = 0; d < 6; d++) { float v = __shfl_sync(qm, xd[d], qbase); if (need) hb[d] = v; }
  for (int kk = 0; kk < 4; kk++)
    for (int j = 0; j < 3; j++) {
      float v = __shfl_sync(qm, xd[6 + 3 * kk + j], qbase);
      if (kk == k && need) hl[j] = v;
    }
}

// efc_D and aref of a joint-limit row at violation p < 0 (rare; out of line)
__device__ __noinline__ float2 limit_row(const float *solref, const float *solimp, float dt, float p, float signed_vel, float invweight) {
  float kk, bb, imp;
  kbi(solref, solimp, dt, p, kk, bb, imp);
  return make_float2(1.f / fmaxf(invweight * (1.f - imp) / imp, kMinVal), -bb * signed_vel - kk * imp * p);
}

// Memory-backed arguments of the out-of-line leg-leg row evaluation (N generalized vectors)
template <int N>
struct LegLegIO {
  S6 cd[3];
  V3 ba[3], bo[3];
  float vb[N][6], vl[N][3];
};
__device__ __noinline__ void legleg_rows3(const EnvShared &es, const LegLegIO<3> &io, float esgn, bool et2, int part_all, int slots, int slots_w, float *out) {
  contact_rows_legleg<3>(es, io.cd, io.ba, io.bo, io.vb, io.vl, esgn, et2, 0xffffffffu, part_all, slots, slots_w, out);
}
__device__ __noinline__ void legleg_rows1(const EnvShared &es, const LegLegIO<1> &io, float esgn, bool et2, int part_all, int slots, int slots_w, float *out) {
  contact_rows_legleg<1>(es, io.cd, io.ba, io.bo, io.vb, io.vl, esgn, et2, 0xffffffffu, part_all, slots, slots_w, out);
}
// Leg-leg slots hold raw rows (J.qvel in buffer A, J.warmstart in B, J.qacc_smooth in C of this lane's column `rowA`):
// turns B and C into Jaref and returns this lane's cost at both start points.
__device__ __noinline__ float2 legleg_start_cost(const EnvShared &es, int ss_mask, float *rowA) {
  float *rowB = rowA + kMaxCon * kBlock, *rowC = rowB + kMaxCon * kBlock;
  float cw = 0.f, cs = 0.f;
#pragma unroll 1
  for (int rem = ss_mask; rem; rem &= rem - 1) {
    const int c = __ffs(rem) - 1;
    const ContactSlot &s = es.con[c];
    float aref = -s.b * rowA[c * kBlock] - s.kimp;
    float xw = rowB[c * kBlock] - aref, xs = rowC[c * kBlock] - aref;
    rowB[c * kBlock] = xw; rowC[c * kBlock] = xs;
    const float mw = fminf(xw, 0.f), ms = fminf(xs, 0.f);
    cw = fmaf(0.5f * s.D * mw, mw, cw);
    cs = fmaf(0.5f * s.D * ms, ms, cs);
  }
  return make_float2(cw, cs);
}
// Wrenches of the leg-leg contacts on this leg's link2 / link3 chains (io[0], io[1]); `ja0` points at edge 0 of slot 0 of
// the chosen start's Jaref buffer as seen from the quad.
__device__ __noinline__ void legleg_wrench(const EnvShared &es, int ss_mask, const float *ja0, int part_all, S6 (&io)[2]) {
  S6 S1 = io[0], S2 = io[1];
#pragma unroll 1
  for (int rem = ss_mask; rem; rem &= rem - 1) {
    const int c = __ffs(rem) - 1;
    const ContactSlot &s = es.con[c];
    const float *ja = ja0 + c * kBlock;
    const float D = s.D, mu = s.mu;
    const float f0 = -D * fminf(ja[0], 0.f), f1 = -D * fminf(ja[1], 0.f), f2 = -D * fminf(ja[2], 0.f), f3 = -D * fminf(ja[3], 0.f);
    const float Fn = (f0 + f1) + (f2 + f3), Ft1 = mu * (f0 - f1), Ft2 = mu * (f2 - f3);
    const V3 g = V3{s.frame[0] * Fn + s.frame[3] * Ft1 + s.frame[6] * Ft2, s.frame[1] * Fn + s.frame[4] * Ft1 + s.frame[7] * Ft2,
                    s.frame[2] * Fn + s.frame[5] * Ft1 + s.frame[8] * Ft2};
    const V3 r = V3{s.r[0], s.r[1], s.r[2]};
    const S6 w = S6{cross(r, g), g};
    const int pc = (part_all >> (4 * c)) & 15;
    const int dd1 = pc & 3, dd2 = (pc >> 2) & 3;
    const float s1w = (dd2 == 1 ? 1.f : 0.f) - (dd1 == 1 ? 1.f : 0.f), s2w = (dd2 == 2 ? 1.f : 0.f) - (dd1 == 2 ? 1.f : 0.f);
    S1 = fma6(s1w, w, S1);
    S2 = fma6(s2w, w, S2);
  }
  io[0] = S1; io[1] = S2;
}
// This lane's view of the env's contact slots when they were not all plane contacts of its own spheres
struct SlotScan {
  int n_ss, css, ss_mask, part_all, own_list, own_count;
  float knee_hits, torso_hits;
};
__device__ __noinline__ void scan_slots(uint32_t knee_mask, uint32_t torso_mask, const EnvShared &es, int ncon, int k, SlotScan &o) {
  int n_ss = 0, css = 0, ss_mask = 0, part_all = 0, own_list = 0, own_count = 0;
  float knee_hits = 0.f, torso_hits = 0.f;
#pragma unroll 1
  for (int c = 0; c < ncon; c++) {
    const ContactSlot &s = es.con[c];
    const bool is_ss = (s.code1 >= 0 && s.code2 >= 0);
    if (is_ss) { n_ss++; css = c; ss_mask |= 1 << c; }
    {
      const int pc = participation(s, k);
      part_all |= pc << (4 * c);
      if (pc && !is_ss) { own_list |= c << (3 * own_count); own_count++; }
    }
    if (s.s1 >= 0) { knee_hits += (float)((knee_mask >> s.s1) & 1u); torso_hits += (float)((torso_mask >> s.s1) & 1u); }
    if (s.s2 >= 0) { knee_hits += (float)((knee_mask >> s.s2) & 1u); torso_hits += (float)((torso_mask >> s.s2) & 1u); }
  }
  o.n_ss = n_ss; o.css = css; o.ss_mask = ss_mask; o.part_all = part_all; o.own_list = own_list; o.own_count = own_count;
  o.knee_hits = knee_hits; o.torso_hits = torso_hits;
}

// ---- collision candidates and the rare collision paths (out of line: see woodbury_direction) -------------------------
struct Cand {  // one candidate contact of a lane
  float dist;      // penetration (< 0) or kInf
  V3 pos, n;       // world contact point, normal geom1 -> geom2
  int code1, code2, s1, s2, box;
};
// raw candidate -> contact slot (the quad completes frame, friction and impedance afterwards); ty: 0 plane, 1 box, 2 leg-leg
__device__ __forceinline__ void write_raw_slot(ContactSlot &s, const Cand &c, V3 C, int ty) {
  s.r[0] = c.pos.x - C.x; s.r[1] = c.pos.y - C.y; s.r[2] = c.pos.z - C.z;
  s.frame[0] = c.n.x; s.frame[1] = c.n.y; s.frame[2] = c.n.z;
  s.dist = c.dist;
  s.code1 = c.code1; s.code2 = c.code2; s.s1 = c.s1; s.s2 = c.s2; s.ty = ty; s.box = c.box;
}

// sphere-box: the broad phase keeps the max_geom_pairs pairs with the smallest bounding-sphere distance over all 8*nbox
// pairs (pair index = sphere*nbox + box, ties to the lower index); the pair of rank i gets its narrow phase on lane i.
// One pass per lane keeps its own 4 best pairs sorted, a 4-round merge over the quad ranks them, and the 4 narrow phases
// then run side by side.  Called by the whole warp when the model has boxes.
__device__ __noinline__ void box_candidate(const BlockShared &sh, const EnvShared &es, int k, int qbase, V3 sc0, V3 sc1, Cand &out) {
  const PupperModelDesc &m = sh.m;
  const unsigned qm = 0xffffffffu;
  out.dist = kInf; out.pos = out.n = V3{0.f, 0.f, 0.f};
  out.code1 = out.code2 = out.s1 = out.s2 = -1; out.box = 0;
  const int nbox = m.nbox, maxp = m.max_geom_pairs;
  float k0 = kInf, k1 = kInf, k2 = kInf, k3 = kInf;
  int i0 = 0x7fffffff, i1 = 0x7fffffff, i2 = 0x7fffffff, i3 = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const float rs = m.sphere_radius[2 * k + i];
    const V3 sci = i == 0 ? sc0 : sc1;
#pragma unroll 1
    for (int bx = 0; bx < nbox; bx++) {
      V3 d = V3{m.box_pos[bx][0], m.box_pos[bx][1], m.box_pos[bx][2]} - sci;
      const float key = sqrtf(dot(d, d)) - (rs + sh.d.box_rbound[bx]);
      const int id = (2 * k + i) * nbox + bx;  // ids grow along the pass, so strict '<' keeps ties in index order
      const bool c0 = key < k0, c1 = key < k1, c2 = key < k2, c3 = key < k3;
      k3 = c2 ? k2 : (c3 ? key : k3); i3 = c2 ? i2 : (c3 ? id : i3);
      k2 = c1 ? k1 : (c2 ? key : k2); i2 = c1 ? i1 : (c2 ? id : i2);
      k1 = c0 ? k0 : (c1 ? key : k1); i1 = c0 ? i0 : (c1 ? id : i1);
      k0 = c0 ? key : k0; i0 = c0 ? id : i0;
    }
  }
  const int nr = min(min(maxp, 4), 8 * nbox);
  int mine = -1;
#pragma unroll 1
  for (int r = 0; r < nr; r++) {
    float bk = k0;
    int bi = i0;
#pragma unroll
    for (int sft = 1; sft <= 2; sft <<= 1) {
      float ok = __shfl_xor_sync(qm, bk, sft);
      int oi = __shfl_xor_sync(qm, bi, sft);
      if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
    }
    if (bi == i0) { k0 = k1; i0 = i1; k1 = k2; i1 = i2; k2 = k3; i2 = i3; k3 = kInf; i3 = 0x7fffffff; }  // this lane's head won: pop it
    if (r == k) mine = bi;
  }
  if (mine >= 0) {  // narrow phase of the pair ranked k
    const int sph = mine / nbox, bx = mine - sph * nbox;
    V3 c = V3{es.sph[sph][0], es.sph[sph][1], es.sph[sph][2]};
    V3 pp, nn;
    float d = sphere_box(c, m.sphere_radius[sph], m.box_pos[bx], m.box_mat[bx], m.box_size[bx], pp, nn);
    out.dist = d < 0.f ? d : kInf;
    out.pos = pp; out.n = nn;
    out.code1 = (sph >> 1) * 4 + 1 + (sph & 1); out.code2 = -1; out.s1 = sph; out.s2 = -1; out.box = bx;
  }
}

// sphere-sphere (leg-leg): the max_geom_pairs closest of the 24 pairs, 6 per lane in MJX pair order; the pair of rank r
// gets its narrow phase on lane r.  Called by the whole warp when some pair of the warp penetrates.
__device__ __noinline__ void ss_candidate(const BlockShared &sh, const EnvShared &es, int k, int qbase, Cand &out) {
  const PupperModelDesc &m = sh.m;
  const unsigned qm = 0xffffffffu;
  out.dist = kInf; out.pos = out.n = V3{0.f, 0.f, 0.f};
  out.code1 = out.code2 = out.s1 = out.s2 = -1; out.box = 0;
  const int maxp = m.max_geom_pairs;
  float pd[6];
  int pa[6], pb[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const int ab = sh.d.ss_pair[6 * k + i];
    const int a = ab & 255, b = ab >> 8;
    V3 d = V3{es.sph[b][0] - es.sph[a][0], es.sph[b][1] - es.sph[a][1], es.sph[b][2] - es.sph[a][2]};
    pd[i] = sqrtf(dot(d, d)) - (m.sphere_radius[a] + m.sphere_radius[b]);
    pa[i] = a; pb[i] = b;
  }
  uint32_t taken = 0u;
  for (int r = 0; r < maxp && r < 24; r++) {
    float bk = kInf;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 6; i++)
      if (!((taken >> i) & 1u) && (pd[i] < bk || (pd[i] == bk && 6 * k + i < bi))) { bk = pd[i]; bi = 6 * k + i; }
#pragma unroll
    for (int sft = 1; sft <= 2; sft <<= 1) {
      float ok = __shfl_xor_sync(qm, bk, sft);
      int oi = __shfl_xor_sync(qm, bi, sft);
      if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
    }
    int owner = min(bi / 6, 3), li = bi - owner * 6;
    int a = 0, b = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) if (i == li) { a = pa[i]; b = pb[i]; }
    a = __shfl_sync(qm, a, qbase + owner);
    b = __shfl_sync(qm, b, qbase + owner);
    if (owner == k) taken |= 1u << li;
    if ((r & 3) == k && r < 4 && bk < 0.f) {
      V3 ca_ = V3{es.sph[a][0], es.sph[a][1], es.sph[a][2]}, cb_ = V3{es.sph[b][0], es.sph[b][1], es.sph[b][2]};
      V3 n = cb_ - ca_;
      float dn = normalize3(n);
      if (dn == 0.f) n = V3{1.f, 0.f, 0.f};
      float d = dn - (m.sphere_radius[a] + m.sphere_radius[b]);
      out.dist = d < 0.f ? d : kInf;
      out.n = n;
      out.pos = ca_ + (m.sphere_radius[a] + d * 0.5f) * n;
      out.code1 = (a >> 1) * 4 + 1 + (a & 1); out.code2 = (b >> 1) * 4 + 1 + (b & 1); out.s1 = a; out.s2 = b;
    }
  }
}

// The ranking cut: some env of the warp has more penetrating candidates than contact slots, so the max_contact_points
// smallest distances over [plane 0..7, box 8..11, sphere-sphere 12..15] are kept, ties to the lower index.  Returns the
// env's contact count; the winners' raw slots are written in rank order.  Called by the whole warp.
struct CutIO { Cand c[4]; };
__device__ __noinline__ int cut_candidates(EnvShared &es, int k, int qbase, int maxc, V3 C, CutIO &io) {
  const unsigned qm = 0xffffffffu;
  float cdist[4];
#pragma unroll
  for (int i = 0; i < 4; i++) cdist[i] = io.c[i].dist;
  int ncon = 0;
  for (int r = 0; r < maxc; r++) {
    // local best of this lane's 4 candidates (ids grow with i, so the first minimum is also the lowest id)
    const float bk = fminf(fminf(cdist[0], cdist[1]), fminf(cdist[2], cdist[3]));
    const int bl = cdist[0] == bk ? 0 : (cdist[1] == bk ? 1 : (cdist[2] == bk ? 2 : 3));
    const int bi = bk < kInf ? (bl < 2 ? 2 * k + bl : (bl == 2 ? 8 + k : 12 + k)) : 0x7fffffff;
    float mk = bk;
    int mi = bi;
#pragma unroll
    for (int sft = 1; sft <= 2; sft <<= 1) {
      float ok = __shfl_xor_sync(qm, mk, sft);
      int oi = __shfl_xor_sync(qm, mi, sft);
      const bool take = (ok < mk) | ((ok == mk) & (oi < mi));
      mk = take ? ok : mk; mi = take ? oi : mi;
    }
    if (!__any_sync(qm, mk < 0.f)) break;  // warp-uniform exit
    if (mk < 0.f && mi == bi && bk < 0.f) {  // this lane owns the winner: publish the raw candidate in slot r
      Cand w = io.c[bl];
      w.dist = bk;
      write_raw_slot(es.con[r], w, C, bl < 2 ? 0 : bl - 1);
#pragma unroll
      for (int i = 0; i < 4; i++) if (i == bl) cdist[i] = kInf;
    }
    if (mk < 0.f) ncon = r + 1;
  }
  return ncon;
}

// Rare path, one leg-leg contact: with A the arrow matrix already factorised in io.H (M + diagonal rows + world-vs-leg
// contact blocks) and J, D the <=4 active pyramid-edge rows of that contact, (A + J^T D J) x = g is solved as
//   (D^-1 + J A^-1 J^T) y = J A^-1 g ,  x = A^-1 (g - J^T y).
// The rows have no base columns (the base moves both legs alike) and 3 entries on each of the two legs, held by those
// legs' lanes.  Quads without such a contact run along with zero rows (y = 0).  Called by the whole warp; kept OUT OF
// LINE with a memory-backed argument block so that its ~800 instructions stay out of the substep loop's address range
// (the loop is instruction-fetch bound: tools/line_hist.py).  On exit io.tb / io.tl hold x.
struct WoodburyIO {
  TreeMat H;
  float gb[6], gl[3];
  S6 cd[3];
  float tb[6], tl[3];  // in: A^-1 g ; out: the corrected solution
};
__device__ __noinline__ void woodbury_direction(const EnvShared &es, WoodburyIO &io, int part_all, int css, bool one_ss, int qbase, float ja) {
  const unsigned qm = 0xffffffffu;
  TreeMat H = io.H;
  float gb[6], gl[3], tl[3];
  S6 cd[3];
#pragma unroll
  for (int d = 0; d < 6; d++) gb[d] = io.gb[d];
#pragma unroll
  for (int j = 0; j < 3; j++) { gl[j] = io.gl[j]; tl[j] = io.tl[j]; cd[j] = io.cd[j]; }
  const ContactSlot &s = es.con[css];
  const int pc = (part_all >> (4 * css)) & 15;
  const int d1 = pc & 3, d2 = (pc >> 2) & 3, dep = d1 | d2;
  const float sg = one_ss ? ((d2 ? 1.f : 0.f) - (d1 ? 1.f : 0.f)) : 0.f;
  const V3 r = V3{s.r[0], s.r[1], s.r[2]};
  float cn[3], ct1[3], ct2[3];  // contact-frame components of this leg's columns
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const V3 col = cd[j].l + cross(cd[j].a, r);
    const float w = (dep != 0 && j <= dep) ? sg : 0.f;
    cn[j] = w * (s.frame[0] * col.x + s.frame[1] * col.y + s.frame[2] * col.z);
    ct1[j] = w * (s.frame[3] * col.x + s.frame[4] * col.y + s.frame[5] * col.z);
    ct2[j] = w * (s.frame[6] * col.x + s.frame[7] * col.y + s.frame[8] * col.z);
  }
  const float de = (one_ss && ja < 0.f) ? s.D : 0.f;  // weight of this lane's edge (0: inactive)
  float je[4][3], S[4][4], t[4], dinv[4];
  const float zb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const float dd = __shfl_sync(qm, de, qbase + e);
    const float on = dd > 0.f ? 1.f : 0.f;
    dinv[e] = dd > 0.f ? 1.f / dd : 1.f;
    const float em = ((e & 1) ? -s.mu : s.mu) * on;
#pragma unroll
    for (int j = 0; j < 3; j++) je[e][j] = fmaf(em, (e >> 1) ? ct2[j] : ct1[j], on * cn[j]);
    t[e] = qsum(je[e][0] * tl[0] + je[e][1] * tl[1] + je[e][2] * tl[2], qm);
  }
#pragma unroll 1
  for (int e = 0; e < 4; e++) {
    float ub[6], ul[3];
    float jr[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; q++) if (q == e) { jr[0] = je[q][0]; jr[1] = je[q][1]; jr[2] = je[q][2]; }
    tree_solve(H, zb, jr, ub, ul, qm);  // column e of A^-1 J^T
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float v = qsum(je[i][0] * ul[0] + je[i][1] * ul[1] + je[i][2] * ul[2], qm);
      if (i <= e) {
#pragma unroll
        for (int q = 0; q < 4; q++) if (q == e) S[q][i] = v;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; e++) S[e][e] += dinv[e];
  // 4x4 Cholesky of S (SPD: D^-1 > 0 on the diagonal, J A^-1 J^T >= 0) and the two triangular solves
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      float v = S[i][j];
#pragma unroll
      for (int q = 0; q < j; q++) v = fmaf(-S[i][q], S[j][q], v);
      S[i][j] = (i == j) ? sqrtf(v) : v / S[j][j];
    }
  float y[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float v = t[i];
#pragma unroll
    for (int q = 0; q < i; q++) v = fmaf(-S[i][q], y[q], v);
    y[i] = v / S[i][i];
  }
#pragma unroll
  for (int i = 3; i >= 0; i--) {
    float v = y[i];
#pragma unroll
    for (int q = i + 1; q < 4; q++) v = fmaf(-S[q][i], y[q], v);
    y[i] = v / S[i][i];
  }
  float g2[3];
#pragma unroll
  for (int j = 0; j < 3; j++) g2[j] = gl[j] - (y[0] * je[0][j] + y[1] * je[1][j] + y[2] * je[2][j] + y[3] * je[3][j]);
  float xb2[6], xl2[3];
  tree_solve(H, gb, g2, xb2, xl2, qm);
#pragma unroll
  for (int d = 0; d < 6; d++) io.tb[d] = xb2[d];
#pragma unroll
  for (int j = 0; j < 3; j++) io.tl[j] = xl2[j];
}

// ---------------------------------------------------------------------------------------------------
// One mjx.forward for the env of this quad.  Outputs qacc (ab, al).
// ---------------------------------------------------------------------------------------------------
template <bool DBG>
__device__ __forceinline__ void forward(const BlockShared &sh, EnvShared &es, float *rows, LaneState &L, int k, unsigned qm, int qbase,
                                     float ab[6], float al[3], bool want_stale, DbgOut *dbg, const int tid) {
  const PupperModelDesc &m = sh.m;
  const int b0 = 2 + 3 * k;   // first body of this leg
  const float dt = m.timestep;

  PHASE_SYNC_AT(0);
  // ---- kinematics (A.2) ------------------------------------------------------------------------
  Q4 q1 = qnormalize(Q4{L.qb[3], L.qb[4], L.qb[5], L.qb[6]});
  L.qb[3] = q1.w; L.qb[4] = q1.x; L.qb[5] = q1.y; L.qb[6] = q1.z;
  const V3 p1 = V3{L.qb[0], L.qb[1], L.qb[2]};
  const M3 R1 = qmat(q1);
  V3 pos[3], axis[3], xip[3];
  Q4 rot[3];
  float Iw[3][6];  // rotated body inertia (xx,yy,zz,xy,xz,yz)
  {
    V3 pp = p1;
    Q4 pq = q1;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int b = b0 + j;
      pos[j] = pp + KROT(V3{m.body_pos[b][0], m.body_pos[b][1], m.body_pos[b][2]}, pq);
      Q4 q = qmul(pq, Q4{m.body_quat[b][0], m.body_quat[b][1], m.body_quat[b][2], m.body_quat[b][3]});
      float sn, cs;
      sincos_small(L.ql[j] * 0.5f, &sn, &cs);
#if PUPPER_ZFOLD
      axis[j] = rotate_z(q);
      rot[j] = qmul_zrot(q, cs, sn);
#else
      axis[j] = KROT(V3{0.f, 0.f, 1.f}, q);
      rot[j] = qmul(q, Q4{cs, 0.f, 0.f, sn});
#endif
      xip[j] = pos[j] + KROT(V3{m.body_ipos[b][0], m.body_ipos[b][1], m.body_ipos[b][2]}, rot[j]);
      M3 Ri = qmat(qmul(rot[j], Q4{m.body_iquat[b][0], m.body_iquat[b][1], m.body_iquat[b][2], m.body_iquat[b][3]}));
      const float *di = &es.inertia[(b - 1) * 3];
      Iw[j][0] = Ri.m[0] * di[0] * Ri.m[0] + Ri.m[1] * di[1] * Ri.m[1] + Ri.m[2] * di[2] * Ri.m[2];
      Iw[j][1] = Ri.m[3] * di[0] * Ri.m[3] + Ri.m[4] * di[1] * Ri.m[4] + Ri.m[5] * di[2] * Ri.m[5];
      Iw[j][2] = Ri.m[6] * di[0] * Ri.m[6] + Ri.m[7] * di[1] * Ri.m[7] + Ri.m[8] * di[2] * Ri.m[8];
      Iw[j][3] = Ri.m[0] * di[0] * Ri.m[3] + Ri.m[1] * di[1] * Ri.m[4] + Ri.m[2] * di[2] * Ri.m[5];
      Iw[j][4] = Ri.m[0] * di[0] * Ri.m[6] + Ri.m[1] * di[1] * Ri.m[7] + Ri.m[2] * di[2] * Ri.m[8];
      Iw[j][5] = Ri.m[3] * di[0] * Ri.m[6] + Ri.m[4] * di[1] * Ri.m[7] + Ri.m[5] * di[2] * Ri.m[8];
      pp = pos[j];
      pq = rot[j];
    }
  }
  // collision sphere centres (knee on link2, foot on link3) and the foot site
  V3 sc[2];
  sc[0] = pos[1] + KROT(V3{m.sphere_pos[2 * k][0], m.sphere_pos[2 * k][1], m.sphere_pos[2 * k][2]}, rot[1]);
  sc[1] = pos[2] + KROT(V3{m.sphere_pos[2 * k + 1][0], m.sphere_pos[2 * k + 1][1], m.sphere_pos[2 * k + 1][2]}, rot[2]);
  es.sph[2 * k][0] = sc[0].x; es.sph[2 * k][1] = sc[0].y; es.sph[2 * k][2] = sc[0].z;
  es.sph[2 * k + 1][0] = sc[1].x; es.sph[2 * k + 1][1] = sc[1].y; es.sph[2 * k + 1][2] = sc[1].z;
  // base inertial frame
  const V3 xip_b = p1 + KROT(V3{es.ipos[0], es.ipos[1], es.ipos[2]}, q1);
  float Iwb[6];
  {
    M3 Ri = qmat(qmul(q1, Q4{m.body_iquat[1][0], m.body_iquat[1][1], m.body_iquat[1][2], m.body_iquat[1][3]}));
    const float *di = &es.inertia[0];
    Iwb[0] = Ri.m[0] * di[0] * Ri.m[0] + Ri.m[1] * di[1] * Ri.m[1] + Ri.m[2] * di[2] * Ri.m[2];
    Iwb[1] = Ri.m[3] * di[0] * Ri.m[3] + Ri.m[4] * di[1] * Ri.m[4] + Ri.m[5] * di[2] * Ri.m[5];
    Iwb[2] = Ri.m[6] * di[0] * Ri.m[6] + Ri.m[7] * di[1] * Ri.m[7] + Ri.m[8] * di[2] * Ri.m[8];
    Iwb[3] = Ri.m[0] * di[0] * Ri.m[3] + Ri.m[1] * di[1] * Ri.m[4] + Ri.m[2] * di[2] * Ri.m[5];
    Iwb[4] = Ri.m[0] * di[0] * Ri.m[6] + Ri.m[1] * di[1] * Ri.m[7] + Ri.m[2] * di[2] * Ri.m[8];
    Iwb[5] = Ri.m[3] * di[0] * Ri.m[6] + Ri.m[4] * di[1] * Ri.m[7] + Ri.m[5] * di[2] * Ri.m[8];
  }

  // ---- subtree COM (A.3): leaf -> root mass-weighted sum, butterfly over the 4 legs --------------
  const float mb = es.mass[0];
  float ml[3];
#pragma unroll
  for (int j = 0; j < 3; j++) ml[j] = es.mass[b0 - 1 + j];
  V3 C;
  {
    V3 pl = ml[1] * xip[1] + ml[2] * xip[2];
    pl = ml[0] * xip[0] + pl;
    float msum = ml[0] + (ml[1] + ml[2]);
    pl = qsum3(pl, qm);
    msum = qsum(msum, qm);
    V3 pt = mb * xip_b + pl;
    float mt = mb + msum;
    C = mt < kMinVal ? xip_b : V3{pt.x / mt, pt.y / mt, pt.z / mt};
  }

  PHASE_SYNC_AT(1);
  // ---- cinert, cdof (A.3) -------------------------------------------------------------------------
  Inertia ci[3], cib;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    V3 off = xip[j] - C;
    float mm = ml[j], o2 = dot(off, off);
    ci[j] = Inertia{Iw[j][0] + mm * (o2 - off.x * off.x), Iw[j][1] + mm * (o2 - off.y * off.y), Iw[j][2] + mm * (o2 - off.z * off.z),
                    Iw[j][3] - mm * off.x * off.y, Iw[j][4] - mm * off.x * off.z, Iw[j][5] - mm * off.y * off.z, mm * off, mm};
  }
  {
    V3 off = xip_b - C;
    float o2 = dot(off, off);
    cib = Inertia{Iwb[0] + mb * (o2 - off.x * off.x), Iwb[1] + mb * (o2 - off.y * off.y), Iwb[2] + mb * (o2 - off.z * off.z),
                  Iwb[3] - mb * off.x * off.y, Iwb[4] - mb * off.x * off.z, Iwb[5] - mb * off.y * off.z, mb * off, mb};
  }
  S6 cd[3];
#pragma unroll
  for (int j = 0; j < 3; j++) cd[j] = S6{axis[j], cross(axis[j], C - pos[j])};
  V3 ba[3], bo[3];
  {
    V3 ob = C - p1;
    ba[0] = V3{R1.m[0], R1.m[3], R1.m[6]};
    ba[1] = V3{R1.m[1], R1.m[4], R1.m[7]};
    ba[2] = V3{R1.m[2], R1.m[5], R1.m[8]};
#pragma unroll
    for (int i = 0; i < 3; i++) bo[i] = cross(ba[i], ob);
  }

  PHASE_SYNC_AT(2);
  // ---- velocities, RNE bias forces (A.7) -------------------------------------------------------------
  S6 cvb;  // base spatial velocity
  cvb.a = L.vb[3] * ba[0] + L.vb[4] * ba[1] + L.vb[5] * ba[2];
  cvb.l = V3{L.vb[0], L.vb[1], L.vb[2]} + L.vb[3] * bo[0] + L.vb[4] * bo[1] + L.vb[5] * bo[2];
  // cdof_dot of the 3 rotational base dofs uses the translational part only: [0, v_lin x a_i]
  S6 cab;  // base spatial acceleration bias
  {
    V3 vt = V3{L.vb[0], L.vb[1], L.vb[2]};
    V3 w = L.vb[3] * cross(vt, ba[0]) + L.vb[4] * cross(vt, ba[1]) + L.vb[5] * cross(vt, ba[2]);
    cab.a = V3{0.f, 0.f, 0.f};
    cab.l = V3{-m.gravity[0], -m.gravity[1], -m.gravity[2]} + w;
  }
  S6 cv[3], ca[3];
  {
    S6 pv = cvb, pa = cab;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      S6 cdd = motion_cross(pv, cd[j]);
      cv[j] = fma6(L.vl[j], cd[j], pv);
      ca[j] = fma6(L.vl[j], cdd, pa);
      pv = cv[j];
      pa = ca[j];
    }
  }
  if (want_stale) {  // park what the env level needs from this forward pass (before the solver, to free registers)
    if (k == 0) {
      V3 tv = cvb.l + cross(cvb.a, p1 - C);  // xd.vel = lin - off x ang
      float *t = es.st_torso;
      t[0] = p1.x; t[1] = p1.y; t[2] = p1.z; t[3] = q1.w; t[4] = q1.x; t[5] = q1.y; t[6] = q1.z;
      t[7] = cvb.a.x; t[8] = cvb.a.y; t[9] = cvb.a.z; t[10] = tv.x; t[11] = tv.y; t[12] = tv.z; t[13] = C.x; t[14] = C.y; t[15] = C.z;
    }
    V3 fs_ = pos[2] + rotate(V3{m.site_pos[1 + k][0], m.site_pos[1 + k][1], m.site_pos[1 + k][2]}, rot[2]);
    V3 lv_ = cv[2].l + cross(cv[2].a, pos[2] - C);
    float *q = es.st_leg[k];
    q[0] = fs_.x; q[1] = fs_.y; q[2] = fs_.z; q[3] = pos[2].x; q[4] = pos[2].y; q[5] = pos[2].z;
    q[6] = cv[2].a.x; q[7] = cv[2].a.y; q[8] = cv[2].a.z; q[9] = lv_.x; q[10] = lv_.y; q[11] = lv_.z;
    if (DBG && dbg) {
#pragma unroll
      for (int j = 0; j < 3; j++) { dbg->pos[j] = pos[j]; dbg->rot[j] = rot[j]; dbg->ang[j] = cv[j].a; dbg->vel[j] = cv[j].l + cross(cv[j].a, pos[j] - C); }
    }
  }
  float bias_l[3], bias_b[6];
  TreeMat M;
  {
    // backward pass: composite forces and composite inertias, leaf -> root
    S6 cf = S6{V3{0.f, 0.f, 0.f}, V3{0.f, 0.f, 0.f}};
    Inertia crb = Inertia{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, V3{0.f, 0.f, 0.f}, 0.f};
#pragma unroll
    for (int j = 2; j >= 0; j--) {
      S6 f = inert_mul(ci[j], ca[j]) + motion_cross_force(cv[j], inert_mul(ci[j], cv[j]));
#if PUPPER_ZFOLD
      cf = (j == 2) ? f : cf + f;
      crb = (j == 2) ? ci[j] : crb + ci[j];
#else
      cf = cf + f;
      crb = crb + ci[j];
#endif
      bias_l[j] = dot6(cd[j], cf);
      S6 F = inert_mul(crb, cd[j]);  // crb_cdof of dof j
#pragma unroll
      for (int jj = 0; jj <= j; jj++) M.D[tri(j, jj)] = dot6(F, cd[jj]);
      M.D[tri(j, j)] += m.dof_armature[6 + 3 * k + j];
      M.C[j][0] = F.l.x; M.C[j][1] = F.l.y; M.C[j][2] = F.l.z;
#pragma unroll
      for (int i = 0; i < 3; i++) M.C[j][3 + i] = dot(ba[i], F.a) + dot(bo[i], F.l);
    }
    // base: own body + the 4 leg subtrees
#if PUPPER_ZFOLD
    S6 fb = S6{cross(cib.h, cab.l), cib.m * cab.l} + motion_cross_force(cvb, inert_mul(cib, cvb));  // cab.a = 0
#else
    S6 fb = inert_mul(cib, cab) + motion_cross_force(cvb, inert_mul(cib, cvb));
#endif
    S6 cfb = fb + qsum6(cf, qm);
    Inertia crbb;
    crbb.xx = cib.xx + qsum(crb.xx, qm); crbb.yy = cib.yy + qsum(crb.yy, qm); crbb.zz = cib.zz + qsum(crb.zz, qm);
    crbb.xy = cib.xy + qsum(crb.xy, qm); crbb.xz = cib.xz + qsum(crb.xz, qm); crbb.yz = cib.yz + qsum(crb.yz, qm);
    crbb.h = cib.h + qsum3(crb.h, qm);
    crbb.m = cib.m + qsum(crb.m, qm);
    bias_b[0] = cfb.l.x; bias_b[1] = cfb.l.y; bias_b[2] = cfb.l.z;
#pragma unroll
    for (int i = 0; i < 3; i++) bias_b[3 + i] = dot(ba[i], cfb.a) + dot(bo[i], cfb.l);
    // base block of M: translation dofs cdof = [0, e_d]; rotation dofs cdof = [a_i, a_i x o]
#pragma unroll
    for (int i = 0; i < 21; i++) M.B[i] = 0.f;
    M.B[tri(0, 0)] = crbb.m + m.dof_armature[0];
    M.B[tri(1, 1)] = crbb.m + m.dof_armature[1];
    M.B[tri(2, 2)] = crbb.m + m.dof_armature[2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      S6 F = inert_mul(crbb, S6{ba[i], bo[i]});
      M.B[tri(3 + i, 0)] = F.l.x; M.B[tri(3 + i, 1)] = F.l.y; M.B[tri(3 + i, 2)] = F.l.z;
#pragma unroll
      for (int ii = 0; ii <= i; ii++) M.B[tri(3 + i, 3 + ii)] = dot(ba[ii], F.a) + dot(bo[ii], F.l);
      M.B[tri(3 + i, 3 + i)] += m.dof_armature[3 + i];
    }
  }

  // ---- actuation, smooth forces (A.7) ---------------------------------------------------------------
  float frc[3], fs_l[3], fs_b[6];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const int u = 3 * k + j;
    float f = es.kp * L.ctrl[j] + (-es.kp * L.ql[j] + -es.kd * L.vl[j]);
    f = fminf(fmaxf(f, m.act_forcerange[u][0]), m.act_forcerange[u][1]);
    frc[j] = f;
    fs_l[j] = -m.dof_damping[6 + u] * L.vl[j] - bias_l[j] + f;
  }
#pragma unroll
  for (int d = 0; d < 6; d++) fs_b[d] = -m.dof_damping[d] * L.vb[d] - bias_b[d];
  if (want_stale) { es.st_leg[k][12] = frc[0]; es.st_leg[k][13] = frc[1]; es.st_leg[k][14] = frc[2]; }

  PHASE_SYNC_AT(3);
  // ---- collision (A.5): keep only contacts that can act (dist < 0), at most max_contact_points ------
  __syncwarp(qm);  // sphere centres visible to the quad
  int ncon = 0;
  int n_ss = 0, css = 0;  // leg-leg contacts of this env: count and slot of the (last) one
  int ss_mask = 0;        // slots of this env that hold a leg-leg contact
  float knee_hits = 0.f, torso_hits = 0.f;
  int own_list = 0, own_count = 0;  // world-vs-leg contacts in which this lane's leg takes part (3 bits per entry)
  int part_all = 0;                 // participation code of this lane in every contact (4 bits per contact)
  bool plane_only = false;          // warp-uniform: the lists above were filled by the collision fast path
  {
    // Four candidates per lane: its own two spheres against the plane z = 0, the sphere-box pair ranked k by the broad
    // phase, the leg-leg sphere pair ranked k.  The last two are rare (no boxes on flat terrain, legs seldom touch) and
    // live out of line (box_candidate / ss_candidate), as does the ranking cut used when an env has more penetrating
    // candidates than contact slots (cut_candidates).
    Cand cd_[4];
#pragma unroll
    for (int i = 0; i < 2; i++) {
      float r = m.sphere_radius[2 * k + i];
      float d = sc[i].z - r;
      cd_[i].dist = d < 0.f ? d : kInf;
      cd_[i].n = V3{0.f, 0.f, 1.f};
      cd_[i].pos = sc[i] - (r + 0.5f * d) * cd_[i].n;
      cd_[i].code1 = -1; cd_[i].code2 = k * 4 + 1 + i; cd_[i].s1 = -1; cd_[i].s2 = 2 * k + i; cd_[i].box = 0;
    }
#pragma unroll
    for (int i = 2; i < 4; i++) {
      cd_[i].dist = kInf; cd_[i].pos = cd_[i].n = V3{0.f, 0.f, 0.f};
      cd_[i].code1 = cd_[i].code2 = cd_[i].s1 = cd_[i].s2 = -1; cd_[i].box = 0;
    }
    if (m.nbox > 0) {  // warp-uniform (a model property)
      Cand io;
      box_candidate(sh, es, k, qbase, sc[0], sc[1], io);
      cd_[2] = io;
    }
    // sphere-sphere: 24 leg-leg pairs; nothing to do unless one penetrates.  For that test the pairs are dealt to the lanes
    // by symmetry (own spheres against both spheres of the next leg, and half of the four pairs with the opposite leg), the
    // other legs' centres come by shuffle; the out-of-line path then ranks the pairs in MJX order.
    {
      const int n1 = qbase | ((k + 1) & 3), n2 = qbase | ((k + 2) & 3);
      const float rK = m.sphere_radius[2 * k], rF = m.sphere_radius[2 * k + 1];
      const V3 aK = V3{__shfl_sync(qm, sc[0].x, n1), __shfl_sync(qm, sc[0].y, n1), __shfl_sync(qm, sc[0].z, n1)};
      const V3 aF = V3{__shfl_sync(qm, sc[1].x, n1), __shfl_sync(qm, sc[1].y, n1), __shfl_sync(qm, sc[1].z, n1)};
      const V3 oK = V3{__shfl_sync(qm, sc[0].x, n2), __shfl_sync(qm, sc[0].y, n2), __shfl_sync(qm, sc[0].z, n2)};
      const V3 oF = V3{__shfl_sync(qm, sc[1].x, n2), __shfl_sync(qm, sc[1].y, n2), __shfl_sync(qm, sc[1].z, n2)};
      const float raK = __shfl_sync(qm, rK, n1), raF = __shfl_sync(qm, rF, n1), roK = __shfl_sync(qm, rK, n2), roF = __shfl_sync(qm, rF, n2);
      auto pen = [](V3 p, V3 q, float rr) { const V3 d = p - q; return dot(d, d) < rr * rr; };  // |d| - rr < 0 (rr > 0)
      bool any_neg = pen(sc[0], aK, rK + raK) | pen(sc[0], aF, rK + raF) | pen(sc[1], aK, rF + raK) | pen(sc[1], aF, rF + raF);
      // opposite leg: lanes 0,1 test (own knee, its knee) (own knee, its foot); lanes 2,3 (own knee, its foot) (own foot, its foot)
      const bool lowk = k < 2;
      any_neg |= pen(sc[0], lowk ? oK : oF, rK + (lowk ? roK : roF));
      any_neg |= pen(lowk ? sc[0] : sc[1], oF, (lowk ? rK : rF) + roF);
      if (__any_sync(qm, any_neg)) {  // warp-uniform
        Cand io;
        ss_candidate(sh, es, k, qbase, io);
        cd_[3] = io;
      }
    }
    // final cut: the max_contact_points smallest dist over [plane 0..7, box 8..11, sphere-sphere 12..15]
    const int maxc = m.max_contact_points;
    const bool a0 = cd_[0].dist < kInf, a1 = cd_[1].dist < kInf, a2 = cd_[2].dist < kInf, a3 = cd_[3].dist < kInf;
    const unsigned q0 = (__ballot_sync(qm, a0) >> qbase) & 15u, q1b = (__ballot_sync(qm, a1) >> qbase) & 15u;
    const unsigned q2 = (__ballot_sync(qm, a2) >> qbase) & 15u, q3 = (__ballot_sync(qm, a3) >> qbase) & 15u;
    const int nplane = __popc(q0) + __popc(q1b), total = nplane + __popc(q2) + __popc(q3);
    if (__all_sync(qm, total <= maxc)) {
      // Usual case: no env of the warp has more penetrating candidates than slots, so every one is kept and no ranking
      // is needed.  Slots are handed out in candidate order (the SET of contacts is what the solver sees; the order only
      // moves float32 rounding of sums over contacts).
      const unsigned below = (1u << k) - 1u;
      const int s0 = __popc(q0 & below) + __popc(q1b & below), s1 = s0 + (a0 ? 1 : 0);
      const int s2 = nplane + __popc(q2 & below), s3 = nplane + __popc(q2) + __popc(q3 & below);
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const bool on = i == 0 ? a0 : a1;
        const int slot = i == 0 ? s0 : s1;
        if (on) write_raw_slot(es.con[slot], cd_[i], C, 0);
      }
      if (__any_sync(qm, a2 | a3)) {  // rare: box or leg-leg candidates
        if (a2) write_raw_slot(es.con[s2], cd_[2], C, 1);
        if (a3) write_raw_slot(es.con[s3], cd_[3], C, 2);
      } else {  // plane contacts only: each lane knows its own contacts without looking at the slots
        plane_only = true;
        own_count = (a0 ? 1 : 0) + (a1 ? 1 : 0);
        own_list = a0 ? (s0 | (a1 ? s1 << 3 : 0)) : (a1 ? s1 : 0);
        part_all = (a0 ? 4 << (4 * s0) : 0) | (a1 ? 8 << (4 * s1) : 0);  // body2 at depth 1 (knee sphere) / depth 2 (foot sphere)
        const float kh = (a0 ? (float)((sh.c.knee_sphere_mask >> (2 * k)) & 1u) : 0.f) + (a1 ? (float)((sh.c.knee_sphere_mask >> (2 * k + 1)) & 1u) : 0.f);
        const float th = (a0 ? (float)((sh.c.torso_sphere_mask >> (2 * k)) & 1u) : 0.f) + (a1 ? (float)((sh.c.torso_sphere_mask >> (2 * k + 1)) & 1u) : 0.f);
        knee_hits = qsum(kh, qm); torso_hits = qsum(th, qm);
      }
      ncon = total;
    } else {
      CutIO io;
#pragma unroll
      for (int i = 0; i < 4; i++) io.c[i] = cd_[i];
      ncon = cut_candidates(es, k, qbase, maxc, C, io);
    }
  }
  __syncwarp(qm);  // raw contact slots visible to the quad
  const int ncon_w = __reduce_max_sync(qm, ncon);
  // Complete the slots in parallel: lane k derives frame, friction, impedance and row weight of slot k (then k + 4),
  // instead of the winner lane doing it alone inside the selection loop.
#pragma unroll 1
  for (int c0 = 0; c0 < ncon_w; c0 += 4) {  // warp-uniform trip count (<= 2)
    const int c = c0 + k;
    if (c < ncon) {
      ContactSlot &s = es.con[c];
      const int c1 = s.code1, c2 = s.code2, s1 = s.s1, s2 = s.s2, ty = s.ty, bxi = s.box;
      const float bk = s.dist;
      make_frame(V3{s.frame[0], s.frame[1], s.frame[2]}, s.frame);
      // geom friction: one DR draw for every geom (es.friction >= 0) or the model's per-geom values
      const bool drf = es.friction >= 0.f;
      float f1 = s1 >= 0 ? (drf ? es.friction : m.sphere_friction[s1]) : (drf ? es.friction : m.floor_friction);
      float f2 = s2 >= 0 ? (drf ? es.friction : m.sphere_friction[s2]) : (ty == 1 ? (drf ? es.friction : m.box_friction[bxi]) : (drf ? es.friction : m.floor_friction));
      if (ty == 0) f1 = drf ? es.friction : m.floor_friction;
      float mu = fmaxf(f1, f2);
      float w1 = c1 >= 0 ? m.body_invweight0[2 + 3 * (c1 >> 2) + (c1 & 3)] : 0.f;
      float w2 = c2 >= 0 ? m.body_invweight0[2 + 3 * (c2 >> 2) + (c2 & 3)] : 0.f;
      float t = w1 + w2;
      float invw = t + mu * mu * t;
      invw = invw * 2.f * mu * mu / m.impratio;
      const float *solref = ty == 0 ? m.plane_sphere_solref : (ty == 1 ? m.sphere_box_solref : m.sphere_sphere_solref);
      const float *solimp = ty == 0 ? m.plane_sphere_solimp : (ty == 1 ? m.sphere_box_solimp : m.sphere_sphere_solimp);
      float kk, bb, imp;
      kbi(solref, solimp, dt, bk, kk, bb, imp);
      float Rr = fmaxf(invw * (1.f - imp) / imp, kMinVal);
      s.mu = mu; s.D = 1.f / Rr; s.b = bb; s.kimp = kk * imp * bk;
    }
  }
  __syncwarp(qm);  // completed contact slots visible to the quad
  if (!plane_only) {  // rare (boxes, leg-leg contacts, ranking cut): derive this lane's lists from the slots, out of line
    SlotScan sc_;
    scan_slots(sh.c.knee_sphere_mask, sh.c.torso_sphere_mask, es, ncon, k, sc_);
    n_ss = sc_.n_ss; css = sc_.css; ss_mask = sc_.ss_mask; part_all = sc_.part_all; own_list = sc_.own_list; own_count = sc_.own_count;
    knee_hits = sc_.knee_hits; torso_hits = sc_.torso_hits;
  }

  // A leg-leg contact couples two legs, so its rows do not fit the arrow structure.  One such contact (the common
  // rare case) is applied to the arrow solve as a rank-<=4 update (Woodbury, below); two or more go the dense way.
  const bool one_ss = n_ss == 1, dense_env = n_ss >= 2;
  const int ss_mask_w = plane_only ? 0 : (int)__reduce_or_sync(qm, (unsigned)ss_mask);  // slots holding a leg-leg contact somewhere in the warp
  const int nown_w = __reduce_max_sync(qm, own_count);  // most world-vs-leg contacts on one leg, over the warp
  if (want_stale && k == 0) { es.st_hits[0] = knee_hits; es.st_hits[1] = torso_hits; }
  PHASE_SYNC_AT(4);
  // ---- constraint rows handled by this lane (A.6): 3 friction-loss, 3 limits, one pyramid edge per contact
  // (contact-edge row scalars live in shared memory: row[buffer][contact][thread])
  float *rowA = rows + tid, *rowB = rowA + kMaxCon * kBlock, *rowC = rowB + kMaxCon * kBlock;
  float *rowQ = rows + (tid & ~3);  // the same buffers seen from lane 0 of the quad: edge e of contact c at rowQ[c*kBlock + e]
  const float esgn = (k & 1) ? -1.f : 1.f;  // pyramid edge of this lane: Jn + esgn*mu*Jt[k>>1]
  const bool et2 = (k >> 1) != 0;
  float fl[3], rff[3], fD[3], fA[3];   // friction-loss rows: loss, R*loss, D, aref
  float lD[3], lA[3], lsign[3];        // limit rows: D, aref, Jacobian entry (+-1, 0 when inactive)
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const int d = 6 + 3 * k + j;
    fl[j] = sh.d.fric_loss[d];
    fD[j] = sh.d.fric_D[d];
    rff[j] = sh.d.fric_rf[d];
    fA[j] = -sh.d.fric_b * L.vl[j];
    float dlo = L.ql[j] - m.jnt_range[3 * k + j][0], dhi = m.jnt_range[3 * k + j][1] - L.ql[j];
    float p = fminf(dlo, dhi);
    lsign[j] = 0.f; lD[j] = 0.f; lA[j] = 0.f;
    if (p < 0.f) {  // joint past a limit (rare): efc_D and aref of its row, out of line
      lsign[j] = dlo < dhi ? 1.f : -1.f;
      const float2 r = limit_row(m.jnt_solref, m.jnt_solimp, dt, p, lsign[j] * L.vl[j], m.dof_invweight0[d]);
      lD[j] = r.x; lA[j] = r.y;
    }
  }

  // ---- unconstrained acceleration: M qacc_smooth = qfrc_smooth ---------------------------------------------
  float sb[6], sl[3];
  {
    TreeMat F = M;
    tree_factor(F, nullptr, qm);
    tree_solve(F, fs_b, fs_l, sb, sl, qm);
  }

  PHASE_SYNC_AT(5);
  // ---- Newton solver, one iteration (A.8) ------------------------------------------------------------------
  // cost at the warm start and at qacc_smooth.  (M qacc_smooth is taken as qfrc_smooth.)
  float Maw_b[6], Maw_l[3];
  tree_matvec(M, L.wb, L.wl, Maw_b, Maw_l, qm);
  float cw_con = 0.f, cs_con = 0.f;  // constraint cost of this lane's contact rows at the warm start / at qacc_smooth
  {
    float v3b[3][6], v3l[3][3];
#pragma unroll
    for (int d = 0; d < 6; d++) { v3b[0][d] = L.vb[d]; v3b[1][d] = L.wb[d]; v3b[2][d] = sb[d]; }
#pragma unroll
    for (int j = 0; j < 3; j++) { v3l[0][j] = L.vl[j]; v3l[1][j] = L.wl[j]; v3l[2][j] = sl[j]; }
    // contact rows at qvel / warm start / qacc_smooth, their reference accelerations and both start costs: world-vs-leg
    // contacts by the touching leg's lane (all four edges), leg-leg contacts edge by edge with a sum over the quad
    contact_rows_own<3, true, DBG>(es, cd, ba, bo, v3b, v3l, own_list, own_count, nown_w, part_all, rowQ, cw_con, cs_con);
    if (ss_mask_w) {  // leg-leg contact somewhere in the warp (rare): out of line, arguments through memory
      LegLegIO<3> io;
#pragma unroll
      for (int j = 0; j < 3; j++) { io.cd[j] = cd[j]; io.ba[j] = ba[j]; io.bo[j] = bo[j]; }
#pragma unroll
      for (int n = 0; n < 3; n++) {
#pragma unroll
        for (int d = 0; d < 6; d++) io.vb[n][d] = v3b[n][d];
#pragma unroll
        for (int j = 0; j < 3; j++) io.vl[n][j] = v3l[n][j];
      }
      legleg_rows3(es, io, esgn, et2, part_all, ss_mask, ss_mask_w, rowA);
    }
  }
  float cost_w, cost_s, gauss_w;
  float jaw_f[3], jaw_l[3], jas_f[3], jas_l[3];
  {
    float cw = cw_con, cs = cs_con;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      // friction-loss row (Huber): 0.5 D x^2 inside |x| < R f, f (|x| - R f / 2) outside; a row without friction loss
      // (fl = 0) has R f = 0 and contributes f * (...) = 0.  Branch-free on purpose (branches are costly here).
      {
        const float xw = L.wl[j] - fA[j], xs = sl[j] - fA[j];
        jaw_f[j] = xw; jas_f[j] = xs;
        const float aw = fabsf(xw), as = fabsf(xs), hr = 0.5f * rff[j];
        cw += aw >= rff[j] ? fl[j] * (aw - hr) : 0.5f * fD[j] * xw * xw;
        cs += as >= rff[j] ? fl[j] * (as - hr) : 0.5f * fD[j] * xs * xs;
      }
      {
        const float xw = lsign[j] * L.wl[j] - lA[j], xs = lsign[j] * sl[j] - lA[j];
        jaw_l[j] = xw; jas_l[j] = xs;
        const float mw = fminf(xw, 0.f), ms = fminf(xs, 0.f);
        cw = fmaf(0.5f * lD[j] * mw, mw, cw);
        cs = fmaf(0.5f * lD[j] * ms, ms, cs);
      }
    }
    if (ss_mask_w) {  // leg-leg slots hold raw rows: turn this lane's edge into Jaref and add its cost (rare, out of line)
      const float2 dc = legleg_start_cost(es, ss_mask, rowA);
      cw += dc.x; cs += dc.y;
    }
    __syncwarp(qm);  // Jaref rows written by the owning lanes are visible to every edge's lane
    if (DBG && want_stale && dbg) {  // constraint rows as the solver sees them (PupperStepOut.dbg_efc)
#pragma unroll
      for (int j = 0; j < 3; j++) { dbg->efc_fD[j] = fl[j] > 0.f ? fD[j] : 0.f; dbg->efc_fA[j] = fA[j]; dbg->efc_lD[j] = lD[j]; dbg->efc_lA[j] = lA[j]; }
#pragma unroll
      for (int c = 0; c < kMaxCon; c++) {
        const bool on = c < ncon;
        const bool is_ss = on && es.con[c].code1 >= 0 && es.con[c].code2 >= 0;
        dbg->efc_cD[c] = on ? es.con[c].D : 0.f;
        // own contacts: aref parked in buffer A by contact_rows_own; leg-leg slots: buffer A holds J.qvel of this lane's edge
        dbg->efc_cA[c] = on ? (is_ss ? -es.con[c].b * rowA[c * kBlock] - es.con[c].kimp : rowA[c * kBlock]) : 0.f;
      }
    }
    float gw = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) gw = fmaf(Maw_l[j] - fs_l[j], L.wl[j] - sl[j], gw);
    gw = qsum(gw, qm);
#pragma unroll
    for (int d = 0; d < 6; d++) gw = fmaf(Maw_b[d] - fs_b[d], L.wb[d] - sb[d], gw);
    gauss_w = 0.5f * gw;
    cost_w = qsum(cw, qm) + gauss_w;
    cost_s = qsum(cs, qm);  // gauss(qacc_smooth) = 0
  }
  const bool use_w = cost_w < cost_s;
  float *rowJ = use_w ? rowB : rowC;  // Jaref of the contact-edge rows at the chosen start
  float xb[6], xl[3], Mab[6], Mal[3];  // qacc, M qacc
  float fJ[3], lJ[3];
  const float gauss = use_w ? gauss_w : 0.f;
#pragma unroll
  for (int d = 0; d < 6; d++) { xb[d] = use_w ? L.wb[d] : sb[d]; Mab[d] = use_w ? Maw_b[d] : fs_b[d]; }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    xl[j] = use_w ? L.wl[j] : sl[j]; Mal[j] = use_w ? Maw_l[j] : fs_l[j];
    fJ[j] = use_w ? jaw_f[j] : jas_f[j];
    lJ[j] = use_w ? jaw_l[j] : jas_l[j];
  }

  PHASE_SYNC_AT(6);
  // forces, J^T f, gradient; Hessian additions
  float gb[6], gl[3];
  // H is built in place in M's registers; the copy of M in shared memory serves the line search's M*search
  {
    float *msm = const_cast<float *>(sh.mat) + tid;
#pragma unroll
    for (int i = 0; i < 21; i++) msm[i * kBlock] = M.B[i];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int d = 0; d < 6; d++) msm[(21 + j * 6 + d) * kBlock] = M.C[j][d];
#pragma unroll
    for (int i = 0; i < 6; i++) msm[(39 + i) * kBlock] = M.D[i];
  }
  TreeMat &H = M;
  float Badd[21];
#pragma unroll
  for (int i = 0; i < 21; i++) Badd[i] = 0.f;
  {
    float qc_l[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; j++) {
      {  // friction-loss row: force -D x inside the quadratic zone (and the row enters H), -+f outside
        const bool lin = fabsf(fJ[j]) >= rff[j];
        qc_l[j] += lin ? -copysignf(fl[j], fJ[j]) : -fD[j] * fJ[j];
        H.D[tri(j, j)] += lin ? 0.f : fD[j];
      }
      {  // limit row: active when Jaref < 0
        const float ml = fminf(lJ[j], 0.f);
        qc_l[j] = fmaf(-lsign[j] * lD[j], ml, qc_l[j]);
        H.D[tri(j, j)] += lJ[j] < 0.f ? lD[j] : 0.f;
      }
    }
    S6 S1 = S6{V3{0.f, 0.f, 0.f}, V3{0.f, 0.f, 0.f}}, S2 = S1, Sb = S1;  // wrenches on link2 / link3 chains, base
    // World-vs-leg contacts, each handled by the lane whose leg touches (it has the chain's cdofs): the four pyramid-edge
    // forces -> contact-frame force -> world wrench about C on this leg's chain and on the base, and the Hessian block
    // Jc^T W Jc with the 3x3 contact-frame weight W of the active edges.  The 4 lanes of a quad work on different contacts
    // at the same time; the trip count is the warp-wide maximum of contacts per leg (typically 1-2).
#pragma unroll 1
    for (int i = 0; i < nown_w; i++) {
      if (i < own_count) {
        const int c = (own_list >> (3 * i)) & 7;
        const ContactSlot &s = es.con[c];
        const float *ja = rowJ - k + c * kBlock;  // the four edges' Jaref
        const float D = s.D, mu = s.mu;
        const float j0 = ja[0], j1 = ja[1], j2 = ja[2], j3 = ja[3];
        const float d0 = j0 < 0.f ? D : 0.f, d1 = j1 < 0.f ? D : 0.f, d2 = j2 < 0.f ? D : 0.f, d3 = j3 < 0.f ? D : 0.f;
        const float f0 = -d0 * j0, f1 = -d1 * j1, f2 = -d2 * j2, f3 = -d3 * j3;
        const float Fn = (f0 + f1) + (f2 + f3), Ft1 = mu * (f0 - f1), Ft2 = mu * (f2 - f3);
        const V3 g = V3{s.frame[0] * Fn + s.frame[3] * Ft1 + s.frame[6] * Ft2, s.frame[1] * Fn + s.frame[4] * Ft1 + s.frame[7] * Ft2,
                        s.frame[2] * Fn + s.frame[5] * Ft1 + s.frame[8] * Ft2};
        const V3 r = V3{s.r[0], s.r[1], s.r[2]};
        const int pc = (part_all >> (4 * c)) & 15;
        const int dep = (pc & 3) | ((pc >> 2) & 3);
        const float sg = (pc >> 2) ? 1.f : -1.f;  // + as body2, - as body1
        const S6 w = S6{sg * cross(r, g), sg * g};
        if (dep == 1) S1 = S1 + w; else S2 = S2 + w;
        Sb = Sb + w;
        const float W00 = (d0 + d1) + (d2 + d3);
        if (!dense_env && W00 > 0.f) {
          const float W01 = mu * (d0 - d1), W02 = mu * (d2 - d3), W11 = mu * mu * (d0 + d1), W22 = mu * mu * (d2 + d3);
          float Jc[9][3], T[9][3];
#pragma unroll
          for (int d = 0; d < 9; d++) {
            V3 col;
            if (d < 3) col = V3{d == 0 ? 1.f : 0.f, d == 1 ? 1.f : 0.f, d == 2 ? 1.f : 0.f};
            else if (d < 6) col = bo[d - 3] + cross(ba[d - 3], r);
            else col = (d - 6 <= dep) ? cd[d - 6].l + cross(cd[d - 6].a, r) : V3{0.f, 0.f, 0.f};
            Jc[d][0] = s.frame[0] * col.x + s.frame[1] * col.y + s.frame[2] * col.z;
            Jc[d][1] = s.frame[3] * col.x + s.frame[4] * col.y + s.frame[5] * col.z;
            Jc[d][2] = s.frame[6] * col.x + s.frame[7] * col.y + s.frame[8] * col.z;
#if !PUPPER_F2
            T[d][0] = W00 * Jc[d][0] + W01 * Jc[d][1] + W02 * Jc[d][2];
            T[d][1] = W01 * Jc[d][0] + W11 * Jc[d][1];
            T[d][2] = W02 * Jc[d][0] + W22 * Jc[d][2];
#endif
          }
#if PUPPER_F2
          // The same products on register pairs: the base part of each contact-frame row of Jc (dofs 0-5) is three pairs,
          // every update is "broadcast scalar x pair + pair" (FFMA2), so the block costs about two thirds of the scalar form.
          float2 J0[3], J1[3], J2[3], T0[3], T1[3], T2[3];  // rows normal / tangent 1 / tangent 2, pairs (0,1) (2,3) (4,5)
#pragma unroll
          for (int q = 0; q < 3; q++) {
            J0[q] = make_float2(Jc[2 * q][0], Jc[2 * q + 1][0]);
            J1[q] = make_float2(Jc[2 * q][1], Jc[2 * q + 1][1]);
            J2[q] = make_float2(Jc[2 * q][2], Jc[2 * q + 1][2]);
            T0[q] = __ffma2_rn(make_float2(W02, W02), J2[q], __ffma2_rn(make_float2(W01, W01), J1[q], __fmul2_rn(make_float2(W00, W00), J0[q])));
            T1[q] = __ffma2_rn(make_float2(W11, W11), J1[q], __fmul2_rn(make_float2(W01, W01), J0[q]));
            T2[q] = __ffma2_rn(make_float2(W22, W22), J2[q], __fmul2_rn(make_float2(W02, W02), J0[q]));
          }
#pragma unroll
          for (int d = 6; d < 9; d++) {
            T[d][0] = W00 * Jc[d][0] + W01 * Jc[d][1] + W02 * Jc[d][2];
            T[d][1] = W01 * Jc[d][0] + W11 * Jc[d][1];
            T[d][2] = W02 * Jc[d][0] + W22 * Jc[d][2];
          }
#pragma unroll
          for (int i2 = 0; i2 < 6; i2++) {
            const float2 t0p = T0[i2 >> 1], t1p = T1[i2 >> 1], t2p = T2[i2 >> 1];
            const float t0 = (i2 & 1) ? t0p.y : t0p.x, t1 = (i2 & 1) ? t1p.y : t1p.x, t2 = (i2 & 1) ? t2p.y : t2p.x;
#pragma unroll
            for (int jp = 0; jp <= i2; jp += 2) {
              const float2 v = __ffma2_rn(make_float2(t2, t2), J2[jp >> 1], __ffma2_rn(make_float2(t1, t1), J1[jp >> 1], __fmul2_rn(make_float2(t0, t0), J0[jp >> 1])));
              Badd[tri(i2, jp)] += v.x;
              if (jp + 1 <= i2) Badd[tri(i2, jp + 1)] += v.y;
            }
          }
#pragma unroll
          for (int j = 0; j < 3; j++) {
            R6 c = row6(H.C[j]);
            const float2 a0 = make_float2(T[6 + j][0], T[6 + j][0]), a1 = make_float2(T[6 + j][1], T[6 + j][1]), a2 = make_float2(T[6 + j][2], T[6 + j][2]);
            c.p0 = __ffma2_rn(a2, J2[0], __ffma2_rn(a1, J1[0], __ffma2_rn(a0, J0[0], c.p0)));
            c.p1 = __ffma2_rn(a2, J2[1], __ffma2_rn(a1, J1[1], __ffma2_rn(a0, J0[1], c.p1)));
            c.p2 = __ffma2_rn(a2, J2[2], __ffma2_rn(a1, J1[2], __ffma2_rn(a0, J0[2], c.p2)));
            unrow6(c, H.C[j]);
#pragma unroll
            for (int jj = 0; jj <= j; jj++) H.D[tri(j, jj)] += T[6 + j][0] * Jc[6 + jj][0] + T[6 + j][1] * Jc[6 + jj][1] + T[6 + j][2] * Jc[6 + jj][2];
          }
#else
#pragma unroll
          for (int i2 = 0; i2 < 6; i2++)
#pragma unroll
            for (int j = 0; j <= i2; j++) Badd[tri(i2, j)] += T[i2][0] * Jc[j][0] + T[i2][1] * Jc[j][1] + T[i2][2] * Jc[j][2];
#pragma unroll
          for (int j = 0; j < 3; j++) {
#pragma unroll
            for (int d = 0; d < 6; d++) H.C[j][d] += T[6 + j][0] * Jc[d][0] + T[6 + j][1] * Jc[d][1] + T[6 + j][2] * Jc[d][2];
#pragma unroll
            for (int jj = 0; jj <= j; jj++) H.D[tri(j, jj)] += T[6 + j][0] * Jc[6 + jj][0] + T[6 + j][1] * Jc[6 + jj][1] + T[6 + j][2] * Jc[6 + jj][2];
          }
#endif
        }
      }
    }
    Sb = qsum6(Sb, qm);  // the base carries every world-vs-leg contact
    if (ss_mask_w) {
      // Leg-leg contacts: equal and opposite wrenches on the two legs' chains, nothing on the base (its columns cancel).
      // Their Hessian rows couple two legs: Woodbury / dense paths below.  Rare, out of line.
      S6 io[2] = {S1, S2};
      legleg_wrench(es, ss_mask, rowJ - k, part_all, io);
      S1 = io[0]; S2 = io[1];
    }
    // qfrc_constraint and gradient
    gl[0] = Mal[0] - fs_l[0] - (qc_l[0] + dot6(cd[0], S1 + S2));
    gl[1] = Mal[1] - fs_l[1] - (qc_l[1] + dot6(cd[1], S1 + S2));
    gl[2] = Mal[2] - fs_l[2] - (qc_l[2] + dot6(cd[2], S2));
    gb[0] = Mab[0] - fs_b[0] - Sb.l.x; gb[1] = Mab[1] - fs_b[1] - Sb.l.y; gb[2] = Mab[2] - fs_b[2] - Sb.l.z;
#pragma unroll
    for (int i = 0; i < 3; i++) gb[3 + i] = Mab[3 + i] - fs_b[3 + i] - (dot(ba[i], Sb.a) + dot(bo[i], Sb.l));
  }

  PHASE_SYNC_AT(7);
  // Newton direction: search = -H^-1 grad
  float hb[6], hl[3];
  if (__any_sync(qm, dense_env)) {  // very rare: some env of this warp has two or more leg-leg contacts (H holds M + diagonal terms there)
    DenseIO io;
    io.H = H;
#pragma unroll
    for (int d = 0; d < 6; d++) { io.gb[d] = gb[d]; io.hb[d] = 0.f; }
#pragma unroll
    for (int j = 0; j < 3; j++) { io.gl[j] = gl[j]; io.hl[j] = 0.f; io.cd[j] = cd[j]; io.ba[j] = ba[j]; io.bo[j] = bo[j]; }
    dense_newton_direction(es, ncon, k, dense_env, qbase, io, rowJ);
#pragma unroll
    for (int d = 0; d < 6; d++) hb[d] = io.hb[d];
#pragma unroll
    for (int j = 0; j < 3; j++) hl[j] = io.hl[j];
  }
  tree_factor(H, Badd, qm);
  {
    float tb[6], tl[3];
    tree_solve(H, gb, gl, tb, tl, qm);
    if (__any_sync(qm, one_ss)) {  // rare: one leg-leg contact in some env of the warp -> rank-<=4 Woodbury correction, out of line
      WoodburyIO io;
      io.H = H;
#pragma unroll
      for (int d = 0; d < 6; d++) { io.gb[d] = gb[d]; io.tb[d] = tb[d]; }
#pragma unroll
      for (int j = 0; j < 3; j++) { io.gl[j] = gl[j]; io.tl[j] = tl[j]; io.cd[j] = cd[j]; }
      woodbury_direction(es, io, part_all, css, one_ss, qbase, rowJ[css * kBlock]);
      if (one_ss) {
#pragma unroll
        for (int d = 0; d < 6; d++) tb[d] = io.tb[d];
#pragma unroll
        for (int j = 0; j < 3; j++) tl[j] = io.tl[j];
      }
    }
    if (!dense_env) {
#pragma unroll
      for (int d = 0; d < 6; d++) hb[d] = tb[d];
#pragma unroll
      for (int j = 0; j < 3; j++) hl[j] = tl[j];
    }
  }
#pragma unroll
  for (int d = 0; d < 6; d++) hb[d] = -hb[d];
#pragma unroll
  for (int j = 0; j < 3; j++) hl[j] = -hl[j];

  PHASE_SYNC_AT(8);
  // ---- line search along `search` (A.8.3) ---------------------------------------------------------------------
  float alpha;
  {
    float mvb[6], mvl[3];
    tree_matvec_smem(sh.mat + tid, hb, hl, mvb, mvl, qm);
    {
      float v1b[1][6], v1l[1][3];
#pragma unroll
      for (int d = 0; d < 6; d++) v1b[0][d] = hb[d];
#pragma unroll
      for (int j = 0; j < 3; j++) v1l[0][j] = hl[j];
      float unused0 = 0.f, unused1 = 0.f;
      contact_rows_own<1, false>(es, cd, ba, bo, v1b, v1l, own_list, own_count, nown_w, part_all, rowQ, unused0, unused1);  // jv of the contact-edge rows
      if (ss_mask_w) {
        LegLegIO<1> io;
#pragma unroll
        for (int j = 0; j < 3; j++) { io.cd[j] = cd[j]; io.ba[j] = ba[j]; io.bo[j] = bo[j]; io.vl[0][j] = hl[j]; }
#pragma unroll
        for (int d = 0; d < 6; d++) io.vb[0][d] = hb[d];
        legleg_rows1(es, io, esgn, et2, part_all, ss_mask, ss_mask_w, rowA);
      }
      __syncwarp(qm);
    }
    float sn = hl[0] * hl[0] + hl[1] * hl[1] + hl[2] * hl[2];
    float q1l = 0.f, q2l = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      q1l = fmaf(hl[j], Mal[j] - fs_l[j], q1l);
      q2l = fmaf(hl[j], mvl[j], q2l);
    }
    sn = qsum(sn, qm); q1l = qsum(q1l, qm); q2l = qsum(q2l, qm);
#pragma unroll
    for (int d = 0; d < 6; d++) { sn = fmaf(hb[d], hb[d], sn); q1l = fmaf(hb[d], Mab[d] - fs_b[d], q1l); q2l = fmaf(hb[d], mvb[d], q2l); }
    const float gq0 = gauss, gq1 = q1l, gq2 = 0.5f * q2l;
    // Quadratic models of this lane's rows, built once per line search and kept in lane-private shared memory (the
    // stage loop then needs two 16-byte loads per friction row, one 16-byte + one 4-byte load per contact-edge row).
    // The cost model at alpha = 0 (MJX's p0) falls out of the same pass: same sums, same order as eval3 below.
    float4 *lsf = const_cast<float4 *>(sh.lsf) + tid;
    float4 *lsq = const_cast<float4 *>(sh.lsq) + tid;
    float *lsc = const_cast<float *>(sh.lsc) + tid;
    float fb0 = 0.f, fb1 = 0.f, fb2 = 0.f;  // friction rows: quadratic-zone coefficients, common to all step sizes
    float fc[3][4];                         // friction rows: linear-zone corrections (kept for the alpha = 0 pass only)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const float ja = fJ[j], jv = hl[j];
      const float qa = 0.5f * ja * ja * fD[j], qb = jv * ja * fD[j], qc = 0.5f * jv * jv * fD[j];
      const float lv = fl[j] * jv, hr = -0.5f * rff[j];
      fb0 += qa; fb1 += qb; fb2 += qc;
      // linear zones replace (qa, qb, qc) by (f(-R f/2 -+ ja), -+f jv, 0): the difference is added where they apply
      fc[j][0] = fl[j] * (hr - ja) - qa; fc[j][1] = -lv - qb; fc[j][2] = fl[j] * (hr + ja) - qa; fc[j][3] = lv - qb;
      lsf[(2 * j) * kBlock] = make_float4(ja, jv, rff[j], qc);
      lsf[(2 * j + 1) * kBlock] = make_float4(fc[j][0], fc[j][1], fc[j][2], fc[j][3]);
    }
    float z0 = fb0, z1 = fb1, z2 = fb2;  // sums at alpha = 0
#pragma unroll
    for (int j = 0; j < 3; j++) {
      {
        const float x = fJ[j];
        const bool neg = x <= -rff[j];
        const float lin = (neg || (x >= rff[j])) ? 1.f : 0.f;
        z0 = fmaf(lin, neg ? fc[j][0] : fc[j][2], z0);
        z1 = fmaf(lin, neg ? fc[j][1] : fc[j][3], z1);
        z2 = fmaf(-lin, 0.5f * hl[j] * hl[j] * fD[j], z2);
      }
    }
    const bool lim_w = __any_sync(qm, (lsign[0] != 0.f) | (lsign[1] != 0.f) | (lsign[2] != 0.f));
    if (lim_w) {
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const float ja = lJ[j], jv = lsign[j] * hl[j];
        const float on = ja < 0.f ? 1.f : 0.f;
        z0 = fmaf(on, 0.5f * ja * ja * lD[j], z0); z1 = fmaf(on, jv * ja * lD[j], z1); z2 = fmaf(on, 0.5f * jv * jv * lD[j], z2);
      }
    }
#pragma unroll
    for (int c = 0; c < kMaxCon; c++) {
      const bool con = c < ncon;
      const float ja = con ? rowJ[c * kBlock] : 0.f, jv = con ? rowA[c * kBlock] : 0.f, D = con ? es.con[c].D : 0.f;
      const float qa = 0.5f * ja * ja * D, qb = jv * ja * D, qc = 0.5f * jv * jv * D;
      lsq[c * kBlock] = make_float4(ja, jv, qa, qb);
      lsc[c * kBlock] = qc;
      const float on = ja < 0.f ? 1.f : 0.f;
      z0 = fmaf(on, qa, z0); z1 = fmaf(on, qb, z1); z2 = fmaf(on, qc, z2);
    }
    const float gtol = m.tolerance * m.ls_tolerance * (sqrtf(sn) * m.meaninertia * 18.f);

    // evaluates the 1-D cost model at three step sizes at once (one pass over this lane's rows)
    auto evalN = [&](auto npc, const float *a, LSPoint *out) {
      constexpr int NP = decltype(npc)::value;
      float s0[NP], s1[NP], s2[NP];
#pragma unroll
      for (int p = 0; p < NP; p++) { s0[p] = fb0; s1[p] = fb1; s2[p] = fb2; }
#pragma unroll
      for (int j = 0; j < 3; j++) {
        {
          const float4 a4 = lsf[(2 * j) * kBlock], c4 = lsf[(2 * j + 1) * kBlock];
          const float ja = a4.x, jv = a4.y, rf = a4.z, qc = a4.w;
#pragma unroll
          for (int p = 0; p < NP; p++) {
            const float x = fmaf(a[p], jv, ja);
            const bool neg = x <= -rf;
            const float lin = (neg || (x >= rf)) ? 1.f : 0.f;  // branch-free: 0/1 weight of the correction
            s0[p] = fmaf(lin, neg ? c4.x : c4.z, s0[p]);
            s1[p] = fmaf(lin, neg ? c4.y : c4.w, s1[p]);
            s2[p] = fmaf(-lin, qc, s2[p]);
          }
        }
      }
      if (lim_w) {  // some joint of this warp is past a limit (rare); rows that are not have lD = 0, ja = jv = 0: they add zeros
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const float ja = lJ[j], jv = lsign[j] * hl[j];
          const float qa = 0.5f * ja * ja * lD[j], qb = jv * ja * lD[j], qc = 0.5f * jv * jv * lD[j];
#pragma unroll
          for (int p = 0; p < NP; p++) {
            const float on = fmaf(a[p], jv, ja) < 0.f ? 1.f : 0.f;
            s0[p] = fmaf(on, qa, s0[p]); s1[p] = fmaf(on, qb, s1[p]); s2[p] = fmaf(on, qc, s2[p]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kMaxCon; c++) {  // unrolled, zero coefficients past the env's contacts: no loop or divergence branches
        const float4 q4 = lsq[c * kBlock];
        const float ja = q4.x, jv = q4.y, qa = q4.z, qb = q4.w, qc = lsc[c * kBlock];
#pragma unroll
        for (int p = 0; p < NP; p++) {
          const float on = fmaf(a[p], jv, ja) < 0.f ? 1.f : 0.f;
          s0[p] = fmaf(on, qa, s0[p]); s1[p] = fmaf(on, qb, s1[p]); s2[p] = fmaf(on, qc, s2[p]);
        }
      }
#pragma unroll
      for (int p = 0; p < NP; p++) {
        float t0 = gq0 + qsum(s0[p], qm), t1 = gq1 + qsum(s1[p], qm), t2 = gq2 + qsum(s2[p], qm);
        out[p].alpha = a[p];
        out[p].cost = a[p] * a[p] * t2 + a[p] * t1 + t0;
        out[p].d0 = 2.f * a[p] * t2 + t1;
        out[p].d1 = 2.f * t2 + (t2 == 0.f ? kMinVal : 0.f);
      }
    };
    LSPoint p0, lo, hi;
    {
      const float t0 = gq0 + qsum(z0, qm), t1 = gq1 + qsum(z1, qm), t2 = gq2 + qsum(z2, qm);
      p0 = LSPoint{0.f, t0, t1, 2.f * t2 + (t2 == 0.f ? kMinVal : 0.f)};
    }
    lo = p0; hi = p0;
    bool swap = true, ls_on = true;
    int ls_it = 0;  // DBG: bracket refinements this env ran (MJX's loop counter)
    const int nstage = 1 + m.ls_iterations;  // stage 0: the Newton point of p0; stages 1..: bracket refinements
    {
      const float a1[1] = {p0.alpha - p0.d0 / p0.d1};
      LSPoint pt1[1];
      evalN(std::integral_constant<int, 1>{}, a1, pt1);
      lo = pt1[0];
      if (lo.d0 < p0.d0) { hi = p0; } else { hi = lo; lo = p0; }
    }
#pragma unroll 1
    for (int stage = 1; stage < nstage; stage++) {
      float a3[3];
      if (!swap || ((lo.d0 < 0.f) && (lo.d0 > -gtol)) || ((hi.d0 > 0.f) && (hi.d0 < gtol))) ls_on = false;
      if (!__any_sync(qm, ls_on)) break;  // warp-uniform exit; finished envs idle through the remaining stages
      a3[0] = lo.alpha - lo.d0 / lo.d1; a3[1] = hi.alpha - hi.d0 / hi.d1; a3[2] = 0.5f * (lo.alpha + hi.alpha);
      LSPoint pt[3];
      evalN(std::integral_constant<int, 3>{}, a3, pt);
      if (DBG && ls_on) ls_it++;
      if (ls_on) {
        const LSPoint lon = pt[0], hin = pt[1], mid = pt[2];
        const bool s1 = in_bracket(lo, lon); lo = ls_select(s1, lon, lo);
        const bool s2 = in_bracket(lo, mid); lo = ls_select(s2, mid, lo);
        const bool s3 = in_bracket(lo, hin); lo = ls_select(s3, hin, lo);
        const bool t1 = in_bracket(hi, hin); hi = ls_select(t1, hin, hi);
        const bool t2 = in_bracket(hi, mid); hi = ls_select(t2, mid, hi);
        const bool t3 = in_bracket(hi, lon); hi = ls_select(t3, lon, hi);
        swap = s1 | s2 | s3 | t1 | t2 | t3;
      }
    }
    bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
    alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
    if (!improved) alpha = 0.f;
    if (DBG && want_stale && dbg) {  // solver decisions (PupperStepOut.dbg_solver)
      int fz = 0, lim = 0, con = 0;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        fz |= (fJ[j] <= -rff[j] ? 2 : (fJ[j] >= rff[j] ? 3 : 1)) << (2 * (3 * k + j));
        lim |= ((lsign[j] != 0.f && lJ[j] < 0.f) ? 1 : 0) << (3 * k + j);
      }
      for (int c = 0; c < ncon; c++) con |= (rowJ[c * kBlock] < 0.f ? 1 : 0) << (4 * c + k);
#pragma unroll
      for (int sft = 1; sft <= 2; sft <<= 1) {
        fz |= __shfl_xor_sync(qm, fz, sft); lim |= __shfl_xor_sync(qm, lim, sft); con |= __shfl_xor_sync(qm, con, sft);
      }
      dbg->solver[0] = use_w ? 1 : 0; dbg->solver[1] = ls_it; dbg->solver[2] = ncon; dbg->solver[3] = fz; dbg->solver[4] = lim;
      dbg->solver[5] = con; dbg->solver[6] = __float_as_int(alpha); dbg->solver[7] = n_ss;
    }
  }
#pragma unroll
  for (int d = 0; d < 6; d++) ab[d] = fmaf(alpha, hb[d], xb[d]);
#pragma unroll
  for (int j = 0; j < 3; j++) al[j] = fmaf(alpha, hl[j], xl[j]);

  if (DBG && want_stale && dbg) {
#pragma unroll
    for (int j = 0; j < 3; j++) dbg->qacc_l[j] = al[j];
#pragma unroll
    for (int d = 0; d < 6; d++) dbg->qacc_b[d] = ab[d];
  }
  es.ncon = ncon;
}

}  // namespace pupper

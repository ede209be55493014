// pupper_math.cuh -- small device math used by the env-step kernel (quaternions, spatial vectors,
// threefry).  Conventions follow what the reference's path computes through brax.math / mjx math
// (SURVEY.md A.10): quaternions (w,x,y,z); spatial 6-vectors [angular(3), linear(3)].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pupper {

struct V3 {
  float x, y, z;
};
struct Q4 {
  float w, x, y, z;
};
struct S6 {  // spatial vector: angular a, linear l
  V3 a, l;
};

#ifndef PUPPER_F2
#define PUPPER_F2 1  // packed FP32x2 (FFMA2 / FADD2 / FMUL2, sm_100) for 6-vector arithmetic: same IEEE results per element, half the
#endif               // instructions to fetch -- the step kernel is instruction-fetch bound in its straight-line phases
#ifndef PUPPER_F2_V3
#define PUPPER_F2_V3 0  // ... and for the (x, y) halves of 3-vector sums / scalings: the register pairing costs more MOVs than it saves (static count +48)
#endif
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
#if PUPPER_F2_V3
__device__ __forceinline__ V3 operator+(V3 a, V3 b) {
  const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  return V3{r.x, r.y, a.z + b.z};
}
__device__ __forceinline__ V3 operator-(V3 a, V3 b) {
  const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
  return V3{r.x, r.y, a.z - b.z};
}
__device__ __forceinline__ V3 operator*(float s, V3 a) {
  const float2 r = __fmul2_rn(make_float2(s, s), make_float2(a.x, a.y));
  return V3{r.x, r.y, s * a.z};
}
__device__ __forceinline__ V3 operator*(V3 a, float s) { return s * a; }
#else
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{s * a.x, s * a.y, s * a.z}; }
#endif
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ V3 fma3(float s, V3 a, V3 b) {  // s*a + b
#if PUPPER_F2_V3
  const float2 r = __ffma2_rn(make_float2(s, s), make_float2(a.x, a.y), make_float2(b.x, b.y));
  return V3{r.x, r.y, fmaf(s, a.z, b.z)};
#else
  return V3{fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)};
#endif
}
__device__ __forceinline__ float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

#if PUPPER_F2
// a 6-vector as three register pairs: (a.x, a.y) (a.z, l.x) (l.y, l.z)
struct P6 { float2 p0, p1, p2; };
__device__ __forceinline__ P6 pack6(S6 v) { return P6{make_float2(v.a.x, v.a.y), make_float2(v.a.z, v.l.x), make_float2(v.l.y, v.l.z)}; }
__device__ __forceinline__ S6 unpack6(P6 p) { return S6{V3{p.p0.x, p.p0.y, p.p1.x}, V3{p.p1.y, p.p2.x, p.p2.y}}; }
__device__ __forceinline__ S6 operator+(S6 a, S6 b) {
  const P6 x = pack6(a), y = pack6(b);
  return unpack6(P6{__fadd2_rn(x.p0, y.p0), __fadd2_rn(x.p1, y.p1), __fadd2_rn(x.p2, y.p2)});
}
__device__ __forceinline__ S6 operator*(float s, S6 a) {
  const P6 x = pack6(a);
  const float2 ss = make_float2(s, s);
  return unpack6(P6{__fmul2_rn(ss, x.p0), __fmul2_rn(ss, x.p1), __fmul2_rn(ss, x.p2)});
}
__device__ __forceinline__ S6 fma6(float s, S6 a, S6 b) {
  const P6 x = pack6(a), y = pack6(b);
  const float2 ss = make_float2(s, s);
  return unpack6(P6{__ffma2_rn(ss, x.p0, y.p0), __ffma2_rn(ss, x.p1, y.p1), __ffma2_rn(ss, x.p2, y.p2)});
}
#else
__device__ __forceinline__ S6 operator+(S6 a, S6 b) { return S6{a.a + b.a, a.l + b.l}; }
__device__ __forceinline__ S6 operator*(float s, S6 a) { return S6{s * a.a, s * a.l}; }
__device__ __forceinline__ S6 fma6(float s, S6 a, S6 b) { return S6{fma3(s, a.a, b.a), fma3(s, a.l, b.l)}; }
#endif
#if PUPPER_F2
__device__ __forceinline__ float dot6(S6 a, S6 b) {
  const P6 x = pack6(a), y = pack6(b);
  const float2 t = __ffma2_rn(x.p2, y.p2, __ffma2_rn(x.p1, y.p1, __fmul2_rn(x.p0, y.p0)));
  return t.x + t.y;
}
#else
__device__ __forceinline__ float dot6(S6 a, S6 b) { return dot(a.a, b.a) + dot(a.l, b.l); }
#endif
// motion cross product u x v  = [u.a x v.a, u.l x v.a + u.a x v.l]
__device__ __forceinline__ S6 motion_cross(S6 u, S6 v) { return S6{cross(u.a, v.a), cross(u.l, v.a) + cross(u.a, v.l)}; }
// force cross product v x* f = [v.a x f.a + v.l x f.l, v.a x f.l]
__device__ __forceinline__ S6 motion_cross_force(S6 v, S6 f) { return S6{cross(v.a, f.a) + cross(v.l, f.l), cross(v.a, f.l)}; }

// spatial inertia about the subtree COM: [Ixx,Iyy,Izz,Ixy,Ixz,Iyz], h = m*off, m
struct Inertia {
  float xx, yy, zz, xy, xz, yz;
  V3 h;
  float m;
};
__device__ __forceinline__ Inertia operator+(Inertia a, Inertia b) {
  return Inertia{a.xx + b.xx, a.yy + b.yy, a.zz + b.zz, a.xy + b.xy, a.xz + b.xz, a.yz + b.yz, a.h + b.h, a.m + b.m};
}
__device__ __forceinline__ S6 inert_mul(const Inertia &I, S6 v) {
  V3 ia = V3{I.xx * v.a.x + I.xy * v.a.y + I.xz * v.a.z, I.xy * v.a.x + I.yy * v.a.y + I.yz * v.a.z,
             I.xz * v.a.x + I.yz * v.a.y + I.zz * v.a.z};
  return S6{ia + cross(I.h, v.l), I.m * v.l - cross(I.h, v.a)};
}

__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
  return Q4{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
__device__ __forceinline__ Q4 qinv(Q4 q) { return Q4{q.w, -q.x, -q.y, -q.z}; }
// rotate(v, q) = 2(u.v)u + (s^2 - u.u)v + 2s(u x v)
__device__ __forceinline__ V3 rotate(V3 v, Q4 q) {
  V3 u = V3{q.x, q.y, q.z};
  float uv = dot(u, v), uu = dot(u, u);
  V3 c = cross(u, v);
  float k = q.w * q.w - uu;
  return V3{2.f * (uv * u.x) + k * v.x + 2.f * q.w * c.x, 2.f * (uv * u.y) + k * v.y + 2.f * q.w * c.y,
            2.f * (uv * u.z) + k * v.z + 2.f * q.w * c.z};
}
// Same rotation as v + 2w (u x v) + 2 u x (u x v): 15 instructions instead of 26, equal to `rotate` up to float32 rounding
// (~1 ulp) for unit quaternions.  Used in the kinematics of the physics loop; the env level keeps brax's form above.
__device__ __forceinline__ V3 rotate_fast(V3 v, Q4 q) {
  const V3 u2 = V3{2.f * q.x, 2.f * q.y, 2.f * q.z};
  const V3 t = cross(u2, v);
  return V3{fmaf(q.y, t.z, fmaf(-q.z, t.y, fmaf(q.w, t.x, v.x))), fmaf(q.z, t.x, fmaf(-q.x, t.z, fmaf(q.w, t.y, v.y))),
            fmaf(q.x, t.y, fmaf(-q.y, t.x, fmaf(q.w, t.z, v.z)))};
}
// rotate((0,0,1), q) and q * (cs,0,0,sn) with the structural zeros dropped (same values for finite inputs)
__device__ __forceinline__ V3 rotate_z(Q4 q) {
  const float k = q.w * q.w - (q.x * q.x + q.y * q.y + q.z * q.z);
  return V3{2.f * (q.z * q.x) + 2.f * q.w * q.y, 2.f * (q.z * q.y) - 2.f * q.w * q.x, 2.f * (q.z * q.z) + k};
}
__device__ __forceinline__ Q4 qmul_zrot(Q4 a, float cs, float sn) {
  return Q4{a.w * cs - a.z * sn, a.x * cs + a.y * sn, a.y * cs - a.x * sn, a.w * sn + a.z * cs};
}
struct M3 {  // row-major 3x3
  float m[9];
};
__device__ __forceinline__ M3 qmat(Q4 q) {
  float w = q.w, x = q.x, y = q.y, z = q.z;
  M3 r;
  r.m[0] = w * w + x * x - y * y - z * z; r.m[1] = 2.f * (x * y - w * z); r.m[2] = 2.f * (x * z + w * y);
  r.m[3] = 2.f * (x * y + w * z); r.m[4] = w * w - x * x + y * y - z * z; r.m[5] = 2.f * (y * z - w * x);
  r.m[6] = 2.f * (x * z - w * y); r.m[7] = 2.f * (y * z + w * x); r.m[8] = w * w - x * x - y * y + z * z;
  return r;
}
// x / (|x| + 1e-6 (|x| == 0)); returns the norm
__device__ __forceinline__ float normalize3(V3 &v) {
  float n = sqrtf(dot(v, v));
  float d = n + (n == 0.f ? 1e-6f : 0.f);
  v = V3{v.x / d, v.y / d, v.z / d};
  return n;
}
__device__ __forceinline__ Q4 qnormalize(Q4 q) {
  float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  float d = n + (n == 0.f ? 1e-6f : 0.f);
  return Q4{q.w / d, q.x / d, q.y / d, q.z / d};
}
// brax math.safe_norm: 0 when every |x_i| <= 1e-8
__device__ __forceinline__ float brax_norm(V3 v) {
  if (fabsf(v.x) <= 1e-8f && fabsf(v.y) <= 1e-8f && fabsf(v.z) <= 1e-8f) return 0.f;
  return sqrtf(dot(v, v));
}

// Branch-free sin/cos for moderate arguments (|x| < ~1e3: joint half-angles, integration half-angles): three-term
// Cody-Waite reduction to [-pi/4, pi/4] and minimax polynomials, ~1 ulp.  The CUDA library version inlines a
// large-argument slow path at every call site, which bloats the straight-line hot path for nothing here.
__device__ __forceinline__ void sincos_small(float x, float *sn, float *cs) {
  const float k = rintf(x * 0.636619772f);
  float r = fmaf(k, -1.57079601e+00f, x);
  r = fmaf(k, -3.13916473e-07f, r);
  r = fmaf(k, -5.39030253e-15f, r);
  const float r2 = r * r;
  float s = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  s = fmaf(s, r2, -1.6666654611e-1f);
  s = fmaf(s * r2, r, r);
  float c = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  c = fmaf(c, r2, 4.166664568298827e-2f);
  c = fmaf(c, r2, -0.5f);
  c = fmaf(c, r2, 1.0f);
  const int q = (int)k;
  const bool swap = q & 1;
  const float ss = swap ? c : s, cc = swap ? s : c;
  *sn = (q & 2) ? -ss : ss;
  *cs = ((q + 1) & 2) ? -cc : cc;
}

// ---- jax 0.5.0 threefry2x32 (partitionable derivation), SURVEY.md A.11 ------------------------------
// Out of line on purpose: the env-level code calls it ~15 times per step, and inlined copies (~75 instructions each) only add
// code for the instruction caches to stream (the step is instruction-fetch bound).
__device__ __noinline__ uint2 threefry2x32(uint2 key, uint32_t c0, uint32_t c1) {
  const uint32_t ks0 = key.x, ks1 = key.y, ks2 = key.x ^ key.y ^ 0x1BD11BDAu;
  uint32_t x0 = c0 + ks0, x1 = c1 + ks1;
#define TF_R(r) x0 += x1; x1 = __funnelshift_l(x1, x1, r); x1 ^= x0;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)
  x0 += ks1; x1 += ks2 + 1u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24)
  x0 += ks2; x1 += ks0 + 2u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)
  x0 += ks0; x1 += ks1 + 3u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24)
  x0 += ks1; x1 += ks2 + 4u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)
  x0 += ks2; x1 += ks0 + 5u;
#undef TF_R
  return make_uint2(x0, x1);
}
__device__ __forceinline__ uint2 split_key(uint2 key, uint32_t i) { return threefry2x32(key, 0u, i); }
// element `i` of jax.random.uniform(key, (n,), lo, hi): three separately rounded float ops, no FMA
__device__ __forceinline__ float uniform(uint2 key, uint32_t i, float lo, float hi) {
  uint2 o = threefry2x32(key, 0u, i);
  float f = __fsub_rn(__uint_as_float(((o.x ^ o.y) >> 9) | 0x3f800000u), 1.0f);
  float v = __fadd_rn(__fmul_rn(f, __fsub_rn(hi, lo)), lo);
  return fmaxf(lo, v);
}

}  // namespace pupper

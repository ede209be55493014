/*
 * pupper_oracle.c -- CPU oracle of the batched PupperV3Env step.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates, one function per stage, what the reference's hot path computes:
 *   env level   : pupperv3_mjx/environment.py:314-543, rewards.py:9-138, utils.py:34-69,
 *                 domain_randomization.py:180-210 (file:line into /root/reference)
 *   physics [3P]: mujoco_mjx==3.2.7 (mjx/_src/{smooth,collision_driver,collision_primitive,
 *                 collision_convex,constraint,solver,passive,forward,support,math}.py),
 *                 brax==0.12.1 (brax/mjx/pipeline.py, brax/math.py, envs/wrappers/training.py),
 *                 jax==0.5.0 threefry PRNG -- un-vendored pins (requirements.txt:1-5); the published
 *                 algorithms are restated as specified in SURVEY.md Appendix A, dense and in MJX's
 *                 evaluation order (no sparsity tricks; those belong to the CUDA kernel under test).
 * PARITY STATUS: env level pinned against the reference executed under tests/refshim, physics [3P] unpinned against MJX -- see oracle.h.
 *
 * Compiled twice (REAL=double -> *_f64, REAL=float -> *_f32) with -ffp-contract=off.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle.h"

#ifdef ORACLE_F32
typedef float real;
#define SUF(name) name##_f32
#define R_SQRT sqrtf
#define R_SIN sinf
#define R_COS cosf
#define R_POW powf
#define R_EXP expf
#define R_ABS fabsf
#else
typedef double real;
#define SUF(name) name##_f64
#define R_SQRT sqrt
#define R_SIN sin
#define R_COS cos
#define R_POW pow
#define R_EXP exp
#define R_ABS fabs
#endif

#define NB PUPPER_NBODY
#define NV PUPPER_NV
#define NQ PUPPER_NQ
#define NU PUPPER_NU
#define MAXEFC ORACLE_MAX_EFC
#define MJ_MINVAL ((real)1e-15)
#define MJ_MINIMP ((real)1e-4)
#define MJ_MAXIMP ((real)0.9999)

static inline real r_min(real a, real b) { return a < b ? a : b; }
static inline real r_max(real a, real b) { return a > b ? a : b; }
static inline real r_clip(real x, real lo, real hi) { return r_min(r_max(x, lo), hi); }

/* ------------------------------------------------------------------------------------------
 * PRNG: jax 0.5.0 threefry2x32, partitionable derivation.  Always float32.
 * ---------------------------------------------------------------------------------------- */
#ifndef ORACLE_F32 /* shared helpers are emitted once (by the f64 translation unit) */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

void oracle_threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t *out) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
  for (int i = 0; i < 5; i++) {
    for (int j = 0; j < 4; j++) {
      x0 += x1;
      x1 = rotl32(x1, R[i & 1][j]);
      x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
  out[0] = x0;
  out[1] = x1;
}

/* jax.random.uniform element `index` of a draw with the given key: bits -> [0,1) -> affine -> max */
float oracle_uniform(uint32_t k0, uint32_t k1, uint32_t index, float lo, float hi) {
  uint32_t o[2];
  oracle_threefry2x32(k0, k1, 0u, index, o);
  uint32_t bits = ((o[0] ^ o[1]) >> 9) | 0x3f800000u;
  float f;
  memcpy(&f, &bits, 4);
  f = f - 1.0f;
  volatile float scale = hi - lo; /* volatile: keep the three roundings separate */
  volatile float prod = f * scale;
  float v = prod + lo;
  return v > lo ? v : lo;
}

/* index drawn by jax.random.choice(key, a, axis, p=p) (utils.py:67) */
int oracle_choice(uint32_t k0, uint32_t k1, const float *p, int n) {
  float cum[PUPPER_MAX_LAT];
  volatile float acc = 0.0f;
  for (int i = 0; i < n; i++) {
    acc = acc + p[i];
    cum[i] = acc;
  }
  float u = oracle_uniform(k0, k1, 0u, 0.0f, 1.0f);
  volatile float one_minus = 1.0f - u;
  float r = cum[n - 1] * one_minus;
  int idx = 0;
  for (int i = 0; i < n; i++) idx += (cum[i] < r);
  return idx;
}

int oracle_sizeof_env(void) { return (int)sizeof(OracleEnv); }
int oracle_sizeof_dr(void) { return (int)sizeof(OracleDR); }
int oracle_sizeof_debug(void) { return (int)sizeof(OracleDebug); }
#endif

static inline void split_key(const uint32_t key[2], uint32_t i, uint32_t out[2]) {
  oracle_threefry2x32(key[0], key[1], 0u, i, out);
}

/* External randoms (include/pupper_env.h PupperRand): when set, every draw of the env step / reset takes the raw
 * [0, 1) uniform of its row from this env's table instead of threefry, and gets jax.random.uniform's affine map. */
static __thread const float *g_ext = NULL;
static float du(int row, const uint32_t key[2], uint32_t index, float lo, float hi) {
  if (!g_ext) return oracle_uniform(key[0], key[1], index, lo, hi);
  float f = g_ext[row];
  volatile float scale = hi - lo;
  volatile float prod = f * scale;
  float v = prod + lo;
  return v > lo ? v : lo;
}
static int dchoice(int row, const uint32_t key[2], const float *p, int n) {
  if (!g_ext) return oracle_choice(key[0], key[1], p, n);
  float cum[PUPPER_MAX_LAT];
  volatile float acc = 0.0f;
  for (int i = 0; i < n; i++) { acc = acc + p[i]; cum[i] = acc; }
  float u = du(row, key, 0u, 0.0f, 1.0f);
  volatile float one_minus = 1.0f - u;
  float r = cum[n - 1] * one_minus;
  int idx = 0;
  for (int i = 0; i < n; i++) idx += (cum[i] < r);
  return idx;
}

/* ------------------------------------------------------------------------------------------
 * small math (MJX math.py / brax math.py conventions)
 * ---------------------------------------------------------------------------------------- */
static inline void cross3(const real a[3], const real b[3], real o[3]) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline real dot3(const real a[3], const real b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static inline void quat_mul(const real a[4], const real b[4], real o[4]) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
/* rotate(v, q) = 2(u.v)u + (s^2 - u.u)v + 2s(u x v) */
static inline void rotate(const real v[3], const real q[4], real o[3]) {
  const real s = q[0];
  const real *u = q + 1;
  real uv = dot3(u, v), uu = dot3(u, u), c[3];
  cross3(u, v, c);
  for (int i = 0; i < 3; i++) o[i] = 2 * (uv * u[i]) + (s * s - uu) * v[i] + 2 * s * c[i];
}
static inline void quat_inv(const real q[4], real o[4]) { o[0] = q[0]; o[1] = -q[1]; o[2] = -q[2]; o[3] = -q[3]; }
static inline void quat_to_mat(const real q[4], real m[9]) {
  real w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
/* x / (n + 1e-6 (n == 0)), returns n */
static inline real normalize_n(real *x, int n) {
  real s = 0;
  for (int i = 0; i < n; i++) s += x[i] * x[i];
  real nrm = R_SQRT(s);
  real d = nrm + (real)1e-6 * (nrm == 0 ? (real)1 : (real)0);
  for (int i = 0; i < n; i++) x[i] = x[i] / d;
  return nrm;
}
/* brax math.safe_norm: 0 if all |x_i| <= 1e-8 */
static inline real brax_norm3(const real x[3]) {
  if (R_ABS(x[0]) <= (real)1e-8 && R_ABS(x[1]) <= (real)1e-8 && R_ABS(x[2]) <= (real)1e-8) return 0;
  return R_SQRT(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
}
/* spatial vectors [ang(3), lin(3)]; cinert = [Ixx,Iyy,Izz,Ixy,Ixz,Iyz, m*off(3), m] */
static inline void inert_mul(const real I[10], const real v[6], real o[6]) {
  real a0 = I[0] * v[0] + I[3] * v[1] + I[4] * v[2];
  real a1 = I[3] * v[0] + I[1] * v[1] + I[5] * v[2];
  real a2 = I[4] * v[0] + I[5] * v[1] + I[2] * v[2];
  real c[3], c2[3];
  cross3(I + 6, v + 3, c);
  cross3(I + 6, v, c2);
  o[0] = a0 + c[0]; o[1] = a1 + c[1]; o[2] = a2 + c[2];
  o[3] = I[9] * v[3] - c2[0]; o[4] = I[9] * v[4] - c2[1]; o[5] = I[9] * v[5] - c2[2];
}
static inline void motion_cross(const real u[6], const real v[6], real o[6]) {
  real a[3], b[3], c[3];
  cross3(u, v, a);
  cross3(u + 3, v, b);
  cross3(u, v + 3, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
static inline void motion_cross_force(const real v[6], const real f[6], real o[6]) {
  real a[3], b[3], c[3];
  cross3(v, f, a);
  cross3(v + 3, f + 3, b);
  cross3(v, f + 3, c);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
  o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}

/* dense Cholesky (lower) and solve; returns 0 on success */
static int chol_factor(const real *A, real *L, int n) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      real s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
      if (i == j) {
        if (!(s > 0)) return -1;
        L[i * n + i] = R_SQRT(s);
      } else {
        L[i * n + j] = s / L[j * n + j];
      }
    }
  return 0;
}
static void chol_solve(const real *L, const real *b, real *x, int n) {
  real y[NV];
  for (int i = 0; i < n; i++) {
    real s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k];
    y[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    real s = y[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

/* ------------------------------------------------------------------------------------------
 * model in working precision (+ per-env DR leaves applied)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int parent[NB];
  real body_pos[NB][3], body_quat[NB][4], body_ipos[NB][3], body_iquat[NB][4], body_mass[NB], body_inertia[NB][3];
  real body_invweight0[NB];
  real armature[NV], damping[NV], frictionloss[NV], dof_invweight0[NV];
  real dof_solref[2], dof_solimp[5], jnt_range[NU][2], jnt_solref[2], jnt_solimp[5];
  real gain[NU], bias1[NU], bias2[NU], frc_lo[NU], frc_hi[NU];
  real floor_friction, sphere_friction[PUPPER_NSPHERE], box_friction[PUPPER_MAX_BOX];
  int sphere_body[PUPPER_NSPHERE], sphere_geomid[PUPPER_NSPHERE], floor_geomid, nbox, box_geomid[PUPPER_MAX_BOX];
  real sphere_pos[PUPPER_NSPHERE][3], sphere_radius[PUPPER_NSPHERE];
  real box_pos[PUPPER_MAX_BOX][3], box_mat[PUPPER_MAX_BOX][9], box_size[PUPPER_MAX_BOX][3];
  real ps_solref[2], ps_solimp[5], sb_solref[2], sb_solimp[5], ss_solref[2], ss_solimp[5];
  int site_body[PUPPER_NSITE];
  real site_pos[PUPPER_NSITE][3];
  real timestep, gravity[3], impratio, tolerance, ls_tolerance, meaninertia;
  int ls_iterations, max_geom_pairs, max_contact_points, frictionloss_rows;
} Model;

#define CPY(dst, src, n) for (int _i = 0; _i < (n); _i++) ((real *)(dst))[_i] = (real)((const float *)(src))[_i]

static void model_load(Model *M, const PupperModelDesc *d, const OracleDR *dr) {
  for (int b = 0; b < NB; b++) M->parent[b] = d->body_parent[b];
  CPY(M->body_pos, d->body_pos, NB * 3); CPY(M->body_quat, d->body_quat, NB * 4);
  CPY(M->body_ipos, d->body_ipos, NB * 3); CPY(M->body_iquat, d->body_iquat, NB * 4);
  CPY(M->body_mass, d->body_mass, NB); CPY(M->body_inertia, d->body_inertia, NB * 3);
  CPY(M->body_invweight0, d->body_invweight0, NB);
  CPY(M->armature, d->dof_armature, NV); CPY(M->damping, d->dof_damping, NV);
  CPY(M->frictionloss, d->dof_frictionloss, NV); CPY(M->dof_invweight0, d->dof_invweight0, NV);
  CPY(M->dof_solref, d->dof_solref, 2); CPY(M->dof_solimp, d->dof_solimp, 5);
  CPY(M->jnt_range, d->jnt_range, NU * 2); CPY(M->jnt_solref, d->jnt_solref, 2); CPY(M->jnt_solimp, d->jnt_solimp, 5);
  CPY(M->gain, d->act_gain, NU); CPY(M->bias1, d->act_bias1, NU); CPY(M->bias2, d->act_bias2, NU);
  for (int i = 0; i < NU; i++) { M->frc_lo[i] = (real)d->act_forcerange[i][0]; M->frc_hi[i] = (real)d->act_forcerange[i][1]; }
  M->floor_friction = (real)d->floor_friction;
  CPY(M->sphere_friction, d->sphere_friction, PUPPER_NSPHERE); CPY(M->box_friction, d->box_friction, PUPPER_MAX_BOX);
  for (int s = 0; s < PUPPER_NSPHERE; s++) { M->sphere_body[s] = d->sphere_body[s]; M->sphere_geomid[s] = d->sphere_geomid[s]; }
  M->floor_geomid = d->floor_geomid; M->nbox = d->nbox;
  for (int b = 0; b < PUPPER_MAX_BOX; b++) M->box_geomid[b] = d->box_geomid[b];
  CPY(M->sphere_pos, d->sphere_pos, PUPPER_NSPHERE * 3); CPY(M->sphere_radius, d->sphere_radius, PUPPER_NSPHERE);
  CPY(M->box_pos, d->box_pos, PUPPER_MAX_BOX * 3); CPY(M->box_mat, d->box_mat, PUPPER_MAX_BOX * 9);
  CPY(M->box_size, d->box_size, PUPPER_MAX_BOX * 3);
  CPY(M->ps_solref, d->plane_sphere_solref, 2); CPY(M->ps_solimp, d->plane_sphere_solimp, 5);
  CPY(M->sb_solref, d->sphere_box_solref, 2); CPY(M->sb_solimp, d->sphere_box_solimp, 5);
  CPY(M->ss_solref, d->sphere_sphere_solref, 2); CPY(M->ss_solimp, d->sphere_sphere_solimp, 5);
  for (int s = 0; s < PUPPER_NSITE; s++) M->site_body[s] = d->site_body[s];
  CPY(M->site_pos, d->site_pos, PUPPER_NSITE * 3);
  M->timestep = (real)d->timestep; CPY(M->gravity, d->gravity, 3);
  M->impratio = (real)d->impratio; M->tolerance = (real)d->tolerance; M->ls_tolerance = (real)d->ls_tolerance;
  M->meaninertia = (real)d->meaninertia;
  M->ls_iterations = d->ls_iterations; M->max_geom_pairs = d->max_geom_pairs;
  M->max_contact_points = d->max_contact_points; M->frictionloss_rows = d->frictionloss_rows;
  if (dr) { /* domain_randomization.py:94-110 (values are f32 in the reference) */
    real f = (real)(float)dr->friction;
    M->floor_friction = f;
    for (int s = 0; s < PUPPER_NSPHERE; s++) M->sphere_friction[s] = f;
    for (int b = 0; b < PUPPER_MAX_BOX; b++) M->box_friction[b] = f;
    for (int i = 0; i < NU; i++) { M->gain[i] = (real)(float)dr->kp; M->bias1[i] = -(real)(float)dr->kp; M->bias2[i] = -(real)(float)dr->kd; }
    for (int k = 0; k < 3; k++) M->body_ipos[1][k] = (real)(float)dr->base_ipos[k];
    for (int b = 1; b < NB; b++) {
      M->body_mass[b] = (real)(float)dr->body_mass[b - 1];
      for (int k = 0; k < 3; k++) M->body_inertia[b][k] = (real)(float)dr->body_inertia[(b - 1) * 3 + k];
    }
  }
}

static inline int dof_body(int d) { return d < 6 ? 1 : d - 4; }
static inline int dof_parent(int d) { /* MuJoCo dof_parentid */
  if (d < 6) return d - 1;
  int j = (d - 6) % 3;
  return j == 0 ? 5 : d - 1;
}

/* ------------------------------------------------------------------------------------------
 * mjx.forward, specialised
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  real xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], ximat[NB][9];
  real xanchor[NV][3], xaxis[NV][3]; /* per dof (free dofs share the base anchor) */
  real sphere_xpos[PUPPER_NSPHERE][3], site_xpos[PUPPER_NSITE][3];
  real com[3];
  real cinert[NB][10], cdof[NV][6], cdof_dot[NV][6], cvel[NB][6];
  real qM[NV * NV], qLD[NV * NV];
  real qfrc_passive[NV], qfrc_bias[NV], qfrc_actuator[NV], qfrc_smooth[NV], qacc_smooth[NV];
  int ncon;
  real con_dist[PUPPER_MAX_CON], con_pos[PUPPER_MAX_CON][3], con_frame[PUPPER_MAX_CON][9], con_mu[PUPPER_MAX_CON];
  int con_geom[PUPPER_MAX_CON][2], con_body[PUPPER_MAX_CON][2], con_type[PUPPER_MAX_CON];
  int nefc, nf;
  real efc_J[MAXEFC][NV], efc_D[MAXEFC], efc_aref[MAXEFC], efc_pos[MAXEFC], efc_floss[MAXEFC], efc_force[MAXEFC];
  real qacc[NV], qfrc_constraint[NV];
  real ls_alpha, cost_start, cost_end, warm_cost, smooth_cost;
  int ls_iters, used_warmstart;
  int zone0[MAXEFC]; /* row zones at the start point of the Newton iteration (0 inactive, 1 quadratic, 2 / 3 linear) */
} Data;

static void kinematics(const Model *M, real *qpos, Data *D) {
  for (int k = 0; k < 3; k++) D->xpos[0][k] = 0;
  D->xquat[0][0] = 1; D->xquat[0][1] = D->xquat[0][2] = D->xquat[0][3] = 0;
  /* free joint: pos = qpos[0:3], quat = normalize(qpos[3:7]) written back */
  normalize_n(qpos + 3, 4);
  for (int k = 0; k < 3; k++) D->xpos[1][k] = qpos[k];
  for (int k = 0; k < 4; k++) D->xquat[1][k] = qpos[3 + k];
  for (int d = 0; d < 6; d++)
    for (int k = 0; k < 3; k++) { D->xanchor[d][k] = qpos[k]; D->xaxis[d][k] = (k == 2); }
  for (int b = 2; b < NB; b++) {
    int p = M->parent[b], d = b + 4;
    real r[3], q[4], ql[4];
    rotate(M->body_pos[b], D->xquat[p], r);
    for (int k = 0; k < 3; k++) D->xpos[b][k] = D->xpos[p][k] + r[k];
    quat_mul(D->xquat[p], M->body_quat[b], q);
    const real zaxis[3] = {0, 0, 1};
    rotate(zaxis, q, D->xaxis[d]);
    for (int k = 0; k < 3; k++) D->xanchor[d][k] = D->xpos[b][k];
    real half = qpos[7 + b - 2] * (real)0.5;
    ql[0] = R_COS(half); ql[1] = 0; ql[2] = 0; ql[3] = R_SIN(half);
    quat_mul(q, ql, D->xquat[b]);
  }
  for (int b = 0; b < NB; b++) {
    real r[3], q[4];
    quat_to_mat(D->xquat[b], D->xmat[b]);
    rotate(M->body_ipos[b], D->xquat[b], r);
    for (int k = 0; k < 3; k++) D->xipos[b][k] = D->xpos[b][k] + r[k];
    quat_mul(D->xquat[b], M->body_iquat[b], q);
    quat_to_mat(q, D->ximat[b]);
  }
  for (int s = 0; s < PUPPER_NSPHERE; s++) {
    int b = M->sphere_body[s];
    real r[3];
    rotate(M->sphere_pos[s], D->xquat[b], r);
    for (int k = 0; k < 3; k++) D->sphere_xpos[s][k] = D->xpos[b][k] + r[k];
  }
  for (int s = 0; s < PUPPER_NSITE; s++) {
    int b = M->site_body[s];
    real r[3];
    rotate(M->site_pos[s], D->xquat[b], r);
    for (int k = 0; k < 3; k++) D->site_xpos[s][k] = D->xpos[b][k] + r[k];
  }
}

static void com_pos(const Model *M, Data *D) {
  /* subtree COM of the root body, accumulated leaf -> root: body + (sum over children) */
  real pos[NB][3], mass[NB];
  for (int b = 0; b < NB; b++) {
    mass[b] = M->body_mass[b];
    for (int k = 0; k < 3; k++) pos[b][k] = D->xipos[b][k] * M->body_mass[b];
  }
  for (int leg = 0; leg < 4; leg++) {
    int b1 = 2 + 3 * leg;
    for (int b = b1 + 1; b >= b1; b--) { /* link2 += link3, link1 += link2 */
      for (int k = 0; k < 3; k++) pos[b][k] += pos[b + 1][k];
      mass[b] += mass[b + 1];
    }
  }
  real cp[3] = {0, 0, 0}, cm = 0;
  for (int leg = 0; leg < 4; leg++) {
    for (int k = 0; k < 3; k++) cp[k] += pos[2 + 3 * leg][k];
    cm += mass[2 + 3 * leg];
  }
  for (int k = 0; k < 3; k++) pos[1][k] += cp[k];
  mass[1] += cm;
  for (int k = 0; k < 3; k++) D->com[k] = mass[1] < MJ_MINVAL ? D->xipos[1][k] : pos[1][k] / mass[1];

  for (int b = 0; b < NB; b++) {
    real off[3], I[9];
    const real *R = D->ximat[b], *di = M->body_inertia[b];
    real m = M->body_mass[b];
    for (int k = 0; k < 3; k++) off[k] = D->xipos[b][k] - (b == 0 ? D->xipos[0][k] : D->com[k]);
    /* (ximat * inertia) @ ximat.T + h h^T m with h = cross(off, -I3) */
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        real s = 0;
        for (int k = 0; k < 3; k++) s += (R[i * 3 + k] * di[k]) * R[j * 3 + k];
        I[i * 3 + j] = s;
      }
    real h[9] = {0, off[2], -off[1], -off[2], 0, off[0], off[1], -off[0], 0};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        real s = 0;
        for (int k = 0; k < 3; k++) s += h[i * 3 + k] * h[j * 3 + k];
        I[i * 3 + j] += s * m;
      }
    real *c = D->cinert[b];
    c[0] = I[0]; c[1] = I[4]; c[2] = I[8]; c[3] = I[1]; c[4] = I[2]; c[5] = I[5];
    c[6] = off[0] * m; c[7] = off[1] * m; c[8] = off[2] * m; c[9] = m;
  }
  for (int d = 0; d < NV; d++) {
    real o[3];
    for (int k = 0; k < 3; k++) o[k] = D->com[k] - D->xanchor[d][k];
    real *c = D->cdof[d];
    if (d < 3) {
      for (int k = 0; k < 6; k++) c[k] = (k == 3 + d);
    } else {
      real a[3];
      if (d < 6) { for (int k = 0; k < 3; k++) a[k] = D->xmat[1][k * 3 + (d - 3)]; }
      else { for (int k = 0; k < 3; k++) a[k] = D->xaxis[d][k]; }
      c[0] = a[0]; c[1] = a[1]; c[2] = a[2];
      cross3(a, o, c + 3);
    }
  }
}

static void crb_and_factor(const Model *M, Data *D) {
  real crb[NB][10];
  memcpy(crb, D->cinert, sizeof(crb));
  for (int b = NB - 1; b >= 2; b--) {
    int p = M->parent[b];
    for (int k = 0; k < 10; k++) crb[p][k] += crb[b][k];
  }
  real f[NV][6];
  for (int d = 0; d < NV; d++) inert_mul(crb[dof_body(d)], D->cdof[d], f[d]);
  memset(D->qM, 0, sizeof(D->qM));
  for (int i = 0; i < NV; i++) {
    for (int j = i; j >= 0; j = dof_parent(j)) {
      real s = 0;
      for (int k = 0; k < 6; k++) s += f[i][k] * D->cdof[j][k];
      if (i == j) s += M->armature[i];
      D->qM[i * NV + j] = s;
      D->qM[j * NV + i] = s;
    }
  }
  chol_factor(D->qM, D->qLD, NV);
}

/* ---- collision (A.5) ------------------------------------------------------------------- */
static void make_frame(const real n_in[3], real fr[9]) {
  real a[3] = {n_in[0], n_in[1], n_in[2]};
  normalize_n(a, 3);
  real b[3] = {0, 0, 0};
  if (R_ABS(a[1]) < (real)0.5) b[1] = 1; else b[2] = 1;
  real ab = dot3(a, b);
  for (int k = 0; k < 3; k++) b[k] -= a[k] * ab;
  normalize_n(b, 3);
  real c[3];
  cross3(a, b, c);
  for (int k = 0; k < 3; k++) { fr[k] = a[k]; fr[3 + k] = b[k]; fr[6 + k] = c[k]; }
}

typedef struct { real dist, pos[3], n[3], mu; int g1, g2, b1, b2, type; } Cand;

/* box as a 6-face polytope; vertex v = 4*ix+2*iy+iz over (-1,+1)^3, MJX face table */
static const int BOX_FACE[6][4] = {{0, 4, 5, 1}, {0, 2, 6, 4}, {6, 7, 5, 4}, {2, 3, 7, 6}, {1, 5, 7, 3}, {0, 1, 3, 2}};
static const real BOX_NORMAL[6][3] = {{0, -1, 0}, {0, 0, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}};

static void sphere_box(const real c_world[3], real radius, const real bpos[3], const real bmat[9],
                       const real size[3], real *dist_out, real pos_out[3], real n_out[3]) {
  real d[3], c[3];
  for (int k = 0; k < 3; k++) d[k] = c_world[k] - bpos[k];
  for (int i = 0; i < 3; i++) c[i] = bmat[0 * 3 + i] * d[0] + bmat[1 * 3 + i] * d[1] + bmat[2 * 3 + i] * d[2]; /* mat^T d */
  real vert[8][3];
  for (int v = 0; v < 8; v++) {
    vert[v][0] = ((v & 4) ? 1 : -1) * size[0];
    vert[v][1] = ((v & 2) ? 1 : -1) * size[1];
    vert[v][2] = ((v & 1) ? 1 : -1) * size[2];
  }
  int best = 0;
  real best_s = 0;
  for (int f = 0; f < 6; f++) {
    real p[3];
    for (int k = 0; k < 3; k++) p[k] = c[k] - BOX_NORMAL[f][k] * radius - vert[BOX_FACE[f][0]][k];
    real s = dot3(p, BOX_NORMAL[f]);
    if (s >= 0) s = (real)-1e12;
    if (f == 0 || s > best_s) { best = f; best_s = s; }
  }
  const real *fn = BOX_NORMAL[best];
  real face[4][3];
  for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) face[i][k] = vert[BOX_FACE[best][i]][k];
  real tmp[3], pt[3];
  for (int k = 0; k < 3; k++) tmp[k] = c[k] - face[0][k];
  real pd = dot3(tmp, fn);
  for (int k = 0; k < 3; k++) pt[k] = c[k] - pd * fn[k];
  int inside = 1, idx = 0;
  real best_e = 0;
  for (int i = 0; i < 4; i++) {
    const real *p0 = face[(i + 3) & 3], *p1 = face[i];
    real e[3], en[3], r[3];
    for (int k = 0; k < 3; k++) { e[k] = p1[k] - p0[k]; r[k] = pt[k] - p0[k]; }
    cross3(e, fn, en);
    real ed = dot3(r, en);
    if (!(ed <= 0)) inside = 0;
    int degenerate = (en[0] == 0 && en[1] == 0 && en[2] == 0);
    real v = (degenerate || ed < 0) ? (real)1e12 : ed;
    if (i == 0 || v < best_e) { best_e = v; idx = i; }
  }
  if (!inside) { /* closest point on the selected edge segment */
    const real *a = face[(idx + 3) & 3], *b = face[idx];
    real ab[3], pa[3];
    for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; pa[k] = pt[k] - a[k]; }
    real t = dot3(pa, ab) / (dot3(ab, ab) + (real)1e-6);
    t = r_clip(t, 0, 1);
    for (int k = 0; k < 3; k++) pt[k] = a[k] + t * ab[k];
  }
  real n[3];
  for (int k = 0; k < 3; k++) n[k] = pt[k] - c[k];
  real dn = normalize_n(n, 3);
  real pos[3];
  for (int k = 0; k < 3; k++) pos[k] = (pt[k] + (c[k] + n[k] * radius)) * (real)0.5;
  *dist_out = dn - radius;
  for (int i = 0; i < 3; i++) {
    n_out[i] = bmat[i * 3 + 0] * n[0] + bmat[i * 3 + 1] * n[1] + bmat[i * 3 + 2] * n[2];
    pos_out[i] = bmat[i * 3 + 0] * pos[0] + bmat[i * 3 + 1] * pos[1] + bmat[i * 3 + 2] * pos[2] + bpos[i];
  }
}

/* stable selection of the k smallest keys (ties -> lower index), ascending: lax.top_k(-key, k) */
static void top_k_smallest(const real *key, int n, int k, int *idx) {
  char used[8 * PUPPER_MAX_BOX + 64];
  memset(used, 0, sizeof(used));
  for (int j = 0; j < k; j++) {
    int best = -1;
    for (int i = 0; i < n; i++)
      if (!used[i] && (best < 0 || key[i] < key[best])) best = i;
    used[best] = 1;
    idx[j] = best;
  }
}

static void collision(const Model *M, Data *D) {
  Cand cand[8 + 2 * PUPPER_MAX_PAIRS + 32];
  int nc = 0;
  const int maxp = M->max_geom_pairs;
  /* group (PLANE, SPHERE): never culled */
  for (int s = 0; s < PUPPER_NSPHERE; s++) {
    Cand *c = &cand[nc++];
    const real *p = D->sphere_xpos[s];
    real r = M->sphere_radius[s];
    c->dist = p[2] - r; /* plane z = 0, normal +z */
    c->n[0] = 0; c->n[1] = 0; c->n[2] = 1;
    for (int k = 0; k < 3; k++) c->pos[k] = p[k] - c->n[k] * (r + (real)0.5 * c->dist);
    c->g1 = M->floor_geomid; c->g2 = M->sphere_geomid[s]; c->b1 = 0; c->b2 = M->sphere_body[s];
    c->mu = r_max(M->floor_friction, M->sphere_friction[s]);
    c->type = 0;
  }
  /* group (SPHERE, BOX): pair index = sphere * nbox + box */
  if (M->nbox > 0) {
    int npair = PUPPER_NSPHERE * M->nbox, sel[8 * PUPPER_MAX_BOX], nsel = npair;
    if (maxp > -1 && npair > maxp) {
      real key[8 * PUPPER_MAX_BOX];
      for (int s = 0; s < PUPPER_NSPHERE; s++)
        for (int b = 0; b < M->nbox; b++) {
          real d[3];
          for (int k = 0; k < 3; k++) d[k] = M->box_pos[b][k] - D->sphere_xpos[s][k];
          real rb = R_SQRT(M->box_size[b][0] * M->box_size[b][0] + M->box_size[b][1] * M->box_size[b][1] + M->box_size[b][2] * M->box_size[b][2]);
          key[s * M->nbox + b] = R_SQRT(dot3(d, d)) - (M->sphere_radius[s] + rb);
        }
      top_k_smallest(key, npair, maxp, sel);
      nsel = maxp;
    } else {
      for (int i = 0; i < npair; i++) sel[i] = i;
    }
    for (int i = 0; i < nsel; i++) {
      int s = sel[i] / M->nbox, b = sel[i] % M->nbox;
      Cand *c = &cand[nc++];
      sphere_box(D->sphere_xpos[s], M->sphere_radius[s], M->box_pos[b], M->box_mat[b], M->box_size[b], &c->dist, c->pos, c->n);
      c->g1 = M->sphere_geomid[s]; c->g2 = M->box_geomid[b]; c->b1 = M->sphere_body[s]; c->b2 = 0;
      c->mu = r_max(M->sphere_friction[s], M->box_friction[b]);
      c->type = 1;
    }
  }
  /* group (SPHERE, SPHERE): all sphere pairs on different legs */
  {
    int pa[28], pb[28], npair = 0, sel[28], nsel;
    real key[28];
    for (int a = 0; a < PUPPER_NSPHERE; a++)
      for (int b = a + 1; b < PUPPER_NSPHERE; b++) {
        if (a / 2 == b / 2) continue; /* knee/foot of one leg: parent-child bodies, filtered */
        real d[3];
        for (int k = 0; k < 3; k++) d[k] = D->sphere_xpos[b][k] - D->sphere_xpos[a][k];
        key[npair] = R_SQRT(dot3(d, d)) - (M->sphere_radius[a] + M->sphere_radius[b]);
        pa[npair] = a; pb[npair] = b; npair++;
      }
    nsel = npair;
    if (maxp > -1 && npair > maxp) { top_k_smallest(key, npair, maxp, sel); nsel = maxp; }
    else for (int i = 0; i < npair; i++) sel[i] = i;
    for (int i = 0; i < nsel; i++) {
      int a = pa[sel[i]], b = pb[sel[i]];
      Cand *c = &cand[nc++];
      real n[3];
      for (int k = 0; k < 3; k++) n[k] = D->sphere_xpos[b][k] - D->sphere_xpos[a][k];
      real dn = normalize_n(n, 3);
      if (dn == 0) { n[0] = 1; n[1] = 0; n[2] = 0; }
      c->dist = dn - (M->sphere_radius[a] + M->sphere_radius[b]);
      for (int k = 0; k < 3; k++) { c->n[k] = n[k]; c->pos[k] = D->sphere_xpos[a][k] + n[k] * (M->sphere_radius[a] + c->dist * (real)0.5); }
      c->g1 = M->sphere_geomid[a]; c->g2 = M->sphere_geomid[b]; c->b1 = M->sphere_body[a]; c->b2 = M->sphere_body[b];
      c->mu = r_max(M->sphere_friction[a], M->sphere_friction[b]);
      c->type = 2;
    }
  }
  /* one condim group: keep the max_contact_points smallest dist, ascending */
  int order[64], ncon = nc;
  if (M->max_contact_points > -1 && nc > M->max_contact_points) {
    real key[64];
    for (int i = 0; i < nc; i++) key[i] = cand[i].dist;
    ncon = M->max_contact_points;
    top_k_smallest(key, nc, ncon, order);
  } else {
    for (int i = 0; i < nc; i++) order[i] = i;
  }
  if (ncon > PUPPER_MAX_CON) ncon = PUPPER_MAX_CON;
  D->ncon = ncon;
  for (int i = 0; i < ncon; i++) {
    const Cand *c = &cand[order[i]];
    D->con_dist[i] = c->dist;
    for (int k = 0; k < 3; k++) D->con_pos[i][k] = c->pos[k];
    make_frame(c->n, D->con_frame[i]);
    D->con_mu[i] = c->mu;
    D->con_geom[i][0] = c->g1; D->con_geom[i][1] = c->g2;
    D->con_body[i][0] = c->b1; D->con_body[i][1] = c->b2;
    D->con_type[i] = c->type;
  }
}

/* ---- constraints (A.6) ------------------------------------------------------------------- */
static void kbi(const Model *M, const real solref[2], const real solimp[5], real pos, real *k, real *b, real *imp) {
  real timeconst = r_max(solref[0], 2 * M->timestep), dampratio = solref[1];
  real dmin = r_clip(solimp[0], MJ_MINIMP, MJ_MAXIMP), dmax = r_clip(solimp[1], MJ_MINIMP, MJ_MAXIMP);
  real width = r_max(MJ_MINVAL, solimp[2]), mid = r_clip(solimp[3], MJ_MINIMP, MJ_MAXIMP), power = r_max(1, solimp[4]);
  *k = 1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  *b = 2 / (dmax * timeconst);
  if (solref[0] <= 0) *k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0) *b = -solref[1] / dmax;
  real x = R_ABS(pos) / width;
  real ia = (1 / R_POW(mid, power - 1)) * R_POW(x, power);
  real ib = 1 - (1 / R_POW(1 - mid, power - 1)) * R_POW(1 - x, power);
  real y = x < mid ? ia : ib;
  real im = dmin + y * (dmax - dmin);
  im = r_clip(im, dmin, dmax);
  if (x > 1) im = dmax;
  *imp = im;
}

static void body_jacp(const Data *D, const real point[3], int body, real jac[NV][3]) {
  /* support.jac: jacp_k = cdof_k.lin + cdof_k.ang x (point - C), masked to the body's ancestor dofs */
  real off[3];
  for (int k = 0; k < 3; k++) off[k] = point[k] - D->com[k];
  for (int d = 0; d < NV; d++) {
    int on = 0;
    if (body >= 1) {
      if (d < 6) on = 1;
      else {
        int leg = (body - 2) / 3, depth = (body - 2) % 3; /* body>=2 */
        on = body >= 2 && (d - 6) / 3 == leg && (d - 6) % 3 <= depth;
      }
    }
    real c[3];
    cross3(D->cdof[d], off, c);
    for (int k = 0; k < 3; k++) jac[d][k] = on ? D->cdof[d][3 + k] + c[k] : 0;
  }
}

static void add_row(const Model *M, Data *D, const real *J, real pos, real invweight, const real solref[2],
                    const real solimp[5], real floss, const real *qvel) {
  int r = D->nefc++;
  real k, b, imp;
  kbi(M, solref, solimp, pos, &k, &b, &imp);
  real R = r_max(invweight * (1 - imp) / imp, MJ_MINVAL);
  real jv = 0;
  for (int d = 0; d < NV; d++) { D->efc_J[r][d] = J[d]; jv += J[d] * qvel[d]; }
  D->efc_aref[r] = -b * jv - k * imp * pos;
  D->efc_D[r] = 1 / R;
  D->efc_pos[r] = pos;
  D->efc_floss[r] = floss;
}

static void make_constraint(const Model *M, Data *D, const real *qpos, const real *qvel) {
  D->nefc = 0;
  D->nf = 0;
  real J[NV];
  if (M->frictionloss_rows) {
    for (int d = 6; d < NV; d++) {
      if (!(M->frictionloss[d] > 0)) continue;
      memset(J, 0, sizeof(J));
      J[d] = 1;
      add_row(M, D, J, 0, M->dof_invweight0[d], M->dof_solref, M->dof_solimp, M->frictionloss[d], qvel);
      D->nf++;
    }
  }
  for (int j = 0; j < NU; j++) {
    real q = qpos[7 + j];
    real dmin = q - M->jnt_range[j][0], dmax = M->jnt_range[j][1] - q;
    real pos = r_min(dmin, dmax);
    int active = pos < 0;
    memset(J, 0, sizeof(J));
    J[6 + j] = active ? (real)((dmin < dmax) * 2 - 1) : 0;
    add_row(M, D, J, active ? pos : 0, M->dof_invweight0[6 + j], M->jnt_solref, M->jnt_solimp, 0, qvel);
  }
  for (int c = 0; c < D->ncon; c++) {
    real j1[NV][3], j2[NV][3], Jc[3][NV];
    body_jacp(D, D->con_pos[c], D->con_body[c][0], j1);
    body_jacp(D, D->con_pos[c], D->con_body[c][1], j2);
    for (int i = 0; i < 3; i++)
      for (int d = 0; d < NV; d++) {
        real s = 0;
        for (int k = 0; k < 3; k++) s += D->con_frame[c][i * 3 + k] * (j2[d][k] - j1[d][k]);
        Jc[i][d] = s;
      }
    real mu = D->con_mu[c];
    real t = M->body_invweight0[D->con_body[c][0]] + M->body_invweight0[D->con_body[c][1]];
    real invweight = t + mu * mu * t;
    invweight = invweight * 2 * mu * mu / M->impratio;
    int active = D->con_dist[c] < 0;
    const real *solref = D->con_type[c] == 0 ? M->ps_solref : D->con_type[c] == 1 ? M->sb_solref : M->ss_solref;
    const real *solimp = D->con_type[c] == 0 ? M->ps_solimp : D->con_type[c] == 1 ? M->sb_solimp : M->ss_solimp;
    for (int e = 0; e < 4; e++) {
      real sgn = (e & 1) ? -mu : mu;
      const real *Jt = Jc[1 + (e >> 1)];
      for (int d = 0; d < NV; d++) J[d] = active ? Jc[0][d] + Jt[d] * sgn : 0;
      add_row(M, D, J, active ? D->con_dist[c] : 0, invweight, solref, solimp, 0, qvel);
    }
  }
}

/* ---- velocity stage, actuation, smooth acceleration (A.7) ------------------------------------ */
static void fwd_velocity(const Model *M, Data *D, const real *qvel) {
  memset(D->cvel[0], 0, sizeof(D->cvel[0]));
  { /* free joint */
    real v[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 6; k++) v[k] += D->cdof[0][k] * qvel[0] + D->cdof[1][k] * qvel[1] + D->cdof[2][k] * qvel[2];
    for (int d = 0; d < 3; d++) memset(D->cdof_dot[d], 0, sizeof(D->cdof_dot[d]));
    for (int d = 3; d < 6; d++) motion_cross(v, D->cdof[d], D->cdof_dot[d]);
    for (int k = 0; k < 6; k++) v[k] += D->cdof[3][k] * qvel[3] + D->cdof[4][k] * qvel[4] + D->cdof[5][k] * qvel[5];
    memcpy(D->cvel[1], v, sizeof(v));
  }
  for (int b = 2; b < NB; b++) {
    int d = b + 4, p = M->parent[b];
    motion_cross(D->cvel[p], D->cdof[d], D->cdof_dot[d]);
    for (int k = 0; k < 6; k++) D->cvel[b][k] = D->cvel[p][k] + D->cdof[d][k] * qvel[d];
  }
  for (int d = 0; d < NV; d++) D->qfrc_passive[d] = -M->damping[d] * qvel[d];
  /* RNE */
  real cacc[NB][6], cfrc[NB][6];
  for (int k = 0; k < 3; k++) { cacc[0][k] = 0; cacc[0][3 + k] = -M->gravity[k]; }
  for (int b = 1; b < NB; b++) {
    int p = M->parent[b];
    real s[6] = {0, 0, 0, 0, 0, 0};
    if (b == 1) { for (int d = 0; d < 6; d++) for (int k = 0; k < 6; k++) s[k] += D->cdof_dot[d][k] * qvel[d]; }
    else { for (int k = 0; k < 6; k++) s[k] = D->cdof_dot[b + 4][k] * qvel[b + 4]; }
    for (int k = 0; k < 6; k++) cacc[b][k] = cacc[p][k] + s[k];
  }
  for (int b = 0; b < NB; b++) {
    real a[6], iv[6], c[6];
    inert_mul(D->cinert[b], cacc[b], a);
    inert_mul(D->cinert[b], D->cvel[b], iv);
    motion_cross_force(D->cvel[b], iv, c);
    for (int k = 0; k < 6; k++) cfrc[b][k] = a[k] + c[k];
  }
  for (int b = NB - 1; b >= 1; b--) {
    int p = M->parent[b];
    for (int k = 0; k < 6; k++) cfrc[p][k] += cfrc[b][k];
  }
  for (int d = 0; d < NV; d++) {
    real s = 0;
    for (int k = 0; k < 6; k++) s += D->cdof[d][k] * cfrc[dof_body(d)][k];
    D->qfrc_bias[d] = s;
  }
}

static void fwd_actuation_acceleration(const Model *M, Data *D, const real *qpos, const real *qvel, const real *ctrl) {
  for (int d = 0; d < NV; d++) D->qfrc_actuator[d] = 0;
  for (int i = 0; i < NU; i++) {
    real force = M->gain[i] * ctrl[i] + (M->bias1[i] * qpos[7 + i] + M->bias2[i] * qvel[6 + i]);
    force = r_clip(force, M->frc_lo[i], M->frc_hi[i]);
    D->qfrc_actuator[6 + i] = force;
  }
  for (int d = 0; d < NV; d++) D->qfrc_smooth[d] = D->qfrc_passive[d] - D->qfrc_bias[d] + D->qfrc_actuator[d];
  chol_solve(D->qLD, D->qfrc_smooth, D->qacc_smooth, NV);
}

/* ---- Newton solver, one iteration (A.8) ---------------------------------------------------------- */
static void mul_m(const Data *D, const real *v, real *o) {
  for (int i = 0; i < NV; i++) {
    real s = 0;
    for (int j = 0; j < NV; j++) s += D->qM[i * NV + j] * v[j];
    o[i] = s;
  }
}

/* row state at residual x: zone 0 inactive, 1 quadratic, 2 friction linear-neg, 3 friction linear-pos */
static inline int row_zone(const Data *D, int r, real x) {
  if (r < D->nf) {
    real rf = D->efc_floss[r] / D->efc_D[r];
    if (x <= -rf) return 2;
    if (x >= rf) return 3;
    return 1;
  }
  return x < 0 ? 1 : 0;
}

typedef struct { real qacc[NV], Ma[NV], Jaref[MAXEFC], force[MAXEFC], qfrc_constraint[NV], gauss, cost; } Ctx;

static void ctx_init(const Data *D, const real *qacc, Ctx *c) {
  for (int i = 0; i < NV; i++) c->qacc[i] = qacc[i];
  for (int r = 0; r < D->nefc; r++) {
    real s = 0;
    for (int d = 0; d < NV; d++) s += D->efc_J[r][d] * qacc[d];
    c->Jaref[r] = s - D->efc_aref[r];
  }
  mul_m(D, qacc, c->Ma);
}

static void update_constraint(const Data *D, Ctx *c) {
  real cost = 0;
  for (int r = 0; r < D->nefc; r++) {
    real x = c->Jaref[r], Dr = D->efc_D[r], f = D->efc_floss[r];
    int z = row_zone(D, r, x);
    if (z == 1) { c->force[r] = -Dr * x; cost += (real)0.5 * Dr * x * x; }
    else if (z == 2) { c->force[r] = f; cost += f * ((real)-0.5 * (f / Dr) - x); }
    else if (z == 3) { c->force[r] = -f; cost += f * ((real)-0.5 * (f / Dr) + x); }
    else c->force[r] = 0;
  }
  for (int d = 0; d < NV; d++) {
    real s = 0;
    for (int r = 0; r < D->nefc; r++) s += D->efc_J[r][d] * c->force[r];
    c->qfrc_constraint[d] = s;
  }
  real g = 0;
  for (int d = 0; d < NV; d++) g += (c->Ma[d] - D->qfrc_smooth[d]) * (c->qacc[d] - D->qacc_smooth[d]);
  c->gauss = (real)0.5 * g;
  c->cost = cost + c->gauss;
}

typedef struct { real alpha, cost, d0, d1; } LSPoint;

static LSPoint ls_point(const Data *D, const Ctx *c, real alpha, const real *jv, const real quad_gauss[3]) {
  real q0 = quad_gauss[0], q1 = quad_gauss[1], q2 = quad_gauss[2];
  real s0 = 0, s1 = 0, s2 = 0;
  for (int r = 0; r < D->nefc; r++) {
    real Dr = D->efc_D[r], ja = c->Jaref[r], f = D->efc_floss[r];
    real x = ja + alpha * jv[r];
    int z = row_zone(D, r, x);
    if (z == 1) { s0 += (real)0.5 * ja * ja * Dr; s1 += jv[r] * ja * Dr; s2 += (real)0.5 * jv[r] * jv[r] * Dr; }
    else if (z == 2) { s0 += f * ((real)-0.5 * (f / Dr) - ja); s1 += -f * jv[r]; }
    else if (z == 3) { s0 += f * ((real)-0.5 * (f / Dr) + ja); s1 += f * jv[r]; }
  }
  q0 += s0; q1 += s1; q2 += s2;
  LSPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.d0 = 2 * alpha * q2 + q1;
  p.d1 = 2 * q2 + (q2 == 0 ? MJ_MINVAL : 0);
  return p;
}
static inline int in_bracket(LSPoint x, LSPoint y) {
  return ((x.d0 < y.d0) && (y.d0 < 0)) || ((x.d0 > y.d0) && (y.d0 > 0));
}

static void solve(const Model *M, Data *D, const real *qacc_warmstart) {
  Ctx warm, smth, *c;
  ctx_init(D, qacc_warmstart, &warm); update_constraint(D, &warm);
  ctx_init(D, D->qacc_smooth, &smth); update_constraint(D, &smth);
  D->warm_cost = warm.cost; D->smooth_cost = smth.cost;
  D->used_warmstart = warm.cost < smth.cost;
  c = D->used_warmstart ? &warm : &smth;
  D->cost_start = c->cost;
  for (int r = 0; r < D->nefc; r++) D->zone0[r] = row_zone(D, r, c->Jaref[r]);
  /* gradient, Hessian, Newton direction */
  real grad[NV], H[NV * NV], L[NV * NV], search[NV];
  for (int d = 0; d < NV; d++) grad[d] = c->Ma[d] - D->qfrc_smooth[d] - c->qfrc_constraint[d];
  memcpy(H, D->qM, sizeof(H));
  for (int r = 0; r < D->nefc; r++) {
    if (row_zone(D, r, c->Jaref[r]) != 1) continue;
    for (int i = 0; i < NV; i++) {
      real ji = D->efc_J[r][i] * D->efc_D[r];
      if (ji == 0) continue;
      for (int j = 0; j < NV; j++) H[i * NV + j] += ji * D->efc_J[r][j];
    }
  }
  chol_factor(H, L, NV);
  chol_solve(L, grad, search, NV);
  for (int d = 0; d < NV; d++) search[d] = -search[d];
  /* line search */
  real snorm = 0;
  for (int d = 0; d < NV; d++) snorm += search[d] * search[d];
  snorm = R_SQRT(snorm);
  real gtol = M->tolerance * M->ls_tolerance * (snorm * M->meaninertia * (real)NV);
  real mv[NV], jv[MAXEFC], quad_gauss[3];
  mul_m(D, search, mv);
  for (int r = 0; r < D->nefc; r++) {
    real s = 0;
    for (int d = 0; d < NV; d++) s += D->efc_J[r][d] * search[d];
    jv[r] = s;
  }
  real sMa = 0, sq = 0, smv = 0;
  for (int d = 0; d < NV; d++) { sMa += search[d] * c->Ma[d]; sq += search[d] * D->qfrc_smooth[d]; smv += search[d] * mv[d]; }
  quad_gauss[0] = c->gauss; quad_gauss[1] = sMa - sq; quad_gauss[2] = (real)0.5 * smv;
  LSPoint p0 = ls_point(D, c, 0, jv, quad_gauss);
  LSPoint lo = ls_point(D, c, p0.alpha - p0.d0 / p0.d1, jv, quad_gauss), hi;
  if (lo.d0 < p0.d0) { hi = p0; } else { hi = lo; lo = p0; }
  int swap = 1, it = 0;
  while (it < M->ls_iterations && swap && !((lo.d0 < 0) && (lo.d0 > -gtol)) && !((hi.d0 > 0) && (hi.d0 < gtol))) {
    LSPoint lo_next = ls_point(D, c, lo.alpha - lo.d0 / lo.d1, jv, quad_gauss);
    LSPoint hi_next = ls_point(D, c, hi.alpha - hi.d0 / hi.d1, jv, quad_gauss);
    LSPoint mid = ls_point(D, c, (real)0.5 * (lo.alpha + hi.alpha), jv, quad_gauss);
    int s1 = in_bracket(lo, lo_next); if (s1) lo = lo_next;
    int s2 = in_bracket(lo, mid); if (s2) lo = mid;
    int s3 = in_bracket(lo, hi_next); if (s3) lo = hi_next;
    int t1 = in_bracket(hi, hi_next); if (t1) hi = hi_next;
    int t2 = in_bracket(hi, mid); if (t2) hi = mid;
    int t3 = in_bracket(hi, lo_next); if (t3) hi = lo_next;
    swap = s1 | s2 | s3 | t1 | t2 | t3;
    it++;
  }
  D->ls_iters = it;
  int improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  real alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  if (!improved) alpha = 0;
  D->ls_alpha = alpha;
  for (int d = 0; d < NV; d++) { c->qacc[d] += search[d] * alpha; c->Ma[d] += mv[d] * alpha; }
  for (int r = 0; r < D->nefc; r++) c->Jaref[r] += jv[r] * alpha;
  update_constraint(D, c);
  D->cost_end = c->cost;
  for (int d = 0; d < NV; d++) { D->qacc[d] = c->qacc[d]; D->qfrc_constraint[d] = c->qfrc_constraint[d]; }
  for (int r = 0; r < D->nefc; r++) D->efc_force[r] = c->force[r];
}

static void forward(const Model *M, Data *D, real *qpos, const real *qvel, const real *ctrl, const real *warm) {
  kinematics(M, qpos, D);
  com_pos(M, D);
  crb_and_factor(M, D);
  collision(M, D);
  make_constraint(M, D, qpos, qvel);
  fwd_velocity(M, D, qvel);
  fwd_actuation_acceleration(M, D, qpos, qvel, ctrl);
  solve(M, D, warm);
}

/* semi-implicit Euler, eulerdamp disabled (A.9) */
static void euler(const Model *M, const Data *D, real *qpos, real *qvel) {
  real dt = M->timestep;
  for (int d = 0; d < NV; d++) qvel[d] = qvel[d] + D->qacc[d] * dt;
  for (int k = 0; k < 3; k++) qpos[k] = qpos[k] + dt * qvel[k];
  real w[3] = {qvel[3], qvel[4], qvel[5]};
  real nrm = normalize_n(w, 3);
  real half = dt * nrm * (real)0.5, s = R_SIN(half);
  real qr[4] = {R_COS(half), w[0] * s, w[1] * s, w[2] * s}, qn[4];
  quat_mul(qpos + 3, qr, qn);
  normalize_n(qn, 4);
  for (int k = 0; k < 4; k++) qpos[3 + k] = qn[k];
  for (int j = 0; j < NU; j++) qpos[7 + j] = qpos[7 + j] + dt * qvel[6 + j];
}

/* ------------------------------------------------------------------------------------------
 * env level (environment.py)
 * ---------------------------------------------------------------------------------------- */
typedef struct { real x_pos[13][3], x_rot[13][4], xd_vel[13][3], xd_ang[13][3]; } BraxX;

static void brax_x_xd(const Data *D, BraxX *X) { /* brax/mjx/pipeline.py (A.9) */
  for (int b = 1; b < NB; b++) {
    real off[3], c[3];
    for (int k = 0; k < 3; k++) { X->x_pos[b - 1][k] = D->xpos[b][k]; off[k] = D->xpos[b][k] - D->com[k]; }
    for (int k = 0; k < 4; k++) X->x_rot[b - 1][k] = D->xquat[b][k];
    cross3(off, D->cvel[b], c); /* vel = lin - off x ang */
    for (int k = 0; k < 3; k++) { X->xd_ang[b - 1][k] = D->cvel[b][k]; X->xd_vel[b - 1][k] = D->cvel[b][3 + k] - c[k]; }
  }
}

static inline real unif(int row, const uint32_t key[2], uint32_t i, float lo, float hi) { return (real)du(row, key, i, lo, hi); }

static void sample_command(const PupperEnvCfg *cfg, const uint32_t key[2], real cmd[3]) { /* environment.py:246-272 */
  uint32_t k[6][2];
  for (uint32_t i = 0; i < 6; i++) split_key(key, i, k[i]);
  real c0 = unif(35, k[1], 0, cfg->lin_vel_x[0], cfg->lin_vel_x[1]);
  real c1 = unif(36, k[2], 0, cfg->lin_vel_y[0], cfg->lin_vel_y[1]);
  real c2 = unif(37, k[3], 0, cfg->ang_vel_yaw[0], cfg->ang_vel_yaw[1]);
  float zp = du(38, k[4], 0, 0.0f, 1.0f);
  float thr = cfg->stand_still_command_threshold;
  if (zp < cfg->zero_command_probability) {
    for (uint32_t i = 0; i < 3; i++) cmd[i] = unif(39 + (int)i, k[5], i, -thr, thr);
  } else { cmd[0] = c0; cmd[1] = c1; cmd[2] = c2; }
}

static void sample_body_orientation(const PupperEnvCfg *cfg, const uint32_t key[2], real out[3]) { /* environment.py:274-298 */
  uint32_t kp[2], kr[2];
  split_key(key, 1, kp);
  split_key(key, 2, kr);
  /* float32 throughout in the reference; the f64 build keeps the f32 draws and widens the trigonometry */
  real pitch = (real)(float)(du(42, kp, 0, -1.0f, 1.0f) * cfg->maximum_pitch_command);
  real roll = (real)(float)(du(43, kr, 0, -1.0f, 1.0f) * cfg->maximum_roll_command);
  real v[3] = {roll, pitch, 0}, c[3], s[3];
#ifdef ORACLE_F32
  const real pi = 3.14159274101257324f;
#else
  const real pi = 3.14159265358979323846;
#endif
  for (int i = 0; i < 3; i++) { real a = v[i] * pi / 360; c[i] = R_COS(a); s[i] = R_SIN(a); }
  real q[4] = {c[0] * c[1] * c[2] - s[0] * s[1] * s[2], s[0] * c[1] * c[2] + c[0] * s[1] * s[2],
               c[0] * s[1] * c[2] - s[0] * c[1] * s[2], c[0] * c[1] * s[2] + s[0] * s[1] * c[2]};
  real z[3] = {(real)cfg->desired_world_z_in_body_frame[0], (real)cfg->desired_world_z_in_body_frame[1], (real)cfg->desired_world_z_in_body_frame[2]};
  rotate(z, q, out);
}

/* utils.sample_lagged_value (utils.py:49-69): push front, categorical pick of one column */
static int lagged(int row, const uint32_t key[2], real *buf, int rows, int L, const real *newv, const float *p, real *out) {
  for (int j = 0; j < rows; j++) {
    for (int l = L - 1; l > 0; l--) buf[j * L + l] = buf[j * L + l - 1];
    buf[j * L + 0] = newv[j];
  }
  int idx = dchoice(row, key, p, L);
  if (idx > L - 1) idx = L - 1; /* jnp.take clips out-of-range indices */
  for (int j = 0; j < rows; j++) out[j] = buf[j * L + idx];
  return idx;
}

typedef struct {
  real qpos[NQ], qvel[NV], warm[NV], last_act[NU], abuf[NU * PUPPER_MAX_LAT], ibuf[6 * PUPPER_MAX_LAT];
  real last_vel[NU], command[3], desired_z[3], air_time[4], kick[2], obs[PUPPER_OBS_DIM * ORACLE_MAX_HIST];
} EnvR;

static void env_load(const OracleEnv *e, EnvR *r) {
#define LD(dst, src, n) for (int _i = 0; _i < (n); _i++) (dst)[_i] = (real)(src)[_i]
  LD(r->qpos, e->qpos, NQ); LD(r->qvel, e->qvel, NV); LD(r->warm, e->qacc_warmstart, NV);
  LD(r->last_act, e->last_act, NU); LD(r->abuf, e->action_buffer, NU * PUPPER_MAX_LAT);
  LD(r->ibuf, e->imu_buffer, 6 * PUPPER_MAX_LAT); LD(r->last_vel, e->last_vel, NU);
  LD(r->command, e->command, 3); LD(r->desired_z, e->desired_world_z, 3); LD(r->air_time, e->feet_air_time, 4);
  LD(r->kick, e->kick, 2); LD(r->obs, e->obs, PUPPER_OBS_DIM * ORACLE_MAX_HIST);
}
static void env_store(OracleEnv *e, const EnvR *r) {
#define ST(dst, src, n) for (int _i = 0; _i < (n); _i++) (dst)[_i] = (double)(src)[_i]
  ST(e->qpos, r->qpos, NQ); ST(e->qvel, r->qvel, NV); ST(e->qacc_warmstart, r->warm, NV);
  ST(e->last_act, r->last_act, NU); ST(e->action_buffer, r->abuf, NU * PUPPER_MAX_LAT);
  ST(e->imu_buffer, r->ibuf, 6 * PUPPER_MAX_LAT); ST(e->last_vel, r->last_vel, NU);
  ST(e->command, r->command, 3); ST(e->desired_world_z, r->desired_z, 3); ST(e->feet_air_time, r->air_time, 4);
  ST(e->kick, r->kick, 2); ST(e->obs, r->obs, PUPPER_OBS_DIM * ORACLE_MAX_HIST);
}

/* _get_obs (environment.py:485-543); mutates rng, imu buffer, obs history */
static void get_obs(const PupperEnvCfg *cfg, const BraxX *X, uint32_t rng[2], EnvR *r, const real *q_joint, OracleDebug *dbg) {
  real inv[4] = {1, 0, 0, 0}, ang[3] = {0, 0, 0};
  if (cfg->use_imu) { quat_inv(X->x_rot[0], inv); rotate(X->xd_ang[0], inv, ang); }
  uint32_t k[6][2];
  for (uint32_t i = 0; i < 6; i++) split_key(rng, i, k[i]);
  rng[0] = k[0][0]; rng[1] = k[0][1];
  real imu[6], down[3] = {0, 0, -1}, g[3];
  rotate(down, inv, g);
  for (uint32_t i = 0; i < 3; i++) {
    real an = (real)(float)(du(4 + (int)i, k[1], i, -1.0f, 1.0f) * cfg->angular_velocity_noise);
    real gn = (real)(float)(du(7 + (int)i, k[2], i, -1.0f, 1.0f) * cfg->gravity_noise);
    imu[i] = ang[i] + an;
    g[i] = g[i] + gn;
  }
  real gnorm = R_SQRT(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  for (int i = 0; i < 3; i++) imu[3 + i] = g[i] / gnorm;
  real lag[6];
  int il = lagged(34, k[5], r->ibuf, 6, cfg->n_imu_latency, imu, cfg->imu_latency_distribution, lag);
  if (dbg) dbg->imu_lag = il;
  real obs[PUPPER_OBS_DIM];
  for (int i = 0; i < 6; i++) obs[i] = lag[i];
  for (int i = 0; i < 3; i++) { obs[6 + i] = r->command[i]; obs[9 + i] = r->desired_z[i]; }
  for (uint32_t i = 0; i < 12; i++) {
    real mn = (real)(float)(du(10 + (int)i, k[3], i, -1.0f, 1.0f) * cfg->motor_angle_noise);
    real ln = (real)(float)(du(22 + (int)i, k[4], i, -1.0f, 1.0f) * cfg->last_action_noise);
    obs[12 + i] = q_joint[i] - (real)cfg->default_pose[i] + mn;
    obs[24 + i] = r->last_act[i] + ln;
  }
  int H = cfg->observation_history;
  for (int i = H * PUPPER_OBS_DIM - 1; i >= PUPPER_OBS_DIM; i--) r->obs[i] = r->obs[i - PUPPER_OBS_DIM];
  for (int i = 0; i < PUPPER_OBS_DIM; i++) r->obs[i] = r_clip(obs[i], -100, 100);
}

static void debug_fill(const Data *D, const BraxX *X, OracleDebug *g) {
#define DB(dst, src, n) for (int _i = 0; _i < (n); _i++) (dst)[_i] = (double)((const real *)(src))[_i]
  DB(g->xpos, D->xpos, NB * 3); DB(g->xquat, D->xquat, NB * 4); DB(g->xipos, D->xipos, NB * 3);
  DB(g->subtree_com, D->com, 3); DB(g->cinert, D->cinert, NB * 10); DB(g->cdof, D->cdof, NV * 6); DB(g->cvel, D->cvel, NB * 6);
  DB(g->qM, D->qM, NV * NV); DB(g->qfrc_bias, D->qfrc_bias, NV); DB(g->qfrc_passive, D->qfrc_passive, NV);
  DB(g->qfrc_actuator, D->qfrc_actuator, NV); DB(g->qfrc_smooth, D->qfrc_smooth, NV); DB(g->qacc_smooth, D->qacc_smooth, NV);
  DB(g->qacc, D->qacc, NV); DB(g->qfrc_constraint, D->qfrc_constraint, NV);
  DB(g->x_pos, X->x_pos, 39); DB(g->x_rot, X->x_rot, 52); DB(g->xd_vel, X->xd_vel, 39); DB(g->xd_ang, X->xd_ang, 39);
  DB(g->site_xpos, D->site_xpos, PUPPER_NSITE * 3); DB(g->sphere_xpos, D->sphere_xpos, PUPPER_NSPHERE * 3);
  memset(g->contact_dist, 0, sizeof(g->contact_dist)); memset(g->contact_geom, 0, sizeof(g->contact_geom));
  DB(g->contact_dist, D->con_dist, D->ncon); DB(g->contact_pos, D->con_pos, D->ncon * 3);
  DB(g->contact_frame, D->con_frame, D->ncon * 9); DB(g->contact_mu, D->con_mu, D->ncon);
  for (int c = 0; c < D->ncon; c++) { g->contact_geom[2 * c] = D->con_geom[c][0]; g->contact_geom[2 * c + 1] = D->con_geom[c][1]; }
  memset(g->efc_J, 0, sizeof(g->efc_J));
  for (int r = 0; r < D->nefc; r++) for (int d = 0; d < NV; d++) g->efc_J[r * NV + d] = (double)D->efc_J[r][d];
  DB(g->efc_D, D->efc_D, D->nefc); DB(g->efc_aref, D->efc_aref, D->nefc); DB(g->efc_force, D->efc_force, D->nefc);
  DB(g->efc_pos, D->efc_pos, D->nefc);
  g->ls_alpha = (double)D->ls_alpha; g->cost_start = (double)D->cost_start; g->cost_end = (double)D->cost_end;
  g->warm_cost = (double)D->warm_cost; g->smooth_cost = (double)D->smooth_cost;
  g->ncon = D->ncon; g->nefc = D->nefc; g->ls_iters = D->ls_iters; g->used_warmstart = D->used_warmstart;
  for (int r = 0; r < D->nefc; r++) g->efc_zone0[r] = D->zone0[r];
}

static void env_reset(const PupperModelDesc *md, const PupperEnvCfg *cfg, const uint32_t key_in[2], const OracleDR *dr,
                      OracleEnv *e, OracleDebug *dbg) { /* environment.py:314-346 */
  Model M;
  Data D;
  EnvR r;
  BraxX X;
  model_load(&M, md, dr);
  memset(&r, 0, sizeof(r));
  uint32_t k[4][2], kq[3][2];
  for (uint32_t i = 0; i < 4; i++) split_key(key_in, i, k[i]);
  /* randomize_qpos (domain_randomization.py:188-210) */
  for (uint32_t i = 0; i < 3; i++) split_key(k[3], i, kq[i]);
  for (int i = 0; i < NQ; i++) r.qpos[i] = (real)cfg->init_q[i];
  for (uint32_t i = 0; i < 3; i++) r.qpos[i] = unif((int)i, kq[1], i, cfg->start_pos_min[i], cfg->start_pos_max[i]);
  {
    float yaw = du(3, kq[2], 0, -3.14159274101257324f, 3.14159274101257324f);
    real h = (real)yaw / 2;
    r.qpos[3] = R_COS(h); r.qpos[4] = 0; r.qpos[5] = 0; r.qpos[6] = R_SIN(h);
  }
  real ctrl[NU] = {0};
  forward(&M, &D, r.qpos, r.qvel, ctrl, r.warm); /* pipeline_init = make_data + forward */
  for (int d = 0; d < NV; d++) r.warm[d] = D.qacc[d];
  brax_x_xd(&D, &X);
  for (int j = 0; j < 6; j++) for (int l = 0; l < cfg->n_imu_latency; l++) r.ibuf[j * cfg->n_imu_latency + l] = (j == 5) ? -1 : 0;
  uint32_t rng[2] = {k[0][0], k[0][1]};
  sample_command(cfg, k[1], r.command);
  sample_body_orientation(cfg, k[2], r.desired_z);
  if (dbg) { memset(dbg, 0, sizeof(*dbg)); debug_fill(&D, &X, dbg); }
  get_obs(cfg, &X, rng, &r, r.qpos + 7, dbg);
  memset(e, 0, sizeof(*e));
  env_store(e, &r);
  if (!g_ext) { e->rng[0] = rng[0]; e->rng[1] = rng[1]; } else { e->rng[0] = key_in[0]; e->rng[1] = key_in[1]; }
  /* AutoResetWrapper.reset: remember the first pipeline state and obs */
  for (int i = 0; i < NQ; i++) e->first_qpos[i] = e->qpos[i];
  for (int i = 0; i < NV; i++) { e->first_qvel[i] = e->qvel[i]; e->first_warmstart[i] = e->qacc_warmstart[i]; }
  memcpy(e->first_obs, e->obs, sizeof(e->obs));
}

static void env_step(const PupperModelDesc *md, const PupperEnvCfg *cfg, const OracleDR *dr, OracleEnv *e,
                     const double *action_in, int episode, OracleDebug *dbg) { /* environment.py:348-483 */
  Model M;
  Data D;
  EnvR r;
  BraxX X;
  model_load(&M, md, dr);
  env_load(e, &r);
  real action[NU];
  for (int i = 0; i < NU; i++) action[i] = (real)action_in[i];
  if (dbg) memset(dbg, 0, sizeof(*dbg));
  /* S1 */
  uint32_t k[5][2], rng[2];
  for (uint32_t i = 0; i < 5; i++) split_key(e->rng, i, k[i]);
  rng[0] = k[0][0]; rng[1] = k[0][1];
  /* S2 kick */
  real kick[2];
  {
    int hit = du(2, k[3], 0, 0.0f, 1.0f) < cfg->kick_probability;
    for (uint32_t i = 0; i < 2; i++) {
      float u = du((int)i, k[2], i, -1.0f, 1.0f) * cfg->kick_vel;
      kick[i] = (real)(u * (float)hit);
      r.qvel[i] = kick[i] + r.qvel[i];
    }
  }
  /* S3 action latency, S4 motor targets */
  real lag[NU], ctrl[NU];
  int al = lagged(3, k[4], r.abuf, NU, cfg->n_latency, action, cfg->latency_distribution, lag);
  for (int i = 0; i < NU; i++)
    ctrl[i] = r_clip((real)cfg->default_pose[i] + lag[i] * (real)cfg->action_scale, (real)cfg->joint_lower[i], (real)cfg->joint_upper[i]);
  /* S5 physics: n_frames x mjx.step */
  for (int f = 0; f < cfg->n_frames; f++) {
    forward(&M, &D, r.qpos, r.qvel, ctrl, r.warm);
    for (int d = 0; d < NV; d++) r.warm[d] = D.qacc[d];
    euler(&M, &D, r.qpos, r.qvel);
  }
  brax_x_xd(&D, &X);
  if (dbg) { debug_fill(&D, &X, dbg); dbg->act_lag = al; for (int i = 0; i < NU; i++) dbg->motor_targets[i] = (double)ctrl[i]; }
  /* S6 obs (reads the not-yet-updated last_act / command / desired_z) */
  get_obs(cfg, &X, rng, &r, r.qpos + 7, dbg);
  const real *q = r.qpos + 7, *qd = r.qvel + 6;
  /* S7 foot contacts */
  real dt = (real)cfg->dt;
  int contact = 0, filt_mm = 0, filt_cm = 0, first = 0;
  for (int f = 0; f < 4; f++) {
    real z = D.site_xpos[1 + f][2] - (real)cfg->foot_radius;
    int lc = (e->last_contact >> f) & 1;
    int c = z < (real)1e-3;
    contact |= c << f;
    filt_mm |= (c | lc) << f;
    filt_cm |= ((z < (real)3e-2) | lc) << f;
    first |= ((r.air_time[f] > 0) && (c | lc)) << f;
    r.air_time[f] += dt;
    if (dbg) dbg->foot_z[f] = (double)z;
  }
  /* S8 termination */
  real up[3] = {0, 0, 1}, rup[3];
  rotate(up, X.x_rot[0], rup);
  int done = rup[2] < (real)cfg->cos_terminal_body_angle;
  real margin = (real)1e30;
  for (int i = 0; i < NU; i++) {
    done |= q[i] < (real)cfg->joint_lower[i];
    done |= q[i] > (real)cfg->joint_upper[i];
    margin = r_min(margin, r_min(q[i] - (real)cfg->joint_lower[i], (real)cfg->joint_upper[i] - q[i]));
  }
  done |= X.x_pos[0][2] < (real)cfg->terminal_body_z;
  if (dbg) { dbg->up_dot = (double)rup[2]; dbg->min_limit_margin = (double)margin; dbg->torso_z = (double)X.x_pos[0][2]; }
  /* S9 rewards (rewards.py) */
  real rw[PUPPER_NREWARD], inv[4], lv[3], av[3], sigma = (real)cfg->tracking_sigma;
  quat_inv(X.x_rot[0], inv);
  rotate(X.xd_vel[0], inv, lv);
  rotate(X.xd_ang[0], inv, av);
  {
    real e0 = r.command[0] - lv[0], e1 = r.command[1] - lv[1];
    rw[PUPPER_R_TRACKING_LIN_VEL] = R_EXP(-(e0 * e0 + e1 * e1) / sigma);
    real e2 = r.command[2] - av[2];
    rw[PUPPER_R_TRACKING_ANG_VEL] = R_EXP(-(e2 * e2) / sigma);
    real wz[3];
    rotate(up, inv, wz);
    real err = 0;
    for (int i = 0; i < 3; i++) err += (wz[i] - r.desired_z[i]) * (wz[i] - r.desired_z[i]);
    rw[PUPPER_R_TRACKING_ORIENTATION] = R_EXP(-err / sigma);
  }
  rw[PUPPER_R_LIN_VEL_Z] = X.xd_vel[0][2] * X.xd_vel[0][2];
  rw[PUPPER_R_ANG_VEL_XY] = X.xd_ang[0][0] * X.xd_ang[0][0] + X.xd_ang[0][1] * X.xd_ang[0][1];
  rw[PUPPER_R_ORIENTATION] = rup[0] * rup[0] + rup[1] * rup[1];
  {
    real s = 0, ja = 0, mw = 0, ar = 0, ss = 0, sv = 0, ab = 0;
    for (int d = 0; d < NV; d++) s += D.qfrc_actuator[d] * D.qfrc_actuator[d];
    for (int i = 0; i < NU; i++) {
      real a = (qd[i] - r.last_vel[i]) / (real)cfg->env_dt;
      ja += a * a;
      mw += R_ABS(D.qfrc_actuator[6 + i] * qd[i]);
      ar += (action[i] - r.last_act[i]) * (action[i] - r.last_act[i]);
      ss += R_ABS(q[i] - (real)cfg->default_pose[i]);
      sv += R_ABS(qd[i]);
    }
    for (int l = 0; l < 4; l++) { real a = q[1 + 3 * l] - (real)cfg->desired_abduction[l]; ab += a * a; }
    real cn = brax_norm3(r.command);
    rw[PUPPER_R_TORQUES] = s; rw[PUPPER_R_JOINT_ACCELERATION] = ja; rw[PUPPER_R_MECHANICAL_WORK] = mw;
    rw[PUPPER_R_ACTION_RATE] = ar;
    rw[PUPPER_R_STAND_STILL] = ss * (cn < (real)0.1 ? 1 : 0);
    rw[PUPPER_R_STAND_STILL_JOINT_VELOCITY] = sv * (cn < (real)cfg->stand_still_command_threshold ? 1 : 0);
    rw[PUPPER_R_ABDUCTION_ANGLE] = ab;
    real at = 0;
    for (int f = 0; f < 4; f++) at += (r.air_time[f] - (real)0.1) * (real)((first >> f) & 1);
    rw[PUPPER_R_FEET_AIR_TIME] = at * (cn > (real)0.05 ? 1 : 0);
  }
  {
    real slip = 0;
    for (int f = 0; f < 4; f++) {
      int b = 4 + 3 * f; /* lower-leg body id */
      real off[3], c[3];
      for (int kk = 0; kk < 3; kk++) off[kk] = D.site_xpos[1 + f][kk] - D.xpos[b][kk];
      cross3(off, X.xd_ang[b - 1], c);
      real vx = X.xd_vel[b - 1][0] - c[0], vy = X.xd_vel[b - 1][1] - c[1];
      real m = (real)((filt_cm >> f) & 1);
      slip += vx * vx * m + vy * vy * m;
    }
    rw[PUPPER_R_FOOT_SLIP] = slip;
  }
  rw[PUPPER_R_TERMINATION] = (real)(done && (e->step < cfg->early_termination_step_threshold));
  {
    real knee = 0, body = 0;
    for (int c = 0; c < D.ncon; c++) {
      if (!(D.con_dist[c] < 0)) continue;
      for (int side = 0; side < 2; side++)
        for (int s = 0; s < PUPPER_NSPHERE; s++)
          if (D.con_geom[c][side] == M.sphere_geomid[s]) {
            knee += (real)((cfg->knee_sphere_mask >> s) & 1u);
            body += (real)((cfg->torso_sphere_mask >> s) & 1u);
          }
    }
    rw[PUPPER_R_KNEE_COLLISION] = knee; rw[PUPPER_R_BODY_COLLISION] = body;
  }
  real total = 0, scaled[PUPPER_NREWARD];
  for (int i = 0; i < PUPPER_NREWARD; i++) { scaled[i] = rw[i] * (real)cfg->reward_scales[i]; total += scaled[i]; }
  real reward = r_clip(total * dt, 0, 10000);
  if (dbg) {
    for (int i = 0; i < PUPPER_NREWARD; i++) dbg->rewards_raw[i] = (double)rw[i];
    dbg->contact_flags = contact | (filt_cm << 4) | (first << 8);
  }
  /* S10 bookkeeping */
  for (int i = 0; i < 2; i++) r.kick[i] = kick[i];
  for (int i = 0; i < NU; i++) { r.last_act[i] = action[i]; r.last_vel[i] = qd[i]; }
  for (int f = 0; f < 4; f++) r.air_time[f] *= (real)(!((filt_mm >> f) & 1));
  int step = e->step + 1;
  /* S11 resample command and desired orientation, both from cmd_rng */
  int resample = step > cfg->resample_velocity_step;
  if (resample) { sample_command(cfg, k[1], r.command); sample_body_orientation(cfg, k[1], r.desired_z); }
  if (done || resample) step = 0;
  if (dbg) dbg->resampled = resample;
  /* S12 outputs */
  env_store(e, &r);
  if (!g_ext) { e->rng[0] = rng[0]; e->rng[1] = rng[1]; } /* external randoms leave the key alone */
  e->last_contact = (uint32_t)contact;
  e->step = step;
  e->reward = (double)reward;
  e->done = (double)done;
  e->metrics[0] = (double)brax_norm3(X.x_pos[0]);
  for (int i = 0; i < PUPPER_NREWARD; i++) e->metrics[1 + i] = (double)scaled[i];

  if (episode) { /* brax EpisodeWrapper then AutoResetWrapper, float32 bookkeeping */
    /* AutoResetWrapper.step zeroes info["steps"] where the previous step was done; EpisodeWrapper.step
     * adds to the episode accumulators and THEN multiplies them by (1 - prev_done), so the first step
     * after an episode end is dropped from the sums (brax 0.12.1 behaviour, kept as is). */
    real keep = e->episode_done != 0.0 ? 0 : 1;
    if (e->episode_done != 0.0) e->steps = 0;
    e->steps += cfg->action_repeat;
    int was_done = e->done != 0.0;
    int trunc = e->steps >= cfg->episode_length;
    e->truncation = (double)(trunc && !was_done);
    if (trunc) e->done = 1.0;
    e->sum_reward = (double)(((real)e->sum_reward + (real)e->reward) * keep);
    e->length = (double)(((real)e->length + (real)cfg->action_repeat) * keep);
    for (int i = 0; i < PUPPER_NMETRIC; i++) e->sum_metrics[i] = (double)(((real)e->sum_metrics[i] + (real)e->metrics[i]) * keep);
    e->episode_done = e->done;
    if (e->done != 0.0) {
      memcpy(e->qpos, e->first_qpos, sizeof(e->qpos)); memcpy(e->qvel, e->first_qvel, sizeof(e->qvel));
      memcpy(e->qacc_warmstart, e->first_warmstart, sizeof(e->qacc_warmstart));
      memcpy(e->obs, e->first_obs, sizeof(e->obs));
    }
  }
}

/* ext_u: NULL, or host float [n][PUPPER_NRAND] raw uniforms in [0, 1) (rows as documented for PupperRand) */
int SUF(oracle_reset_ext)(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const uint32_t *keys,
                          const OracleDR *dr, OracleEnv *envs, OracleDebug *dbg, int n_threads, const float *ext_u) {
  if (!m || !cfg || !keys || !envs || n <= 0 || cfg->observation_history > ORACLE_MAX_HIST) return -1;
#ifdef _OPENMP
  if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for num_threads(n_threads) schedule(static)
#endif
  for (int i = 0; i < n; i++) {
    g_ext = ext_u ? ext_u + (size_t)PUPPER_NRAND * i : NULL;
    env_reset(m, cfg, keys + 2 * i, dr ? dr + i : NULL, envs + i, dbg ? dbg + i : NULL);
    g_ext = NULL;
  }
  return 0;
}
int SUF(oracle_reset)(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const uint32_t *keys,
                      const OracleDR *dr, OracleEnv *envs, OracleDebug *dbg, int n_threads) {
  return SUF(oracle_reset_ext)(m, cfg, n, keys, dr, envs, dbg, n_threads, NULL);
}

int SUF(oracle_step_ext)(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const OracleDR *dr, OracleEnv *envs,
                         const double *action, int episode, OracleDebug *dbg, int n_threads, const float *ext_u) {
  if (!m || !cfg || !envs || !action || n <= 0 || cfg->observation_history > ORACLE_MAX_HIST) return -1;
#ifdef _OPENMP
  if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for num_threads(n_threads) schedule(static)
#endif
  for (int i = 0; i < n; i++) {
    g_ext = ext_u ? ext_u + (size_t)PUPPER_NRAND * i : NULL;
    env_step(m, cfg, dr ? dr + i : NULL, envs + i, action + NU * i, episode, dbg ? dbg + i : NULL);
    g_ext = NULL;
  }
  return 0;
}
int SUF(oracle_step)(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const OracleDR *dr, OracleEnv *envs,
                     const double *action, int episode, OracleDebug *dbg, int n_threads) {
  return SUF(oracle_step_ext)(m, cfg, n, dr, envs, action, episode, dbg, n_threads, NULL);
}

/* Physics only, one env: what the reference reaches through Brax -- pipeline_init (n_frames <= 0: one forward pass, the
 * warm start becomes its qacc) and pipeline_step (n_frames x mjx.step; environment.py:319,366) -- so that the reference's own
 * env-level code can be run with the oracle's physics underneath (tests/refshim).  In/out: qpos[19], qvel[18], warm[18]. */
int SUF(oracle_pipeline)(const PupperModelDesc *m, const OracleDR *dr, int n_frames, double *qpos, double *qvel, double *warm,
                         const double *ctrl_in, OracleDebug *dbg) {
  if (!m || !qpos || !qvel || !warm || !ctrl_in) return -1;
  Model M;
  Data D;
  BraxX X;
  model_load(&M, m, dr);
  real q[NQ], v[NV], w[NV], ctrl[NU];
  for (int i = 0; i < NQ; i++) q[i] = (real)qpos[i];
  for (int i = 0; i < NV; i++) { v[i] = (real)qvel[i]; w[i] = (real)warm[i]; }
  for (int i = 0; i < NU; i++) ctrl[i] = (real)ctrl_in[i];
  if (n_frames <= 0) {
    forward(&M, &D, q, v, ctrl, w);
    for (int d = 0; d < NV; d++) w[d] = D.qacc[d];
  } else {
    for (int f = 0; f < n_frames; f++) {
      forward(&M, &D, q, v, ctrl, w);
      for (int d = 0; d < NV; d++) w[d] = D.qacc[d];
      euler(&M, &D, q, v);
    }
  }
  brax_x_xd(&D, &X);
  if (dbg) { memset(dbg, 0, sizeof(*dbg)); debug_fill(&D, &X, dbg); }
  for (int i = 0; i < NQ; i++) qpos[i] = (double)q[i];
  for (int i = 0; i < NV; i++) { qvel[i] = (double)v[i]; warm[i] = (double)w[i]; }
  return 0;
}

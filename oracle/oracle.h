/*
 * oracle.h -- data exchanged with the CPU oracle (TEST INFRASTRUCTURE, never part of the product path).
 *
 * The oracle is a plain-C restatement of the batched PupperV3Env step of rishihahs/pupperv3-mjx
 * (reference pupperv3_mjx/environment.py:314-543, rewards.py:9-138, utils.py:34-69,
 * domain_randomization.py:188-210) including the third-party physics that path delegates to
 * (mujoco_mjx==3.2.7 mjx.step, brax==0.12.1 mjx pipeline, jax==0.5.0 threefry PRNG; none of them is
 * vendored in the reference nor installable here -- their published algorithms are restated, see
 * SURVEY.md Appendix A).
 * PARITY STATUS: the env level (everything the reference itself implements on the path) is PINNED against the
 * reference's own code executed in the build container (tests/refshim + tests/golden/make_ref_golden.py ->
 * tests/golden/ref_*.npz, checked step by step by tests/test_ref_golden.py); the third-party physics under it is
 * UNPINNED against MJX (no installable MJX, no physics goldens in the reference) and is checked by physics invariants,
 * closed forms and known-answer tests only.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 *
 * All floating-point fields are double at this boundary regardless of the precision the oracle
 * computes in (the f32 build converts on entry/exit; f32 values are exactly representable).
 */
#ifndef PUPPER_ORACLE_H_
#define PUPPER_ORACLE_H_

#include <stdint.h>
#include "../include/pupper_env.h"

#define ORACLE_MAX_HIST 16
#define ORACLE_MAX_EFC 64

/* One env: Brax State fields the reference populates. */
typedef struct OracleEnv {
  double qpos[PUPPER_NQ];
  double qvel[PUPPER_NV];
  double qacc_warmstart[PUPPER_NV];
  double last_act[PUPPER_NU];
  double action_buffer[PUPPER_NU * PUPPER_MAX_LAT]; /* (12, L) row-major: element (j,l) at j*L+l */
  double imu_buffer[6 * PUPPER_MAX_LAT];            /* (6, L_imu) row-major */
  double last_vel[PUPPER_NU];
  double command[3];
  double desired_world_z[3];
  double feet_air_time[4];
  double kick[2];
  double obs[PUPPER_OBS_DIM * ORACLE_MAX_HIST];
  double reward;
  double done;
  double metrics[PUPPER_NMETRIC];
  /* fused brax EpisodeWrapper / AutoResetWrapper state */
  double first_qpos[PUPPER_NQ], first_qvel[PUPPER_NV], first_warmstart[PUPPER_NV];
  double first_obs[PUPPER_OBS_DIM * ORACLE_MAX_HIST];
  double truncation, sum_reward, length, sum_metrics[PUPPER_NMETRIC], episode_done;
  uint32_t rng[2];
  uint32_t last_contact; /* bit k = foot k */
  int32_t step;          /* info["step"] */
  int32_t steps;         /* EpisodeWrapper info["steps"] */
  int32_t pad_;
} OracleEnv;

/* Per-env domain-randomisation leaves (domain_randomization.py:94-110), compact form. */
typedef struct OracleDR {
  double friction;
  double kp;
  double kd;
  double base_ipos[3];
  double body_inertia[13 * 3]; /* bodies 1..13 */
  double body_mass[13];
} OracleDR;

/* Intermediates of the LAST forward pass executed for an env (debug taps for tests). */
typedef struct OracleDebug {
  double xpos[PUPPER_NBODY * 3], xquat[PUPPER_NBODY * 4], xipos[PUPPER_NBODY * 3];
  double subtree_com[3];
  double cinert[PUPPER_NBODY * 10], cdof[PUPPER_NV * 6], cvel[PUPPER_NBODY * 6];
  double qM[PUPPER_NV * PUPPER_NV];
  double qfrc_bias[PUPPER_NV], qfrc_passive[PUPPER_NV], qfrc_actuator[PUPPER_NV];
  double qfrc_smooth[PUPPER_NV], qacc_smooth[PUPPER_NV], qacc[PUPPER_NV], qfrc_constraint[PUPPER_NV];
  double x_pos[13 * 3], x_rot[13 * 4], xd_vel[13 * 3], xd_ang[13 * 3];
  double site_xpos[PUPPER_NSITE * 3], sphere_xpos[PUPPER_NSPHERE * 3];
  double contact_dist[PUPPER_MAX_CON], contact_pos[PUPPER_MAX_CON * 3], contact_frame[PUPPER_MAX_CON * 9];
  double contact_mu[PUPPER_MAX_CON];
  double efc_J[ORACLE_MAX_EFC * PUPPER_NV], efc_D[ORACLE_MAX_EFC], efc_aref[ORACLE_MAX_EFC];
  double efc_force[ORACLE_MAX_EFC], efc_pos[ORACLE_MAX_EFC];
  double ls_alpha, cost_start, cost_end, warm_cost, smooth_cost;
  double foot_z[4];          /* site z - foot_radius */
  double up_dot, min_limit_margin, torso_z; /* termination inputs */
  double rewards_raw[PUPPER_NREWARD];
  double motor_targets[PUPPER_NU];
  int32_t contact_geom[PUPPER_MAX_CON * 2];
  int32_t ncon, nefc, ls_iters, used_warmstart;
  int32_t contact_flags; /* bit k: contact, bit 4+k: contact_filt_cm, bit 8+k: first_contact */
  int32_t act_lag, imu_lag, resampled;
  int32_t efc_zone0[ORACLE_MAX_EFC]; /* row zones at the Newton start point: 0 inactive, 1 quadratic, 2 / 3 linear (-/+) */
} OracleDebug;

#ifdef __cplusplus
extern "C" {
#endif

/* f64 and f32 builds export the same functions with _f64 / _f32 suffixes.
 * dr == NULL: nominal model.  dbg == NULL: no taps.  episode != 0 applies the fused
 * EpisodeWrapper + AutoResetWrapper semantics after the env step.  n_threads <= 0: all cores. */
int oracle_sizeof_env(void);
int oracle_sizeof_dr(void);
int oracle_sizeof_debug(void);

int oracle_reset_f64(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const uint32_t *keys,
                     const OracleDR *dr, OracleEnv *envs, OracleDebug *dbg, int n_threads);
int oracle_step_f64(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const OracleDR *dr,
                    OracleEnv *envs, const double *action, int episode, OracleDebug *dbg, int n_threads);
int oracle_reset_f32(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const uint32_t *keys,
                     const OracleDR *dr, OracleEnv *envs, OracleDebug *dbg, int n_threads);
int oracle_step_f32(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const OracleDR *dr,
                    OracleEnv *envs, const double *action, int episode, OracleDebug *dbg, int n_threads);

/* Same with external randoms: ext_u = NULL, or host float [n][PUPPER_NRAND] raw uniforms in [0, 1), rows as documented
 * for PupperRand in include/pupper_env.h; envs[i].rng is then left unchanged. */
int oracle_reset_ext_f64(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const uint32_t *keys,
                         const OracleDR *dr, OracleEnv *envs, OracleDebug *dbg, int n_threads, const float *ext_u);
int oracle_step_ext_f64(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const OracleDR *dr,
                        OracleEnv *envs, const double *action, int episode, OracleDebug *dbg, int n_threads, const float *ext_u);
int oracle_reset_ext_f32(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const uint32_t *keys,
                         const OracleDR *dr, OracleEnv *envs, OracleDebug *dbg, int n_threads, const float *ext_u);
int oracle_step_ext_f32(const PupperModelDesc *m, const PupperEnvCfg *cfg, int n, const OracleDR *dr,
                        OracleEnv *envs, const double *action, int episode, OracleDebug *dbg, int n_threads, const float *ext_u);

/* Physics only, one env (n_frames <= 0: pipeline_init = one forward pass; else n_frames x mjx.step); see pupper_oracle.c */
int oracle_pipeline_f64(const PupperModelDesc *m, const OracleDR *dr, int n_frames, double *qpos, double *qvel, double *warm,
                        const double *ctrl, OracleDebug *dbg);
int oracle_pipeline_f32(const PupperModelDesc *m, const OracleDR *dr, int n_frames, double *qpos, double *qvel, double *warm,
                        const double *ctrl, OracleDebug *dbg);

/* PRNG known-answer access (jax 0.5.0 threefry, partitionable). */
void oracle_threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t *out);
float oracle_uniform(uint32_t k0, uint32_t k1, uint32_t index, float lo, float hi);
int oracle_choice(uint32_t k0, uint32_t k1, const float *p, int n);

#ifdef __cplusplus
}
#endif
#endif

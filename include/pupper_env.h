/*
 * pupper_env.h — C ABI of libpupper_env.so: the batched PupperV3Env reset/step hot path on B200.
 *
 * What this boundary replaces in the reference (rishihahs/pupperv3-mjx; file:line into that repo):
 *   pupper_model_create   <- PupperV3Env.__init__            pupperv3_mjx/environment.py:35-244
 *                            (mjcf.load + gain/bias override :165-174, id caches :183-203)
 *   pupper_reset          <- PupperV3Env.reset               pupperv3_mjx/environment.py:314-346
 *                            (+ randomize_qpos               pupperv3_mjx/domain_randomization.py:188-210)
 *   pupper_step           <- PupperV3Env.step                pupperv3_mjx/environment.py:348-483
 *                            (+ _get_obs :485-543, rewards.py:9-138, utils.sample_lagged_value utils.py:49-69,
 *                             and Brax PipelineEnv.pipeline_step = 5 x mjx.step, environment.py:366)
 *   PupperDR              <- the 6 batched leaves returned by domain_randomize
 *                                                            pupperv3_mjx/domain_randomization.py:93-112
 *   pupper_episode_*      <- brax EpisodeWrapper/AutoResetWrapper semantics, fused
 *
 * The reference has no C interface of its own (it is pure Python on JAX); these entry points are
 * what an XLA-FFI / ctypes binding for that path binds.
 *
 * Conventions: every function returns 0 on success or a negative PUPPER_E* code; nothing throws;
 * nothing synchronises; work is only enqueued on `stream`; all device buffers are caller-owned
 * (the library owns only the PupperModel constant tables), so every call is CUDA-graph capturable.
 * No torch / JAX types appear here: plain pointers and sizes only.
 */
#ifndef PUPPER_ENV_H_
#define PUPPER_ENV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PUPPER_ABI_VERSION 4

/* Fixed topology of the supported robot family: world + base + 4 legs x 3 links. */
#define PUPPER_NBODY 14
#define PUPPER_NQ 19
#define PUPPER_NV 18
#define PUPPER_NU 12
#define PUPPER_NLEG 4
#define PUPPER_NSPHERE 8 /* per leg: link2 "knee" sphere, link3 "foot" sphere */
#define PUPPER_NSITE 5
#define PUPPER_MAX_BOX 32
#define PUPPER_MAX_CON 8   /* storage bound of contact tables (the CPU oracle's taps use it) */
#define PUPPER_MAX_PAIRS 8 /* storage bound of per-group pair tables */
#define PUPPER_KERNEL_MAX_CON 5   /* what the CUDA path accepts: 1 <= max_contact_points <= 5 (one slot table per env) */
#define PUPPER_KERNEL_MAX_PAIRS 4 /* and 1 <= max_geom_pairs <= 4 (one narrow phase per lane of the env's quad); the
                                     reference model sets 5 / 4 (test_pupper_model.xml:227-230); MJX's -1 = "no limit" is
                                     outside the supported family (pupper_model_create: PUPPER_EUNSUPPORTED) */
#define PUPPER_NEFC_TAP 44 /* rows of PupperStepOut.dbg_efc */
#define PUPPER_NRAND 44    /* uniform draws one reset / step consumes per env (PupperRand) */
#define PUPPER_MAX_LAT 8   /* longest latency distribution (action and IMU) */
#define PUPPER_NREWARD 18
#define PUPPER_NMETRIC 19 /* total_dist + 18 scaled reward terms */
#define PUPPER_OBS_DIM 36

enum {
  PUPPER_OK = 0,
  PUPPER_EINVAL = -1,     /* bad argument (null pointer, n_envs <= 0, sizes out of range) */
  PUPPER_EUNSUPPORTED = -2, /* model/config outside the supported family */
  PUPPER_ECUDA = -3,      /* a CUDA runtime call failed; see pupper_last_cuda_error */
  PUPPER_ENOMEM = -4,
  PUPPER_EVERSION = -5
};

/* Reward terms in the order the reference builds its dict (environment.py:391-444). */
enum {
  PUPPER_R_TRACKING_LIN_VEL = 0,
  PUPPER_R_TRACKING_ANG_VEL,
  PUPPER_R_TRACKING_ORIENTATION,
  PUPPER_R_LIN_VEL_Z,
  PUPPER_R_ANG_VEL_XY,
  PUPPER_R_ORIENTATION,
  PUPPER_R_TORQUES,
  PUPPER_R_JOINT_ACCELERATION,
  PUPPER_R_MECHANICAL_WORK,
  PUPPER_R_ACTION_RATE,
  PUPPER_R_STAND_STILL,
  PUPPER_R_STAND_STILL_JOINT_VELOCITY,
  PUPPER_R_ABDUCTION_ANGLE,
  PUPPER_R_FEET_AIR_TIME,
  PUPPER_R_FOOT_SLIP,
  PUPPER_R_TERMINATION,
  PUPPER_R_KNEE_COLLISION,
  PUPPER_R_BODY_COLLISION
};

/* Compiled model constants (what mujoco's compiler + mjx.put_model would hold), host memory.
 * Produced by pupperv3_mjx_b200.mjcf.compile_model from the MJCF. */
typedef struct PupperModelDesc {
  int32_t abi_version;
  /* kinematic tree, MuJoCo body order: 0 world, 1 base, then per leg link1, link2, link3 */
  int32_t body_parent[PUPPER_NBODY];
  float body_pos[PUPPER_NBODY][3];
  float body_quat[PUPPER_NBODY][4];
  float body_ipos[PUPPER_NBODY][3];
  float body_iquat[PUPPER_NBODY][4];
  float body_mass[PUPPER_NBODY];
  float body_inertia[PUPPER_NBODY][3];
  float body_invweight0[PUPPER_NBODY]; /* translational component only (contacts use nothing else) */
  /* dofs: 0-5 free joint, 6+3k+j = leg k hinge j (axis = body-local z, jnt_pos = 0, qpos0 = 0) */
  float dof_armature[PUPPER_NV];
  float dof_damping[PUPPER_NV];
  float dof_frictionloss[PUPPER_NV];
  float dof_invweight0[PUPPER_NV];
  float dof_solref[2]; /* friction-loss rows */
  float dof_solimp[5];
  float jnt_range[PUPPER_NU][2];
  float jnt_solref[2]; /* limit rows */
  float jnt_solimp[5];
  /* actuators: force = clip(gain*ctrl + bias1*q + bias2*qd, forcerange) */
  float act_gain[PUPPER_NU];
  float act_bias1[PUPPER_NU];
  float act_bias2[PUPPER_NU];
  float act_forcerange[PUPPER_NU][2];
  /* colliding geoms */
  int32_t floor_geomid;
  float floor_friction;
  int32_t sphere_body[PUPPER_NSPHERE]; /* order: leg0 knee, leg0 foot, leg1 knee, ... (= geom id order) */
  int32_t sphere_geomid[PUPPER_NSPHERE];
  float sphere_pos[PUPPER_NSPHERE][3];
  float sphere_radius[PUPPER_NSPHERE];
  float sphere_friction[PUPPER_NSPHERE];
  int32_t nbox;
  int32_t box_geomid[PUPPER_MAX_BOX];
  float box_pos[PUPPER_MAX_BOX][3];
  float box_mat[PUPPER_MAX_BOX][9]; /* row-major world rotation */
  float box_size[PUPPER_MAX_BOX][3]; /* half sizes */
  float box_friction[PUPPER_MAX_BOX];
  /* mixed contact parameters per pair type */
  float plane_sphere_solref[2], plane_sphere_solimp[5];
  float sphere_box_solref[2], sphere_box_solimp[5];
  float sphere_sphere_solref[2], sphere_sphere_solimp[5];
  /* sites: 0 imu, 1..4 feet */
  int32_t site_body[PUPPER_NSITE];
  float site_pos[PUPPER_NSITE][3];
  /* options */
  float timestep;
  float gravity[3];
  float impratio;
  float tolerance;
  float ls_tolerance;
  float meaninertia;
  int32_t iterations;        /* must be 1 (single unrolled Newton body) */
  int32_t ls_iterations;
  int32_t max_geom_pairs;    /* 1..PUPPER_MAX_PAIRS */
  int32_t max_contact_points;/* 1..PUPPER_MAX_CON */
  int32_t frictionloss_rows; /* 1: instantiate joint friction-loss constraint rows (MJX >= 3.2.x) */
} PupperModelDesc;

/* Environment configuration = PupperV3Env ctor kwargs (environment.py:35-121) + reward scales
 * (config.py:19-64), resolved to ids/constants on the host. */
typedef struct PupperEnvCfg {
  int32_t abi_version;
  int32_t observation_history;
  int32_t n_frames;                 /* physics substeps per env step (5.0 in the reference, F6) */
  float env_dt;                     /* self._dt   (environment.py:166)  joint-acceleration reward */
  float dt;                         /* self.dt    = timestep*n_frames  air time, reward scale */
  float action_scale;
  float joint_lower[PUPPER_NU];
  float joint_upper[PUPPER_NU];
  float default_pose[PUPPER_NU];
  float desired_abduction[PUPPER_NLEG];
  int32_t resample_velocity_step;
  float lin_vel_x[2], lin_vel_y[2], ang_vel_yaw[2];
  float zero_command_probability;
  float stand_still_command_threshold;
  float maximum_pitch_command, maximum_roll_command; /* degrees */
  float angular_velocity_noise, gravity_noise, motor_angle_noise, last_action_noise;
  float kick_vel, kick_probability;
  float terminal_body_z;
  float cos_terminal_body_angle;    /* float32(np.cos(terminal_body_angle)) */
  int32_t early_termination_step_threshold;
  float foot_radius;
  int32_t n_latency;                /* len(latency_distribution) */
  float latency_distribution[PUPPER_MAX_LAT];
  int32_t n_imu_latency;
  float imu_latency_distribution[PUPPER_MAX_LAT];
  float desired_world_z_in_body_frame[3];
  int32_t use_imu;
  float reward_scales[PUPPER_NREWARD];
  float tracking_sigma;
  float init_q[PUPPER_NQ];          /* home keyframe with default_pose (environment.py:177,192) */
  float start_pos_min[3], start_pos_max[3]; /* StartPositionRandomization */
  uint32_t knee_sphere_mask;        /* bit s set: sphere s belongs to an upper-leg body */
  uint32_t torso_sphere_mask;       /* bit s set: sphere s belongs to the torso body */
  /* fused brax EpisodeWrapper (only used by pupper_step when PupperEpisode* != NULL) */
  int32_t episode_length;
  int32_t action_repeat;
  int32_t threefry_partitionable;   /* must be 1 (jax 0.5.0 default) */
} PupperEnvCfg;

/* Per-env persistent state, device memory, structure-of-arrays: field f, component c, env e at
 * ptr[c * stride + e] with stride = n_envs padded (caller chooses, >= n_envs, multiple of 32).
 * Updated in place by pupper_step. Field list = SURVEY.md 8(a) E13. */
typedef struct PupperState {
  int32_t stride;
  float *qpos;            /* [19][stride] */
  float *qvel;            /* [18][stride] */
  float *qacc_warmstart;  /* [18][stride] */
  uint32_t *rng;          /* [2][stride]  info["rng"] */
  float *last_act;        /* [12][stride] */
  float *action_buffer;   /* [12*n_latency][stride], element (j, l) at row j*n_latency + l */
  float *imu_buffer;      /* [6*n_imu_latency][stride] */
  float *last_vel;        /* [12][stride] */
  float *command;         /* [3][stride] */
  float *desired_world_z; /* [3][stride] */
  uint32_t *last_contact; /* [stride] bit k = foot k */
  float *feet_air_time;   /* [4][stride] */
  int32_t *step;          /* [stride] info["step"] */
  float *kick;            /* [2][stride] */
  float *obs;             /* [n_envs][H*36] env-major: state.obs, newest observation first */
} PupperState;

/* Domain-randomisation batches, device SoA with the same stride (domain_randomization.py:94-110).
 * Compact form of the reference's DR contract: it draws ONE friction value for every geom, ONE kp
 * and ONE kd multiplier for every actuator, shifts only the base body's COM, and scales every
 * body's inertia (3) and mass. */
typedef struct PupperDR {
  int32_t stride;
  const float *friction;   /* [stride]  geom_friction[:, 0] */
  const float *kp;         /* [stride]  actuator_gainprm[:, 0] (= -biasprm[:, 1]) */
  const float *kd;         /* [stride]  -actuator_biasprm[:, 2] */
  const float *base_ipos;  /* [3][stride]  body_ipos[1] */
  const float *body_inertia; /* [13*3][stride] bodies 1..13 */
  const float *body_mass;  /* [13][stride] bodies 1..13 */
} PupperDR;

/* Step outputs, device, env-major as JAX/Brax hold them. obs lives in PupperState.obs. */
typedef struct PupperStepOut {
  float *reward;   /* [n_envs] */
  float *done;     /* [n_envs] float32 0/1 (environment.py:481) */
  float *metrics;  /* [n_envs][19] total_dist + scaled rewards (= info["rewards"]) */
  /* optional debug taps (NULL to skip): stale forward-pass quantities of the last substep */
  float *dbg_x_pos;      /* [n_envs][13][3] */
  float *dbg_x_rot;      /* [n_envs][13][4] */
  float *dbg_xd_vel;     /* [n_envs][13][3] */
  float *dbg_xd_ang;     /* [n_envs][13][3] */
  float *dbg_qfrc_actuator; /* [n_envs][18] */
  float *dbg_contact_dist;  /* [n_envs][max_contact_points] */
  int32_t *dbg_contact_geom;/* [n_envs][max_contact_points][2] */
  float *dbg_site_xpos;  /* [n_envs][5][3] */
  float *dbg_qacc;       /* [n_envs][18] */
  /* solver decisions of the last substep (oracle: OracleDebug.used_warmstart / ls_iters / ls_alpha / efc_zone0):
   *   [0] 1 if the Newton iteration started from qacc_warmstart, 0 if from qacc_smooth
   *   [1] bracket refinements of the line search this env ran (MJX's loop counter, 0..ls_iterations)
   *   [2] number of active contacts (dist < 0, after the max_contact_points cut)
   *   [3] zone of the 12 friction-loss rows at the start point, 2 bits per hinge dof (1 quadratic, 2 / 3 linear -/+)
   *   [4] bit j: limit row of hinge j active at the start point
   *   [5] bit 4c+e: pyramid edge e of contact slot c active at the start point (slots as in dbg_contact_*)
   *   [6] the accepted step size alpha (float bits)
   *   [7] number of leg-leg contacts among [2] */
  int32_t *dbg_solver;   /* [n_envs][8] */
  /* constraint rows of the last substep as the solver saw them: (efc_D, efc_aref) per row, rows in the order
   *   0-11  friction-loss rows of the 12 hinge dofs,
   *   12-23 joint-limit rows of the 12 hinges (D = 0: limit not violated, no row),
   *   24-43 pyramid edges: 4 per contact slot (n + mu t1, n - mu t1, n + mu t2, n - mu t2), slots as in dbg_contact_*
   *         (D = 0 past the active contacts)
   * (oracle: OracleDebug.efc_D / efc_aref; mjx constraint.py make_constraint) */
  float *dbg_efc;        /* [n_envs][44][2] */
  /* optional (NULL to skip) second destination of the step's observation rows, [n_envs][observation_history * 36], 16-byte
   * aligned: written by pupper_step at the end of the kernel, after the auto-reset block (so it equals state->obs after
   * the call), in 512-byte pieces per warp.  Meant for mapped pinned HOST memory: together with reward / done pointing
   * there too, a host-resident policy gets the step's results without a device-to-host copy being launched after the kernel
   * (the stores travel over PCIe while other CTAs are still computing).  pupper_reset ignores it. */
  float *obs_copy;
} PupperStepOut;

/* External randoms (optional; NULL = every draw is made in-kernel with threefry2x32 from state->rng).
 * u holds RAW uniforms in [0, 1) -- what jax.random.uniform produces before its affine map -- one row per draw; the
 * kernel applies the reference's ranges / thresholds to them (max(lo, u*(hi-lo)+lo), choice via searchsorted).  With
 * external randoms state->rng is neither read nor written, so physics / reward / observation parity can be checked
 * without relying on the restated key tree (draw sites: environment.py:349-361, 499-523, 256-269, 291-293).
 * Rows for pupper_step:  0 kick x, 1 kick y, 2 kick Bernoulli, 3 action-latency choice, 4-6 angular-velocity noise,
 *   7-9 gravity noise, 10-21 motor-angle noise, 22-33 last-action noise, 34 IMU-latency choice, 35-37 command
 *   (lin x, lin y, yaw), 38 zero-command Bernoulli, 39-41 near-zero command, 42 pitch, 43 roll (35-43 are used only in
 *   a step that resamples the command).
 * Rows for pupper_reset: 0-2 start position x, y, z, 3 start yaw; 4-43 as above. */
typedef struct PupperRand {
  int32_t stride;
  const float *u; /* [PUPPER_NRAND][stride] device */
} PupperRand;

/* Fused brax EpisodeWrapper + AutoResetWrapper, device SoA, optional. */
typedef struct PupperEpisode {
  int32_t stride;
  float *first_qpos;      /* [19][stride] snapshot taken by pupper_reset */
  float *first_qvel;      /* [18][stride] */
  float *first_warmstart; /* [18][stride] */
  float *first_obs;       /* [n_envs][H*36] */
  int32_t *steps;         /* [stride] info["steps"] */
  float *truncation;      /* [stride] */
  float *sum_reward;      /* [stride] episode_metrics["sum_reward"] */
  float *length;          /* [stride] */
  float *sum_metrics;     /* [19][stride] */
  float *episode_done;    /* [stride] previous step's done */
  float *totals;          /* [24] device accumulator of COMPLETED episodes, all-reduced across ranks: [0] episodes, [1] sum_reward,
                             [2] length, [3..21] the 19 metric sums, [22] terminations (done without truncation), [23] reserved (0) */
} PupperEpisode;

typedef struct PupperModel PupperModel; /* opaque: device-resident constant tables */
typedef void *pupper_stream_t;          /* cudaStream_t */

int pupper_abi_version(void);
const char *pupper_strerror(int code);
const char *pupper_last_cuda_error(void);

int pupper_model_create(const PupperModelDesc *desc, const PupperEnvCfg *cfg, int device,
                        PupperModel **out);
int pupper_model_destroy(PupperModel *model);

/* keys: device uint32 [n_envs][2] (one JAX PRNG key per env, as vmap(reset) receives them). */
int pupper_reset(const PupperModel *model, int n_envs, const uint32_t *keys, const PupperDR *dr,
                 PupperState *state, PupperStepOut *out, PupperEpisode *episode,
                 const PupperRand *ext_rand, pupper_stream_t stream);

/* action: device float32 [n_envs][12] row-major (as jax.vmap(env.step) receives it).  Every random
 * draw of the step (kick, latency picks, observation noise, command resampling) is made in-kernel
 * with threefry2x32 from state->rng exactly as the reference's key tree.
 * episode != NULL fuses brax EpisodeWrapper + AutoResetWrapper after the env step. */
int pupper_step(const PupperModel *model, int n_envs, const PupperDR *dr, PupperState *state,
                const float *action, const PupperRand *ext_rand, PupperStepOut *out,
                PupperEpisode *episode, pupper_stream_t stream);

/* Measurement helper (bench.py): enqueues an FP32 FMA probe of blocks*256*iters*16 flop, to time with CUDA
 * events for the measured FP32 roofline denominator. device_sink: any device float (never written). */
int pupper_probe_ffma(int blocks, int iters, float *device_sink, pupper_stream_t stream);
int pupper_sizeof(int which); /* 0..6: sizeof PupperModelDesc, EnvCfg, State, DR, StepOut, Episode, Rand (binding self-check) */

/* Number of kernels the last pupper_step / pupper_reset call on this model enqueued. */
int pupper_last_launch_count(const PupperModel *model);

/* Rows of each PupperState SoA field (each row is `stride` 4-byte elements), in declaration order:
 * qpos, qvel, qacc_warmstart, rng, last_act, action_buffer, imu_buffer, last_vel, command,
 * desired_world_z, last_contact, feet_air_time, step, kick  (14 entries; obs is env-major). */
int pupper_state_rows(const PupperEnvCfg *cfg, int32_t *rows_out /* [14] */);

#ifdef __cplusplus
}
#endif
#endif /* PUPPER_ENV_H_ */

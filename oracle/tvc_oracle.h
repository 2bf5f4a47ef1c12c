/*
 * tvc_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Double-precision restatement of the reference hot path
 *   /root/reference/env/enhanced_rocket_tvc_env.py  (EnhancedRocketTVCEnv.step/reset,
 *   MultiObjectiveReward, MissionSuccess, mission phases, termination)
 * plus the slice of PyBullet 3.2.x (Bullet3 btMultiBody, pinned by
 * /root/reference/requirements.txt:11, NOT vendored in the reference) that the
 * env drives.
 *
 * PARITY STATUS: "parity unpinned" for the Bullet rows:
 * PyBullet is not installable in this image and the reference ships no golden
 * vectors.  The env-level rows (S1-S15, R1-R10, Q1-Q22) ARE pinned:
 * tests/golden/make_golden.py runs the reference's own, unmodified Python class on top
 * of oracle/fake_pybullet.py (which forwards the ~18 PyBullet calls to the physics
 * layer below) and the env layer below must reproduce those trajectories.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use this library.  The product (tvc_ai_b200/) never links or calls it.
 */
#ifndef TVC_ORACLE_H
#define TVC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_DELAY 4
#define ORC_HIST 1000
#define ORC_NCOMP 12
#define ORC_NSTATS 16

/* ------------------------------------------------------------------ */
/* Physics layer: the exercised slice of Bullet (rows B1-B9)           */
/* ------------------------------------------------------------------ */

typedef struct orc_body_params {
    double mass;
    double inertia[3];        /* local diagonal */
    double lin_damp, ang_damp;/* Bullet changeDynamics linearDamping/angularDamping (row B5) */
    int32_t use_gyro;         /* btMultiBody::m_useGyroTerm, default false */
    int32_t substeps;         /* numSubSteps */
    double gravity[3];        /* p.setGravity */
    double dt_step;           /* fixedTimeStep */
    double max_vel;           /* m_maxCoordinateVelocity = 100 */
    /* collision geometry + contact model (our documented model) */
    int32_t ground;           /* 0 = no plane */
    int32_t contact_iters;    /* PGS sweeps of the first substep of a step (cold start) */
    int32_t warm_iters;       /* PGS sweeps of the following substeps (warm-started from the previous substep) */
    int32_t reserved_;
    double radius, half_len;
    double cg;                /* COM offset from geometric centre along body z */
    double mu, mu_spin, mu_roll;
    double restitution, rest_threshold, erp, margin;
} orc_body_params;

typedef struct orc_body {
    double pos[3];
    double quat[4];   /* body->world, x y z w, internal (not sign-canonicalised) */
    double vel[3];
    double omega[3];  /* world frame */
    double force[3];  /* accumulated external force, cleared by step */
    double torque[3];
    /* quirk Q3 cleared: a body-fixed force at (0, 0, thrust_arm), re-evaluated at every substep's attitude; cleared by step */
    double thrust_local[3];
    double thrust_arm;
} orc_body;

void orc_body_params_default(orc_body_params *p);
void orc_body_init(orc_body *b, const double pos[3], const double quat[4]);
void orc_apply_external_force(orc_body *b, const double f[3], const double pos_world[3]);
void orc_apply_external_torque(orc_body *b, const double t[3]);
/* one p.stepSimulation(): gravity + K substeps + clearForces.  trace (nullable):
 * K x 13 doubles (pos, quat, vel, omega) after each substep. */
void orc_step_simulation(const orc_body_params *p, orc_body *b, double *trace);
void orc_reported_quat(const double q[4], double out[4]);       /* row B7 */
void orc_euler_from_quat(const double q[4], double rpy[3]);     /* row B8 */
void orc_matrix_from_quat(const double q[4], double m[9]);      /* row B8 */

/* ------------------------------------------------------------------ */
/* Env layer                                                           */
/* ------------------------------------------------------------------ */

enum { ORC_CONTRACT_R = 0, ORC_CONTRACT_X = 1 };
enum { ORC_DIV_OFF = 0, ORC_DIV_FAST = 1, ORC_DIV_EXACT = 2 };

/* quirk switches (SURVEY.md section 8(a) quirk index); same bits and meanings as include/tvc_b200.h TVC_Q_* */
#define ORC_Q_DOUBLE_GRAVITY   (1u << 0)  /* Q1 */
#define ORC_Q_KEEP_CRITERIA    (1u << 1)  /* Q10: criteria history survives reset */
#define ORC_Q_KEEP_REWARD_HIST (1u << 2)  /* Q11: previous_action + reward_history survive reset */
#define ORC_Q_LAGGED_PHASE     (1u << 3)  /* Q8/Q9: obs + R1 use the pre-update phase / success */
#define ORC_Q_THRUST_VECTOR    (1u << 4)  /* Q2; off: normalised direction */
#define ORC_Q_FROZEN_FORCES    (1u << 5)  /* Q3; off: thrust follows the body every substep */
#define ORC_Q_DRAG_CUTOFF      (1u << 6)  /* Q5; off: drag at every speed */
#define ORC_Q_STACKED_DAMPING  (1u << 7)  /* Q6; off: no env damping torque */
#define ORC_Q_EULER_TILT       (1u << 8)  /* Q7; off: angle between body axis and vertical */
#define ORC_Q_DIVERSITY_BONUS  (1u << 9)  /* Q12; off: no bonus */
#define ORC_Q_VARIANCE_PENALTY (1u << 10) /* Q13; off: no penalty */
#define ORC_Q_CLIP_BEFORE_CURIOSITY (1u << 11) /* Q14 (host side) */
#define ORC_Q_SUCCESS_MASKS_TRUNCATION (1u << 12) /* Q16; off: truncation flag also on success */
#define ORC_Q_CRASH_IS_COM_HEIGHT (1u << 13) /* Q17; off: hard or tilted touchdown */
#define ORC_Q_ALL_REFERENCE    (0x3FFFu)
#define ORC_Q_CONTRACT_X       (ORC_Q_ALL_REFERENCE & ~(ORC_Q_KEEP_CRITERIA | ORC_Q_KEEP_REWARD_HIST))

typedef struct orc_config {
    int32_t contract;
    int32_t substeps;
    int32_t max_episode_steps;
    int32_t autoreset;          /* 0 raw gym.Env semantics, 1 same-step autoreset */
    uint32_t quirks;
    int32_t diversity_mode;
    int32_t contact_iters;
    int32_t ground;
    int32_t contact_warm_iters;
    int32_t reserved_;
    double dt_step;
    double gradient_penalty, diversity_bonus;
    /* rocket */
    double mass, radius, length, thrust, gimbal_max_rad;
    double lin_damp, ang_damp;
    /* Contract X */
    double mass_variation;      /* U(1-v, 1+v) */
    double thrust_std, thrust_lo, thrust_hi;
    double cg_offset_max;
    double wind_std;
    double sensor_noise_std;
    double init_tilt_max, init_omega_max;
    double propellant_fraction; /* m = scale*m0*(1 - pf*(1-fuel)) */
    double cg_burn_shift;       /* cg = cg0 + shift*(1-fuel) */
    int32_t delay_steps;        /* actuator delay in control steps (<= ORC_MAX_DELAY) */
    int32_t thrust_curve;       /* 0 constant, 1 model-rocket curve */
    uint64_t seed;
    int64_t env_id_base;
    /* contact material of our ground-contact model */
    double contact_mu, contact_mu_spin, contact_mu_roll, contact_restitution, contact_rest_threshold, contact_erp, contact_margin;
} orc_config;

typedef struct orc_env {
    orc_body body;
    double fuel;            /* fp64 iterated subtraction (row S4) */
    int32_t burn;           /* number of decrements so far */
    int32_t step;
    int32_t phase;          /* 0 boost 1 coast 2 landing 3 touchdown 4 hover 5 complete 6 failed */
    int32_t success;
    int32_t has_prev;
    float prev_action[2];
    int32_t consec;         /* consecutive all-criteria-met pushes (== deque(100) all-true test) */
    int32_t crit_pushes;    /* lifetime pushes into criteria_history (for len>=10 / >=100 tests) */
    int64_t hist_count;     /* lifetime pushes into reward_history */
    double hist[ORC_HIST];  /* ring, slot = push_index % 1000 */
    /* fast diversity bookkeeping (exact for clip duplicates and runs) */
    uint32_t clip_bits[32], run_bits[32];
    int32_t n_clip, n_run;
    double ep_return;
    int32_t episode;
    /* Contract X draws */
    double mass_scale, thrust_scale, cg_offset, wind[2];
    float delay_ring[ORC_MAX_DELAY][2];
} orc_env;

typedef struct orc_step_out {
    float obs[10];
    float final_obs[10];
    double reward;
    int32_t terminated, truncated;
    double comp[ORC_NCOMP]; /* mission, safety, fuel, stability, smooth, altitude, crash, tilt, saturation, adjustment, total_unclipped, diversity_flag */
    /* terminal (pre-autoreset) info */
    double altitude, tilt, omega_mag, fuel, vh, vv;
    double position[3];
    int32_t phase, success, step, criteria_met, term_reason;
    /* diagnostic: distance (m) of the step's closest DISCRETE contact-manifold decision (entry rule, per-point reach rule)
     * from its threshold, 1e30 if the step made none -- the parity reports count steps that sit on such a threshold */
    double contact_margin;
} orc_step_out;

typedef struct orc_sim orc_sim;

void orc_config_default(orc_config *c, int contract);
orc_sim *orc_create(const orc_config *c, int64_t num_envs);
void orc_destroy(orc_sim *s);
int64_t orc_num_envs(const orc_sim *s);
orc_env *orc_env_ptr(orc_sim *s, int64_t i);
orc_config *orc_config_ptr(orc_sim *s);
/* mask nullable (= all).  obs_out: [N,10] float, nullable. */
void orc_reset(orc_sim *s, const uint8_t *mask, float *obs_out);
/* actions [N,2] float.  outs [N] nullable.  threads<=1 -> serial. */
void orc_step(orc_sim *s, const float *actions, orc_step_out *outs, int threads);
/* batched convenience: same as orc_step but SoA outputs (any nullable) */
void orc_step_arrays(orc_sim *s, const float *actions, float *obs, double *reward,
                     uint8_t *terminated, uint8_t *truncated, float *final_obs, int threads);
/* Philox-driven random actions (stream 5), same draws as the device kernel */
void orc_random_actions(const orc_sim *s, int64_t t, float *actions_out);
void orc_stats(orc_sim *s, double out[ORC_NSTATS], int reset_after);
/* substep trace of env 0 for the last orc_step: K x 13 doubles */
const double *orc_last_trace(const orc_sim *s);

/* utilities exposed for known-answer tests */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_fuel_table(int n);
double orc_thrust_curve(int mode, int burn);

#ifdef __cplusplus
}
#endif
#endif

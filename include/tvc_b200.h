/*
 * tvc_b200.h -- C ABI of libtvc_b200.so: the batched, B200-native (sm_100a) replacement for the
 * EnhancedRocketTVCEnv step/reset hot path of NIKHILSAI71/TVC-AI.
 *
 * The reference is pure Python; its "FFI" for this path is the set of PyBullet calls made from
 * env/enhanced_rocket_tvc_env.py.  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference root).  The reference-side binding is a ctypes
 * stub, shown in INTEGRATION.md and implemented in tvc_ai_b200/_abi.py.
 *
 * Conventions
 *  - every function returns 0 on success or a negative TVC_E_* code; tvc_last_error() gives the
 *    message (thread-local).  No exceptions, no exit(), no allocation on the step path.
 *  - *_dev pointers are raw CUDA device pointers owned by the caller (e.g. tensor.data_ptr());
 *    the library owns only the persistent per-env state inside the handle.
 *  - launches are enqueued on the caller's stream and do not synchronise, except where noted.
 *  - a handle is bound to one device; one process per GPU.  Not thread-safe.
 *  - there is NO CPU fallback: tvc_create fails with TVC_E_DEVICE on a non-sm_100 device.
 */
#ifndef TVC_B200_H
#define TVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVC_ABI_VERSION 8

#define TVC_OBS_DIM 10
#define TVC_ACT_DIM 2
#define TVC_NUM_COMPONENTS 12 /* mission, safety, fuel, stability, smoothness, altitude, crash, tilt,
                                 saturation, adjustment, total_unclipped, diversity_flag */
#define TVC_NUM_STATS 16      /* episodes, sum_return, sum_return_sq, sum_length, successes, term_crash,
                                 term_tilt, term_altitude, term_range, truncations, safety_violations,
                                 sum_final_altitude, sum_final_tilt, sum_fuel_left, steps, reserved */
#define TVC_MAX_DELAY 4

enum {
    TVC_OK = 0,
    TVC_E_BADARG = -1,
    TVC_E_CUDA = -2,
    TVC_E_DEVICE = -3, /* not an sm_100 device */
    TVC_E_ABI = -4,
    TVC_E_NOMEM = -5,
    TVC_E_STATE = -6
};

enum { TVC_CONTRACT_R = 0 /* reference-faithful */, TVC_CONTRACT_X = 1 /* extension */ };
enum { TVC_DIV_OFF = 0, TVC_DIV_FAST = 1, TVC_DIV_EXACT = 2 };

/* Quirk switches, SURVEY.md section 8(a) quirk index (lines of env/enhanced_rocket_tvc_env.py).  A set bit reproduces
 * the reference's behaviour; the comment says what the cleared bit does instead.  Contract R sets all of them.
 * Contract X keeps the reference's physics and reward conventions (so that numbers stay comparable and policies
 * transfer) and clears only the two cross-episode leaks Q10 / Q11, which are meaningless under same-step autoreset.
 * Quirks without a bit: Q4 (fuel by iterated fp64 subtraction -- reproduced exactly by the integer burn counter), Q15
 * (deterministic reset: Contract R has no randomness; Contract X draws per-episode conditions), Q18-Q22 (Python-level
 * typing / info quirks, mirrored by the facade in tvc_ai_b200/env.py). */
#define TVC_Q_DOUBLE_GRAVITY   (1u << 0)  /* Q1  :338 + :525-527; off: world gravity only */
#define TVC_Q_KEEP_CRITERIA    (1u << 1)  /* Q10 :300, :399-401; off: reset() clears the success-criteria history */
#define TVC_Q_KEEP_REWARD_HIST (1u << 2)  /* Q11 :301, :82, :172-178; off: reset() clears previous_action / reward_history */
#define TVC_Q_LAGGED_PHASE     (1u << 3)  /* Q8/Q9 :481-482 vs :485-493; off: obs and R1 see this step's phase / success */
#define TVC_Q_THRUST_VECTOR    (1u << 4)  /* Q2  :537-543 thrust 35 [sin yaw, sin pitch, cos pitch cos yaw] (not unit norm);
                                                  off: the same direction normalised, |F| = thrust */
#define TVC_Q_FROZEN_FORCES    (1u << 5)  /* Q3  :474-477 forces held constant in the world frame over the substeps;
                                                  off: thrust force and torque follow the body attitude every substep */
#define TVC_Q_DRAG_CUTOFF      (1u << 6)  /* Q5  :568-580 no drag below |v| = 0.1 m/s; off: drag at every speed */
#define TVC_Q_STACKED_DAMPING  (1u << 7)  /* Q6  :583-585 env torque -0.02 rho w on top of Bullet's damping; off: Bullet's only */
#define TVC_Q_EULER_TILT       (1u << 8)  /* Q7  :614-616 tilt = sqrt(pitch^2 + yaw^2) of the ZYX Euler angles;
                                                  off: the angle between the body axis and the vertical */
#define TVC_Q_DIVERSITY_BONUS  (1u << 9)  /* Q12 :220-224; off: no diversity bonus (as diversity_mode = TVC_DIV_OFF) */
#define TVC_Q_VARIANCE_PENALTY (1u << 10) /* Q13 :213-218; off: no -0.1 var(last 10) penalty */
#define TVC_Q_CLIP_BEFORE_CURIOSITY (1u << 11) /* Q14 :121-123 vs :495-502 (host side: the facades add the curiosity term);
                                                  off: the sum including the curiosity term is clipped to [-1000, 200] */
#define TVC_Q_SUCCESS_MASKS_TRUNCATION (1u << 12) /* Q16 :703-719; off: truncated = step >= max_steps also on success */
#define TVC_Q_CRASH_IS_COM_HEIGHT (1u << 13) /* Q17 :632 crashed = z_com < 0.1; off: the lowest point of the body is within
                                                  1 cm of the ground while sinking faster than 2 m/s or tilted beyond 0.52 rad */
#define TVC_Q_ALL_REFERENCE    0x3FFFu
#define TVC_Q_CONTRACT_X       (TVC_Q_ALL_REFERENCE & ~(TVC_Q_KEEP_CRITERIA | TVC_Q_KEEP_REWARD_HIST))

#define TVC_SEED_KEEP UINT64_MAX /* tvc_reset: keep the current Philox key (every other value, 0 included, is a seed) */

typedef struct tvc_handle tvc_handle;
typedef void *tvc_stream; /* cudaStream_t */

/* Everything the reference hard-codes (enhanced_rocket_tvc_env.py:409-464) or, for Contract X,
 * leaves in YAML nobody reads (config/config.yaml:335-349). */
typedef struct tvc_config {
    int32_t abi_version; /* must be TVC_ABI_VERSION */
    int32_t contract;
    int32_t substeps;          /* R: 4 (:341) */
    int32_t max_episode_steps; /* :282 */
    int32_t autoreset;         /* 0: gym.Env semantics; 1: same-step autoreset (VectorEnv) */
    uint32_t quirks;
    int32_t diversity_mode;
    int32_t contact_iters;      /* block Gauss-Seidel passes of a solve that follows free flight (cold start) */
    int32_t ground;
    int32_t delay_steps;  /* X: actuator delay, control steps */
    int32_t thrust_curve; /* X: 0 constant, 1 model-rocket curve */
    int32_t contact_warm_iters; /* passes of a solve that follows a solve (warm-started from the previous substep) */
    double dt_step;       /* :340 */
    float gradient_penalty, diversity_bonus; /* :83-84 */
    float mass, radius, length, thrust;      /* :412-414, :463 */
    float gimbal_max_rad;                    /* :471 */
    float lin_damp, ang_damp;                /* :453-454 */
    /* Contract X domain randomisation (config/config.yaml:344-349) */
    float mass_variation;
    float thrust_std, thrust_lo, thrust_hi;
    float cg_offset_max;
    float wind_std;
    float sensor_noise_std;
    float init_tilt_max, init_omega_max;
    float propellant_fraction, cg_burn_shift;
    float reserved1;
    /* contact material of our ground-contact model; defaults = Bullet's combination of
     * :349-352 (plane) and :455-458 (rocket) */
    float contact_mu;           /* 0.3 * 0.8 */
    float contact_mu_spin;      /* 0.1 * 0.8 + 0.1 * 0.3 */
    float contact_mu_roll;      /* 0.05 * 0.8 + 0.05 * 0.3 */
    float contact_restitution;  /* 0.1 */
    float contact_rest_threshold; /* 0.2 m/s */
    float contact_erp;          /* 0.2 */
    float contact_margin;       /* 0.05 m */
    float reserved2;
    uint64_t seed;
    int64_t env_id_base; /* global id of env 0 of this handle: Philox counters use global ids, so
                            results do not depend on how envs are sharded over GPUs */
} tvc_config;

/* scripts/curriculum_manager.py:76-94 `conditions` dict */
typedef struct tvc_stage_conditions {
    float max_initial_tilt;
    float max_initial_angular_vel;
    int32_t domain_randomization;
    int32_t sensor_noise;
    float max_gimbal_angle_deg; /* <= 0: keep the env's 18 deg */
    int32_t wind_enabled;
    float wind_force;
    float mass_variation;
} tvc_stage_conditions;

/* Optional per-env info written by tvc_step_ex / tvc_read_info (enhanced_rocket_tvc_env.py:723-742).
 * Any member may be NULL. */
typedef struct tvc_info_soa {
    float *altitude;
    float *tilt_deg;
    float *omega_mag;
    float *fuel;
    float *position; /* [N,3] */
    int32_t *phase;  /* MissionPhase index, :21-29 */
    int32_t *step;
    uint8_t *success;
    uint8_t *criteria_met;
    float *reward_components; /* [N, TVC_NUM_COMPONENTS] (step_ex only) */
} tvc_info_soa;

typedef struct tvc_step_io {
    const float *actions;   /* [N,2]; NULL => Philox actions U(-1,1), stream 5, counter = lifetime step */
    float *obs;             /* [N,10] post-autoreset observation */
    float *reward;          /* [N] */
    uint8_t *terminated;    /* [N] */
    uint8_t *truncated;     /* [N] */
    float *final_obs;       /* [N,10] nullable; written only where terminated|truncated */
    float *actions_out;     /* [N,2] nullable; the actions actually applied (Philox mode) */
    tvc_info_soa info;      /* terminal (pre-autoreset) info */
} tvc_step_io;

/* Portable per-env state blob for tvc_get_state / tvc_set_state (parity tests, checkpoints).  Restoring it into a
 * fresh handle continues the run bit for bit: it carries the diversity window of TVC_DIV_FAST (the two 1000-bit
 * rings behind n_clip / n_run); the 1000-value window of TVC_DIV_EXACT travels through tvc_get/set_reward_history. */
typedef struct tvc_env_state {
    float pos[3];
    float quat[4];
    float vel[3];
    float omega[3];
    float prev_action[2];
    float ep_return;
    int32_t step, burn, phase, success, has_prev, consec;
    int32_t hist_count, episode;
    int32_t n_clip, n_run;
    float ring10[10]; /* slot = push_index % 10 */
    float mass_scale, thrust_scale, cg_offset, wind[2];
    float delay_ring[TVC_MAX_DELAY][2];
    uint32_t clip_bits[32], run_bits[32]; /* TVC_DIV_FAST: bit p % 1000 of push p */
} tvc_env_state;

/* SAC actor for the fused rollout (config 4): Linear(10,256)-ReLU-Linear(256,256)-ReLU-Linear(256,4).
 * Row-major [out,in] float32 device arrays (torch nn.Linear layout); converted to bf16 inside. */
typedef struct tvc_actor_weights {
    const float *w1, *b1; /* [256,10], [256] */
    const float *w2, *b2; /* [256,256], [256] */
    const float *w3, *b3; /* [4,256], [4] : mean(2), log_std(2) */
} tvc_actor_weights;

typedef struct tvc_rollout_io {
    float *obs;          /* [N,10] REQUIRED, in/out: the observation returned by the last reset/step/rollout; updated */
    float *reward_sum;   /* [N] nullable: sum of rewards over the T steps */
    float *actions_last; /* [N,2] nullable */
    float *actions_all;  /* [T,N,2] nullable */
    float *reward_all;   /* [T,N] nullable */
    int32_t deterministic; /* 1: a = tanh(mean) */
    int32_t reserved;
    /* transition record for an on-device replay buffer; all nullable */
    float *obs_all;        /* [T,N,10] observation the actor saw at step t */
    float *next_obs_all;   /* [T,N,10] successor observation (the terminal one when the episode ended at step t) */
    uint8_t *terminated_all; /* [T,N] */
    uint8_t *truncated_all;  /* [T,N] */
} tvc_rollout_io;

int tvc_abi_version(void);
const char *tvc_last_error(void);

/* defaults == what enhanced_rocket_tvc_env.py hard-codes (R) / config.yaml:335-349 (X) */
int tvc_config_default(tvc_config *cfg, int contract);

/* replaces EnhancedRocketTVCEnv.__init__ + _setup_physics (:279-352) for N envs */
int tvc_create(const tvc_config *cfg, int device, int64_t num_envs, tvc_handle **out);
/* replaces close() (:749-753) */
int tvc_destroy(tvc_handle *h);

/* replaces reset() (:381-407).  mask_dev NULL = all envs.  seed != TVC_SEED_KEEP re-keys the Philox streams
 * (Contract X); in Contract R it has no effect on dynamics (quirk Q15). */
int tvc_reset(tvc_handle *h, const uint8_t *mask_dev, uint64_t seed, float *obs_out_dev, tvc_stream stream);

/* replaces step() (:466-518) for N envs: clip -> control -> K substeps -> obs/phase/success/
 * reward/termination (+ same-step autoreset when cfg.autoreset). */
int tvc_step(tvc_handle *h, const float *actions_dev, float *obs_dev, float *reward_dev,
             uint8_t *terminated_dev, uint8_t *truncated_dev, float *final_obs_dev, tvc_stream stream);
int tvc_step_ex(tvc_handle *h, const tvc_step_io *io, tvc_stream stream);

/* Same step through HOST buffers: H2D of actions, launch, D2H of results, stream sync.
 * This is the end-to-end call the Python VectorEnv facade makes for numpy inputs. */
int tvc_step_host(tvc_handle *h, const float *actions_host, float *obs_host, float *reward_host,
                  uint8_t *terminated_host, uint8_t *truncated_host, float *final_obs_host);
/* The same call split in two: enqueue (copies and kernels go to the handle's own stream, nothing is waited for; the
 * host buffers must stay untouched until the sync) and wait.  Several handles -- env slabs of one VectorEnv -- enqueued
 * back to back overlap one slab's device-to-host copy with the next slab's kernels (RocketTVCHostPipelineEnv). */
int tvc_step_host_async(tvc_handle *h, const float *actions_host, float *obs_host, float *reward_host,
                        uint8_t *terminated_host, uint8_t *truncated_host, float *final_obs_host);
int tvc_host_sync(tvc_handle *h);

/* Fused rollout: T env steps per launch with the SAC actor MLP evaluated inside the loop
 * (replaces train.py:546-603's get_action -> step loop for the legacy 2x256 actor). */
int tvc_rollout(tvc_handle *h, const tvc_actor_weights *w, int32_t T, const tvc_rollout_io *io, tvc_stream stream);

/* Row S14 for the whole batch: the intrinsic-curiosity term of enhanced_rocket_tvc_env.py:494-506 with the forward model of
 * :226-269 (Linear(10,256)-ReLU-Linear(256,256)-ReLU-Linear(256,8) over [state8, action2]; never trained -- quirk Q19) on the
 * tensor cores (bf16 operands, fp32 accumulation; csrc/tvc_curiosity.cu).  Call it after tvc_step with that step's actions
 * and results.  Per env: intrinsic = has_prev ? 0.01 * mean((forward([prev_state, clip(action)]) - next8)^2) : 0 with
 * next8 = the step's own successor state (final_obs[:8] where the episode ended, else obs[:8]); reward_out = reward_in +
 * intrinsic (quirk Q14: after the env's clip; clip_sum = 1 clips the sum instead); then prev_state <- obs[:8] and
 * has_prev <- !done (reset() clears state_history, :399-401).  w == NULL reuses the weights packed by an earlier call. */
typedef struct tvc_forward_model {
    const float *w1, *b1; /* [256,10], [256]   (torch nn.Linear layout [out,in]) */
    const float *w2, *b2; /* [256,256], [256] */
    const float *w3, *b3; /* [8,256], [8] */
} tvc_forward_model;

typedef struct tvc_curiosity_io {
    const float *actions;      /* [N,2] REQUIRED */
    const float *obs;          /* [N,10] REQUIRED: what tvc_step returned */
    const float *final_obs;    /* [N,10] nullable (same-step autoreset: terminal observations) */
    const uint8_t *terminated; /* [N] nullable */
    const uint8_t *truncated;  /* [N] nullable */
    float *prev_state;         /* [N,8] REQUIRED, in/out: state_history[-1] */
    uint8_t *has_prev;         /* [N] REQUIRED, in/out: len(state_history) > 0 */
    const float *reward_in;    /* [N] nullable */
    float *reward_out;         /* [N] nullable; may alias reward_in */
    float *intrinsic;          /* [N] nullable: reward_components['curiosity'] */
    int32_t clip_sum;          /* 0: quirk Q14 (reference) */
    int32_t reserved;
} tvc_curiosity_io;
int tvc_curiosity(tvc_handle *h, const tvc_forward_model *w, const tvc_curiosity_io *io, tvc_stream stream);

/* On-device replay ring (replaces train.py:574-591's batch-of-1 agent.update feed).  The ring
 * is caller memory; tvc_rollout fills it in place when tvc_rollout_io.obs_all / actions_all / reward_all / next_obs_all /
 * terminated_all point at its head (T * N consecutive transitions).  tvc_replay_sample draws `batch` uniform indices in
 * [0, filled) from Philox4x32-10 (key = seed, counter = (sample, draw)) and gathers the transitions into the learner's batch
 * tensors in one launch: reward * reward_scale, done = terminated as float (agent/multi_algorithm_agent.py:950-1016 inputs). */
typedef struct tvc_replay_ring {
    const float *obs;          /* [capacity,10] */
    const float *actions;      /* [capacity,2] */
    const float *reward;       /* [capacity] */
    const float *next_obs;     /* [capacity,10] */
    const uint8_t *terminated; /* [capacity] */
    int64_t capacity;
} tvc_replay_ring;
typedef struct tvc_replay_batch {
    float *obs, *actions, *reward, *next_obs, *done; /* [batch,10], [batch,2], [batch], [batch,10], [batch] */
    int64_t *indices;                                 /* [batch] nullable: the ring indices drawn */
} tvc_replay_batch;
/* ctl_dev (nullable): two device words {filled, draw_base}; when given, the kernel reads `filled` from it and adds draw_base to
 * `draw`, so that a launch captured into a CUDA graph follows the ring as it fills and never repeats a draw. */
int tvc_replay_sample(const tvc_replay_ring *ring, int64_t filled, int32_t batch, uint64_t seed, uint64_t draw,
                      const uint64_t *ctl_dev, float reward_scale, const tvc_replay_batch *out, int device, tvc_stream stream);

/* state exchange; dev_blob = N x tvc_env_state on the device */
size_t tvc_state_bytes(const tvc_handle *h);
int tvc_get_state(tvc_handle *h, void *dev_blob, size_t bytes, tvc_stream stream);
int tvc_set_state(tvc_handle *h, const void *dev_blob, size_t bytes, tvc_stream stream);
/* TVC_DIV_EXACT only: the per-env 1000-value reward window, [N][1000] float on the device (slot = push % 1000) */
int tvc_get_reward_history(tvc_handle *h, float *dev_out, size_t bytes, tvc_stream stream);
int tvc_set_reward_history(tvc_handle *h, const float *dev_in, size_t bytes, tvc_stream stream);

/* _get_enhanced_info (:723-742) for the current state */
int tvc_read_info(tvc_handle *h, const tvc_info_soa *dev, tvc_stream stream);

/* episode statistics (what train.py:608-618 consumes).  Synchronises `stream`. */
int tvc_episode_stats(tvc_handle *h, double *host_out, int reset_after, tvc_stream stream);
/* device-side reduction into a caller buffer of TVC_NUM_STATS doubles (feed to ncclAllReduce) */
int tvc_episode_stats_dev(tvc_handle *h, double *dev_out, int reset_after, tvc_stream stream);

/* curriculum_manager.py:191-246 decides the stage on the host; this applies its conditions */
int tvc_set_curriculum(tvc_handle *h, const tvc_stage_conditions *c);

int tvc_get_config(const tvc_handle *h, tvc_config *out);
int64_t tvc_num_envs(const tvc_handle *h);
int64_t tvc_lifetime_steps(const tvc_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* TVC_B200_H */

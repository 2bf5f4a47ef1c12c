// tvc_internal.h -- host-side handle shared by the translation units of libtvc_b200.so
#pragma once
#include "../../include/tvc_b200.h"
#include "tvc_device.cuh"

#include <cstring>
#include <string>

struct tvc_handle {
    int device = 0;
    int num_sms = 0;
    int64_t n = 0;
    int grid = 0;       // CTAs of TVC_BLOCK envs (legacy step kernel, rollout kernel)
    int ngroups = 0;    // 32-env groups (step_kernel_v2 work items, statistics rows)
    int v2_grid = 0;    // persistent grid of step_kernel_v2 (computed on first launch)
    bool order_valid = false;  // the sorted sequence describes the current state (false after reset / set_state / rollout / curriculum)
    bool pdl = true;         // programmatic dependent launch of step_kernel_v2 and of the closing sort kernel
    tvc_config base;   // as created
    tvc_config cur;    // after tvc_set_curriculum
    tvc::DevCfg dc;
    tvc::DevState st;
    int64_t lifetime_steps = 0;  // env steps taken by every env of this handle (Philox action counter)
    double *stats_dev = nullptr;
    double *stats_host = nullptr;  // pinned
    cudaStream_t own_stream = nullptr;
    // staging for tvc_step_host
    float *io_act = nullptr, *io_obs = nullptr, *io_rew = nullptr, *io_final = nullptr;
    uint8_t *io_term = nullptr, *io_trunc = nullptr;
    bool host_pending = false;                // tvc_step_host_async enqueued, tvc_host_sync not yet called
    float *act_pinned = nullptr;              // pinned staging for pageable action buffers (tvc_step_host)
    const float *final_host_seen = nullptr;   // last final_obs_host pointer classified by tvc_step_host
    float *final_host_dev = nullptr;          // its device alias when it is pinned (mapped) host memory, else NULL
    // fused-rollout workspace (tvc_rollout.cu)
    void *rollout_ws = nullptr;
    // packed forward-model weights of tvc_curiosity (tvc_curiosity.cu)
    void *curiosity_ws = nullptr;
};

// tvc_config (ABI) -> the constants the device code reads
inline void make_devcfg(const tvc_config &c, tvc::DevCfg &d) {
    memset(&d, 0, sizeof(d));
    d.contract = c.contract; d.K = c.substeps; d.max_steps = c.max_episode_steps; d.autoreset = c.autoreset;
    d.quirks = c.quirks; d.div_mode = c.diversity_mode; d.contact_iters = c.contact_iters; d.warm_iters = c.contact_warm_iters; d.ground = c.ground;
    d.delay = c.delay_steps; d.thrust_curve = c.thrust_curve;
    const double dt = c.dt_step / (double)c.substeps;
    d.dt = (float)dt; d.inv_dt = (float)(1.0 / dt);
    d.inv_max_steps = (float)(1.0 / (double)c.max_episode_steps);
    d.gp = c.gradient_penalty; d.db = c.diversity_bonus;
    d.mass = c.mass; d.radius = c.radius; d.half_len = 0.5f * c.length; d.thrust = c.thrust; d.gimbal_max = c.gimbal_max_rad;
    d.lin_damp = c.lin_damp; d.ang_damp = c.ang_damp;
    d.mass_var = c.mass_variation; d.thrust_std = c.thrust_std; d.thrust_lo = c.thrust_lo; d.thrust_hi = c.thrust_hi;
    d.cg_max = c.cg_offset_max; d.wind_std = c.wind_std; d.noise_std = c.sensor_noise_std;
    d.tilt_max = c.init_tilt_max; d.omega_max = c.init_omega_max; d.prop_frac = c.propellant_fraction; d.cg_burn = c.cg_burn_shift;
    d.mu = c.contact_mu; d.mu_spin = c.contact_mu_spin; d.mu_roll = c.contact_mu_roll;
    d.restitution = c.contact_restitution; d.rest_thr = c.contact_rest_threshold; d.erp = c.contact_erp; d.margin = c.contact_margin;
    d.seed_lo = (unsigned)c.seed; d.seed_hi = (unsigned)(c.seed >> 32);
    d.env_base = c.env_id_base;
}

const char *tvc_set_err(const std::string &m);
void tvc_rollout_free(tvc_handle *h);
void tvc_curiosity_free(tvc_handle *h);

// tvc_internal.h -- host-side handle shared by the translation units of libtvc_b200.so
#pragma once
#include "../../include/tvc_b200.h"
#include "tvc_device.cuh"

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
    bool v2_defer = false;   // large batches: finished envs are reset by the closing sort kernel; small ones in place
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
};

const char *tvc_set_err(const std::string &m);
void tvc_rollout_free(tvc_handle *h);

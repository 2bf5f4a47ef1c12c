// twin.cpp -- TEST INFRASTRUCTURE ONLY: the device-side model (tvc_ai_b200/csrc/tvc_device.cuh) compiled by g++ for the host.
//
// One "thread" per env in a plain loop, the same per-env call sequence as step_kernel_v2 (load_env -> env_pre ->
// integrate_thread -> env_post -> in-place autoreset -> store_env), on host arrays laid out like the device planes.  It lets
// the CPU test-suite compare the kernels' own source with the fp64 oracle (tests/test_host_twin.py) -- what it cannot see is
// nvcc's FMA contraction and the MUFU approximations, which the -m gpu tests cover.  Nothing in the product links this.
#include "cuda_host_shim.h"

#include "../../tvc_ai_b200/csrc/tvc_internal.h"

#include <vector>

using namespace tvc;

struct twin_handle {
    tvc_config cfg;
    DevCfg dc;
    DevState st;
    long long n;
    std::vector<float4> s0, s1, s2, s3, s4, d0, d1;
    std::vector<float> ring, hist;
    std::vector<unsigned> clipb, runb;
    std::vector<float2> delay;
    unsigned long long t = 0;
};

template <bool X, int DIV, bool FOLLOW>
static void step_all(twin_handle *h, const float *actions, float *obs, float *reward, uint8_t *term, uint8_t *trunc, float *final_obs,
                     float *comp) {
    const DevCfg &c = h->dc;
    const DevState &st = h->st;
    for (long long i = 0; i < h->n; i++) {
        const long long gid = c.env_base + i;
        Env e;
        BodyP P;
        Forces f;
        load_env(st, X, i, e);
        float a0, a1;
        if (actions) { a0 = actions[2 * i]; a1 = actions[2 * i + 1]; }
        else {
            uint4 rr = philox(c.seed_lo, c.seed_hi, gid, ST_ACTION, (unsigned)h->t, (unsigned)(h->t >> 32));
            a0 = 2.0f * u01(rr.x) - 1.0f; a1 = 2.0f * u01(rr.y) - 1.0f;
        }
        env_pre<X>(c, st, i, e, a0, a1, P, f);
        integrate_thread<false, FOLLOW>(c, P, e, f);
        StepResult r;
        env_post<X, DIV>(c, st, i, gid, e, f.a0, f.a1, r);
        reward[i] = r.reward; term[i] = (uint8_t)r.terminated; trunc[i] = (uint8_t)r.truncated;
        if (comp) for (int k = 0; k < 12; k++) comp[12 * i + k] = r.comp[k];
        if (r.terminated | r.truncated) {
            if (final_obs) for (int k = 0; k < 10; k++) final_obs[10 * i + k] = r.obs[k];
            if (c.autoreset) { reset_env(c, X, gid, e, false); build_obs(c, X, gid, e, 0, r.obs); }
        }
        store_env(st, X, i, e);
        for (int k = 0; k < 10; k++) obs[10 * i + k] = r.obs[k];
    }
    h->t++;
}

// tvc_create runs reset_kernel with first_time = 1 (episode 0); tvc_reset re-initialises from the stored env (episode + 1)
static void reset_all(twin_handle *h, float *obs, bool first_time) {
    const bool X = h->dc.contract == TVC_CONTRACT_X;
    for (long long i = 0; i < h->n; i++) {
        Env e;
        if (first_time) { memset(&e, 0, sizeof(e)); e.episode = -1; }
        else load_env(h->st, X, i, e);
        reset_env(h->dc, X, h->dc.env_base + i, e, first_time);
        store_env(h->st, X, i, e);
        if (obs) build_obs(h->dc, X, h->dc.env_base + i, e, 0, obs + 10 * i);
    }
}
extern "C" {

twin_handle *twin_create(const tvc_config *cfg, long long n) {
    twin_handle *h = new twin_handle();
    h->cfg = *cfg; h->n = n;
    make_devcfg(*cfg, h->dc);
    const size_t N = (size_t)n;
    h->s0.assign(N, float4{}); h->s1.assign(N, float4{}); h->s2.assign(N, float4{}); h->s3.assign(N, float4{}); h->s4.assign(N, float4{});
    h->d0.assign(N, float4{}); h->d1.assign(N, float4{});
    h->ring.assign(12 * N, 0.0f); h->hist.assign((size_t)TVC_HIST * N, 0.0f);
    h->clipb.assign(32 * N, 0u); h->runb.assign(32 * N, 0u); h->delay.assign((size_t)TVC_MAX_DELAY * N, float2{});
    memset(&h->st, 0, sizeof(h->st));
    h->st.s0 = h->s0.data(); h->st.s1 = h->s1.data(); h->st.s2 = h->s2.data(); h->st.s3 = h->s3.data(); h->st.s4 = h->s4.data();
    h->st.d0 = h->d0.data(); h->st.d1 = h->d1.data(); h->st.ring = h->ring.data(); h->st.hist = h->hist.data();
    h->st.clipb = h->clipb.data(); h->st.runb = h->runb.data(); h->st.delay = h->delay.data(); h->st.n = n;
    reset_all(h, nullptr, true);
    return h;
}

void twin_destroy(twin_handle *h) { delete h; }

void twin_reset(twin_handle *h, float *obs) { reset_all(h, obs, false); }

// rigid-body state in / out (the teacher-forcing hook of the parity tests): pos3 quat4 vel3 omega3 per env
void twin_set_body(twin_handle *h, long long i, const float *b13) {
    h->s0[i].x = b13[0]; h->s0[i].y = b13[1]; h->s0[i].z = b13[2];
    h->s1[i] = make_float4(b13[3], b13[4], b13[5], b13[6]);
    h->s2[i].x = b13[7]; h->s2[i].y = b13[8]; h->s2[i].z = b13[9];
    h->s3[i].x = b13[10]; h->s3[i].y = b13[11]; h->s3[i].z = b13[12];
}
void twin_get_body(twin_handle *h, long long i, float *b13) {
    b13[0] = h->s0[i].x; b13[1] = h->s0[i].y; b13[2] = h->s0[i].z;
    b13[3] = h->s1[i].x; b13[4] = h->s1[i].y; b13[5] = h->s1[i].z; b13[6] = h->s1[i].w;
    b13[7] = h->s2[i].x; b13[8] = h->s2[i].y; b13[9] = h->s2[i].z;
    b13[10] = h->s3[i].x; b13[11] = h->s3[i].y; b13[12] = h->s3[i].z;
}

// the portable state blob of the ABI (tvc_env_state), same field mapping as set_state_kernel / get_state_kernel
void twin_set_state(twin_handle *h, const tvc_env_state *in) {
    const bool X = h->dc.contract == TVC_CONTRACT_X;
    for (long long i = 0; i < h->n; i++) {
        const tvc_env_state &s = in[i];
        Env e;
        e.px = s.pos[0]; e.py = s.pos[1]; e.pz = s.pos[2]; e.ep_ret = s.ep_return;
        e.qx = s.quat[0]; e.qy = s.quat[1]; e.qz = s.quat[2]; e.qw = s.quat[3];
        const float d = e.qx * e.qx + e.qy * e.qy + e.qz * e.qz + e.qw * e.qw;
        if (fabsf(d - 1.0f) > 1e-6f && d > 0.0f) { const float sc = rsqrtf(d); e.qx *= sc; e.qy *= sc; e.qz *= sc; e.qw *= sc; }
        e.vx = s.vel[0]; e.vy = s.vel[1]; e.vz = s.vel[2]; e.step = s.step;
        e.wx = s.omega[0]; e.wy = s.omega[1]; e.wz = s.omega[2];
        e.burn = s.burn; e.phase = s.phase; e.success = s.success; e.has_prev = s.has_prev; e.consec = s.consec;
        e.ap0 = s.prev_action[0]; e.ap1 = s.prev_action[1];
        e.hist_count = s.hist_count; e.n_clip = s.n_clip; e.n_run = s.n_run;
        e.mass_scale = s.mass_scale; e.thrust_scale = s.thrust_scale; e.cg_off = s.cg_offset;
        e.wind_x = s.wind[0]; e.wind_y = s.wind[1]; e.episode = s.episode;
        store_env(h->st, X, i, e);
        for (int k = 0; k < 10; k++) h->ring[12 * i + k] = s.ring10[k];
        for (int k = 0; k < h->dc.delay; k++) h->delay[(size_t)k * h->n + i] = make_float2(s.delay_ring[k][0], s.delay_ring[k][1]);
        for (int k = 0; k < 32; k++) { h->clipb[(size_t)k * h->n + i] = s.clip_bits[k]; h->runb[(size_t)k * h->n + i] = s.run_bits[k]; }
    }
}
void twin_get_state(twin_handle *h, tvc_env_state *out) {
    const bool X = h->dc.contract == TVC_CONTRACT_X;
    for (long long i = 0; i < h->n; i++) {
        Env e;
        load_env(h->st, X, i, e);
        tvc_env_state s;
        memset(&s, 0, sizeof(s));
        s.pos[0] = e.px; s.pos[1] = e.py; s.pos[2] = e.pz;
        s.quat[0] = e.qx; s.quat[1] = e.qy; s.quat[2] = e.qz; s.quat[3] = e.qw;
        s.vel[0] = e.vx; s.vel[1] = e.vy; s.vel[2] = e.vz;
        s.omega[0] = e.wx; s.omega[1] = e.wy; s.omega[2] = e.wz;
        s.prev_action[0] = e.ap0; s.prev_action[1] = e.ap1; s.ep_return = e.ep_ret;
        s.step = e.step; s.burn = e.burn; s.phase = e.phase; s.success = e.success; s.has_prev = e.has_prev;
        s.consec = e.consec; s.hist_count = e.hist_count; s.episode = e.episode; s.n_clip = e.n_clip; s.n_run = e.n_run;
        for (int k = 0; k < 10; k++) s.ring10[k] = h->ring[12 * i + k];
        s.mass_scale = e.mass_scale; s.thrust_scale = e.thrust_scale; s.cg_offset = e.cg_off; s.wind[0] = e.wind_x; s.wind[1] = e.wind_y;
        for (int k = 0; k < h->dc.delay; k++) { const float2 d = h->delay[(size_t)k * h->n + i]; s.delay_ring[k][0] = d.x; s.delay_ring[k][1] = d.y; }
        for (int k = 0; k < 32; k++) { s.clip_bits[k] = h->clipb[(size_t)k * h->n + i]; s.run_bits[k] = h->runb[(size_t)k * h->n + i]; }
        out[i] = s;
    }
}
void twin_set_step_counter(twin_handle *h, unsigned long long t) { h->t = t; }

int twin_step(twin_handle *h, const float *actions, float *obs, float *reward, uint8_t *term, uint8_t *trunc, float *final_obs,
              float *comp) {
    const bool X = h->dc.contract == TVC_CONTRACT_X, follow = !(h->dc.quirks & Q_FROZEN_FORCES);
    const int div = h->dc.div_mode;
#define GO(XX, DD, FF) step_all<XX, DD, FF>(h, actions, obs, reward, term, trunc, final_obs, comp)
#define GO_D(XX, FF) (div == 0 ? GO(XX, 0, FF) : (div == 1 ? GO(XX, 1, FF) : GO(XX, 2, FF)))
    if (X) { if (follow) GO_D(true, true); else GO_D(true, false); }
    else { if (follow) GO_D(false, true); else GO_D(false, false); }
    return 0;
}

}  // extern "C"

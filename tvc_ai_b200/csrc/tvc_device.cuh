// tvc_device.cuh -- device-side model of the TVC env step (fp32, one thread per env).
//
// Restates, for sm_100a, the reference hot path env/enhanced_rocket_tvc_env.py (step :466-518,
// reward :86-224, phases :635-657, success :659-695, termination :697-721, reset :381-464) and
// the slice of Bullet it drives.  "ref:" below means that
// file.  The fp64 CPU oracle (oracle/tvc_oracle.c) restates the same rows independently; the
// two are compared by tests/ -- this file never includes or calls the oracle.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define TVC_BLOCK 128
#define TVC_WARPS (TVC_BLOCK / 32)
#define TVC_HIST 1000
#define TVC_NSTAT 16

namespace tvc {

struct DevCfg {
    int contract, K, max_steps, autoreset;
    unsigned quirks;
    int div_mode, contact_iters, warm_iters, ground, delay, thrust_curve;
    float dt, inv_dt;  // substep
    float inv_max_steps;
    float gp, db;
    float mass, radius, half_len, thrust, gimbal_max;
    float lin_damp, ang_damp;
    float mass_var, thrust_std, thrust_lo, thrust_hi, cg_max, wind_std, noise_std, tilt_max, omega_max;
    float prop_frac, cg_burn;
    float mu, mu_spin, mu_roll, restitution, rest_thr, erp, margin;
    unsigned seed_lo, seed_hi;
    long long env_base;
    int dbg_only;   // diagnostic builds (-DTVC_DBG) only
};

// Persistent per-env state, SoA planes of 16-byte groups (coalesced LDG.128/STG.128).
struct DevState {
    float4 *s0;  // px py pz ep_return
    float4 *s1;  // qx qy qz qw            (body->world, internal sign)
    float4 *s2;  // vx vy vz step(int)
    float4 *s3;  // wx wy wz flags(int): burn[0:11) phase[11:14) success[14] has_prev[15] consec[16:27)
    float4 *s4;  // a_prev0 a_prev1 hist_count(int) div(int): n_clip[0:10) n_run[10:20) | n_distinct
    float4 *d0;  // X: mass_scale thrust_scale cg_offset wind_x
    float4 *d1;  // X: wind_y episode(int) - -
    float *ring;      // [N][12] last ten clipped totals per env (48 B, three 16-byte loads), slot = push % 10; 2 floats padding
    unsigned *clipb;  // [32][N] fast diversity: value == -1000
    unsigned *runb;   // [32][N] fast diversity: value == predecessor
    float *hist;      // [1000][N] exact diversity only
    float2 *delay;    // [TVC_MAX_DELAY][N] X: actuator delay ring, slot = step % delay
    double *partial;  // [ceil(N/32)][16] episode statistics rows: one owner (CTA or 32-env group) per row per launch
    int *order;       // [N] env ids sorted per 1024-env chunk by class: in contact, may touch, airborne (classify_kernel)
    int *goff;        // [3][nchunks + 1] exclusive scans over the chunks of the per-chunk class counts (goff[c][0] = 0)
    unsigned *counter;  // [0] work-queue head of step_kernel_v2 (zeroed by classify_kernel), [1] classify_kernel's ticket
    uint8_t *cls;       // [N] class of every env for the next step's sort (written by step_kernel_v2; 0xFF = reset me)
    int nchunks;
    long long n;
};

struct DevIO {
    const float2 *actions;
    float *obs, *reward;
    uint8_t *term, *trunc;
    float *final_obs;
    float2 *actions_out;
    float *altitude, *tilt_deg, *omega_mag, *fuel, *position;
    int *phase, *step;
    uint8_t *success, *criteria_met;
    float *comp;
    unsigned long long t;  // lifetime step index (Philox action stream counter)
};

struct Env {
    float px, py, pz, ep_ret;
    float qx, qy, qz, qw;
    float vx, vy, vz;
    int step;
    float wx, wy, wz;
    int burn, phase, success, has_prev, consec;
    float ap0, ap1;
    int hist_count, n_clip, n_run;
    float mass_scale, thrust_scale, cg_off, wind_x, wind_y;
    int episode;
};

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based; Salmon et al. 2011).  Counter = (env_lo, env_hi16 | stream<<16,
// a, b), key = seed.  Same layout as the oracle so both sides draw identical bits.
// ------------------------------------------------------------------------------------------
enum { ST_DR_A = 1, ST_DR_B = 2, ST_NOISE_A = 3, ST_NOISE_B = 4, ST_ACTION = 5, ST_ACTOR = 6, ST_DR_C = 7 };

// Code size matters here: the step kernels run many desynchronised warps through ~90 KB of straight-line code,
// and ncu shows instruction-fetch stalls (no_instruction) dominating once the CTA barriers are gone.  So the
// generators are single out-of-line copies and the hot path uses short, slow-path-free math:
//   rcp_fast / sqrt_fast : MUFU.RCP / MUFU.RSQ based, <= 2 ulp, operands are well-scaled positive numbers
//   sincos_small         : degree-9/8 Taylor polynomials, |x| <= 0.8 rad, error < 4e-8
// (an earlier build, then bound by instruction fetch, measured raw `asm volatile` rcp/sqrt/rsqrt.approx.ftz 10 % slower;
//  with the class-ordered sequence the plain-asm single-MUFU rsqrt below is the faster form)
// reciprocal of a well-scaled positive number (masses, inertias, squared norms): one MUFU.RCP.  __fdividef(1, x) carries
// range handling for |x| > 2^126 (two predicated FMULs, two FSELs per call, 42 call sites: 5 % of the step kernel's
// instructions and 11 % of its stall samples in the ncu source view)
#ifdef TVC_RCP_FDIVIDEF
__device__ __forceinline__ float rcp_fast(float x) { return __fdividef(1.0f, x); }
#else
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif
// operand known to be a normal number (callers clamp it away from the subnormal range): one MUFU.RSQ, without the
// subnormal pre/post-scaling rsqrtf() carries
__device__ __forceinline__ float rsqrt_normal(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// squared magnitudes below 1e-30 (|.| < 1e-15) count as zero; callers of rsqrt_fast pass (near-)unit quaternion norms.
// (rsqrtf()'s subnormal pre/post-scaling measured 2.4 % slower end to end with identical trajectories; an L2 prefetch of
// the reward ring / bit-ring lines right after the state load measured 2.7 % slower.)
__device__ __forceinline__ float sqrt_fast(float x) { return x > 1e-30f ? x * rsqrt_normal(x) : 0.0f; }
__device__ __forceinline__ float rsqrt_fast(float x) { return rsqrt_normal(x); }
__device__ __forceinline__ void sincos_small(float x, float &s, float &c) {
    const float x2 = x * x;
    s = x * (1.0f + x2 * (-1.6666667e-1f + x2 * (8.3333333e-3f + x2 * (-1.9841270e-4f + x2 * 2.7557319e-6f))));
    c = 1.0f + x2 * (-0.5f + x2 * (4.1666667e-2f + x2 * (-1.3888889e-3f + x2 * (2.4801587e-5f - x2 * 2.7557319e-7f))));
}

static __device__ __noinline__ uint4 philox(unsigned seed_lo, unsigned seed_hi, long long gid, unsigned stream,
                                     unsigned a, unsigned b) {
    unsigned c0 = (unsigned)gid, c1 = (unsigned)(((unsigned long long)gid >> 32) & 0xFFFFu) | (stream << 16);
    unsigned c2 = a, c3 = b, k0 = seed_lo, k1 = seed_hi;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        unsigned h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        unsigned h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        unsigned n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// (k + 0.5) / 2^23 : exact in fp32
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 9) + 0.5f) * (1.0f / 8388608.0f); }
__device__ __forceinline__ void box_muller(unsigned x0, unsigned x1, float &n0, float &n1) {
    float r = sqrt_fast(-2.0f * logf(u01(x0)));
    float s, c;
    sincospif(2.0f * u01(x1), &s, &c);   // exact argument reduction, no Payne-Hanek slow path
    n0 = r * c; n1 = r * s;
}
// eight N(0,1) draws for one (env, episode, step) -- one out-of-line copy (sensor noise)
static __device__ __noinline__ void noise8(unsigned seed_lo, unsigned seed_hi, long long gid, unsigned episode, unsigned step,
                                    float n[8]) {
    const uint4 a = philox(seed_lo, seed_hi, gid, ST_NOISE_A, episode, step);
    const uint4 b = philox(seed_lo, seed_hi, gid, ST_NOISE_B, episode, step);
    box_muller(a.x, a.y, n[0], n[1]); box_muller(a.z, a.w, n[2], n[3]);
    box_muller(b.x, b.y, n[4], n[5]); box_muller(b.z, b.w, n[6], n[7]);
}

// ------------------------------------------------------------------------------------------
// Bullet helpers (rows B7, B8)
// ------------------------------------------------------------------------------------------
// btMatrix3x3::setRotation (ref:546 getMatrixFromQuaternion), row-major
__device__ __forceinline__ void quat_to_mat(float x, float y, float z, float w, float R[9]) {
    float d = x * x + y * y + z * z + w * w;
    float s = 2.0f * rcp_fast(d);
    float xs = x * s, ys = y * s, zs = z * s;
    float wx = w * xs, wy = w * ys, wz = w * zs;
    float xx = x * xs, xy = x * ys, xz = x * zs;
    float yy = y * ys, yz = y * zs, zz = z * zs;
    R[0] = 1.0f - (yy + zz); R[1] = xy - wz;          R[2] = xz + wy;
    R[3] = xy + wz;          R[4] = 1.0f - (xx + zz); R[5] = yz - wx;
    R[6] = xz - wy;          R[7] = yz + wx;          R[8] = 1.0f - (xx + yy);
}

// Row B7: the reported orientation is quat -> btMatrix3x3 -> quat, which normalises and fixes the
// sign (w > 0 when trace > 0, else the component of the largest diagonal is positive).
__device__ __forceinline__ void reported_quat(float x, float y, float z, float w, float &ox, float &oy, float &oz,
                                              float &ow) {
    float d = x * x + y * y + z * z + w * w;
    float s = 2.0f * rcp_fast(d);
    float xx = x * x * s, yy = y * y * s, zz = z * z * s;
    float m00 = 1.0f - (yy + zz), m11 = 1.0f - (xx + zz), m22 = 1.0f - (xx + yy);
    float trace = m00 + m11 + m22;
    float lead;
    if (trace > 0.0f) lead = w;
    else {
        int i = m00 < m11 ? (m11 < m22 ? 2 : 1) : (m00 < m22 ? 2 : 0);
        lead = i == 0 ? x : (i == 1 ? y : z);
    }
    float sc = rsqrt_fast(d);
    if (lead < 0.0f) sc = -sc;
    ox = x * sc; oy = y * sc; oz = z * sc; ow = w * sc;
}

// pybullet.c getEulerFromQuaternion (ref:614); only pitch and yaw feed the env (quirk Q7)
static __device__ __noinline__ void euler_gimbal_lock(float x, float y, float sarg, float &pitch, float &yaw) {
    if (sarg < 0.0f) { pitch = -1.5707963267948966f; yaw = 2.0f * atan2f(x, -y); }
    else { pitch = 1.5707963267948966f; yaw = 2.0f * atan2f(-x, y); }
}
__device__ __forceinline__ void euler_pitch_yaw(float x, float y, float z, float w, float &pitch, float &yaw) {
    float sarg = -2.0f * (x * z - w * y);
    if (sarg <= -0.99999f || sarg >= 0.99999f) euler_gimbal_lock(x, y, sarg, pitch, yaw);
    else {
        pitch = asinf(sarg);
        yaw = atan2f(2.0f * (x * y + w * z), w * w + x * x - y * y - z * z);
    }
}

// Row S4 (ref:530-533): fuel after n decrements.  (float)(1.0 - n*0.001) evaluated in fp64 equals
// the float32 rounding of the reference's iterated fp64 subtraction for every n in [0,1000]
// (checked exhaustively in tests/test_oracle.py); thresholds are integer compares on n.
__device__ __forceinline__ float fuel_of(int n) {
    return n >= 1000 ? 0.0f : __double2float_rn(__dsub_rn(1.0, __dmul_rn((double)n, 0.001)));
}

__device__ __forceinline__ float thrust_curve(int mode, int burn) {
    if (mode == 0) return 1.0f;
    float u = burn * 0.001f;
    if (u < 0.05f) return 1.0f + 6.0f * u;
    if (u < 0.15f) return 1.3f - 3.0f * (u - 0.05f);
    if (u < 0.9f) return 1.0f;
    return 1.0f - 5.0f * (u - 0.9f);
}

struct BodyP {
    float mass, inv_mass, Ixy, Iz, inv_Ixy, inv_Iz, cg;
};

__device__ __forceinline__ BodyP body_params(const DevCfg &c, bool X, float mass_scale, float cg_off, float fuel) {
    BodyP P;
    float burnt = 1.0f - fuel;
    float m = X ? c.mass * mass_scale * (1.0f - c.prop_frac * burnt) : c.mass;
    float cg = X ? cg_off + c.cg_burn * burnt : 0.0f;
    float len = 2.0f * c.half_len;
    P.mass = m; P.inv_mass = rcp_fast(m); P.cg = cg;
    P.Ixy = (1.0f / 12.0f) * m * (3.0f * c.radius * c.radius + len * len) + m * cg * cg;   // ref:431
    P.Iz = 0.5f * m * c.radius * c.radius;                                                  // ref:432
    P.inv_Ixy = rcp_fast(P.Ixy); P.inv_Iz = rcp_fast(P.Iz);
    return P;
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ------------------------------------------------------------------------------------------
// Ground contact: our documented model (DESIGN.md "Contact model"), continuous in the state.
//   p0 / p1 : lowest rim point of the bottom / top cap, direction -(R31,R32)/max(rho,1e-3)
//   f0..f2  : body-fixed rim points of the bottom cap at 0, 120, 240 degrees
// Entered when the lowest candidate is within `margin` and some row can bind (contact_needed); then
// all five rows are processed (speculative vn >= -gap/dt, Baumgarte for gap < 0, restitution),
// friction disc per point, spinning / rolling rows limited by the total normal impulse; projected
// Gauss-Seidel on velocities: contact_iters sweeps from a cold start in the first substep of a step,
// warm_iters sweeps warm-started with the previous substep's 18 impulses afterwards.  Bullet's own manifold/solver (row B9) is not
// reproducible without its source; this model is shared with the oracle by specification only.
// ------------------------------------------------------------------------------------------
// Entry rule of the contact model (same in the oracle), evaluated per substep by the env's own thread:
// the lowest candidate is within `margin` AND some row can bind at all -- a row binds only if
// (1+e) * approach speed * dt exceeds its gap, and the approach speed of any point is bounded by
// |vz| + |w| * reach.  When the rule fails the stored impulses are cleared.
__device__ __forceinline__ bool contact_needed_row(const DevCfg &c, const BodyP &P, float R31, float R32, float R33,
                                                   float pz, float vz, float wx, float wy, float wz) {
    const float r = c.radius, h = c.half_len;
    const float rho = sqrt_fast(R31 * R31 + R32 * R32);
    const float inv = rcp_fast(fmaxf(rho, 1e-3f));
    const float low = -r * (R31 * R31 + R32 * R32) * inv;
    const float gmin = pz + fminf(R33 * (-h - P.cg), R33 * (h - P.cg)) + low;
    if (!(gmin < c.margin)) return false;
    const float hh = h + fabsf(P.cg);
    const float reach = sqrt_fast(hh * hh + r * r);
    const float vmax = fabsf(vz) + sqrt_fast(wx * wx + wy * wy + wz * wz) * reach;
    return gmin - 1e-4f < (1.0f + c.restitution) * vmax * c.dt;
}
__device__ __forceinline__ bool contact_needed(const DevCfg &c, const BodyP &P, const float R[9], float pz, float vz,
                                               float wx, float wy, float wz) {
    return contact_needed_row(c, P, R[6], R[7], R[8], pz, vz, wx, wy, wz);
}

#ifdef TVC_PHASE_PROF2
// diagnostic build only (tools/phase_prof2.py): clock64 cycles per phase of step_kernel_v2, per class of group
// (0 = in contact / may touch, 1 = airborne), max over the lanes of a warp, summed over the groups:
// [0] pull (barrier + queue) [1] index + loads + env_pre [2] substeps outside the solver [3] solver entry (geometry,
// effective masses, warm start) [4] sweeps [5] env_post + stores [6] groups [7] substep barrier wait
__device__ unsigned long long g_ph2[2][8];
struct Ph2 { unsigned setup, sweeps, bar; };
#define PH2_ARG , Ph2 *ph2
#define PH2_PASS , ph2
#define PH2_CLK(x) const long long x = clock64()
#else
#define PH2_ARG
#define PH2_PASS
#define PH2_CLK(x)
#endif
// lam: this problem's 18 stored impulses, element j at lam[j * TVC_BLOCK] (shared memory column of the posting
// thread): normal(5), tangent-x(5), tangent-y(5), spin, roll-x, roll-y.  warm: apply them before sweeping.
template <int LS>   // LS: stride between the 18 impulses (TVC_BLOCK for the shared-memory column, 1 for a register array)
__device__ __forceinline__ void solve_contacts(const DevCfg &c, const BodyP &P, const float R[9], float pz, float &vx,
                                               float &vy, float &vz, float &wx, float &wy, float &wz, float *lam,
                                               bool warm, int iters PH2_ARG) {
    PH2_CLK(pc0);
    const float r = c.radius, h = c.half_len;
    const float R31 = R[6], R32 = R[7];
    const float rho = sqrt_fast(R31 * R31 + R32 * R32);
    const float inv = rcp_fast(fmaxf(rho, 1e-3f));
    const float ux = -R31 * inv, uy = -R32 * inv;
    const float zb = -h - P.cg, zt = h - P.cg;

    // world inverse inertia W = R diag(1/I) R^T (symmetric)
    const float ia = P.inv_Ixy, ib = P.inv_Iz;
    const float W00 = ia * (R[0] * R[0] + R[1] * R[1]) + ib * R[2] * R[2];
    const float W01 = ia * (R[0] * R[3] + R[1] * R[4]) + ib * R[2] * R[5];
    const float W02 = ia * (R[0] * R[6] + R[1] * R[7]) + ib * R[2] * R[8];
    const float W11 = ia * (R[3] * R[3] + R[4] * R[4]) + ib * R[5] * R[5];
    const float W12 = ia * (R[3] * R[6] + R[4] * R[7]) + ib * R[5] * R[8];
    const float W22 = ia * (R[6] * R[6] + R[7] * R[7]) + ib * R[8] * R[8];
    const float im = P.inv_mass;

    const float clx[5] = {r * ux, r * ux, r, -0.5f * r, -0.5f * r};
    const float cly[5] = {r * uy, r * uy, 0.0f, 0.8660254037844386f * r, -0.8660254037844386f * r};
    float ax[5], ay[5], az[5], tgt[5], imn[5], im1[5], im2[5];
    float ln[5], l1[5], l2[5];
    // A row whose normal constraint cannot bind (tgt <= vn with zero stored impulses) is an exact no-op together with its
    // friction rows and its warm-start term.  That is the normal case for the top cap (point 1) in every upright pose
    // and for most of the body-fixed rim points: measured with the oracle, an env in contact stands on its rim at a
    // median tilt of 21 degrees with 0.8 of the three fixed points within 4 mm of the plane.  Points 1-4 are therefore
    // visited lazily: each sweep tests `tgt > vn` (and "any stored impulse"), and only a thread that passes computes
    // the point's effective masses and runs the row.  Same results as visiting every row always.  (Sorting the
    // in-contact envs further by their set of near-ground rim points, so that whole warps skip the same rows, was
    // measured: 0.1074 ms against 0.1036 without -- ten classes scatter a group's 32 envs over more cache lines.)
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const float cz = i == 1 ? zt : zb;
        ax[i] = R[0] * clx[i] + R[1] * cly[i] + R[2] * cz;
        ay[i] = R[3] * clx[i] + R[4] * cly[i] + R[5] * cz;
        az[i] = R[6] * clx[i] + R[7] * cly[i] + R[8] * cz;
        const float gap = pz + az[i];
        if (i == 0) {
            // effective masses for n = z, t1 = x, t2 = y at arm (ax,ay,az)
            const float mn = im + (W00 * ay[i] * ay[i] - 2.0f * W01 * ax[i] * ay[i] + W11 * ax[i] * ax[i]);
            const float m1 = im + (W11 * az[i] * az[i] - 2.0f * W12 * az[i] * ay[i] + W22 * ay[i] * ay[i]);
            const float m2 = im + (W00 * az[i] * az[i] - 2.0f * W02 * az[i] * ax[i] + W22 * ax[i] * ax[i]);
            imn[i] = rcp_fast(mn); im1[i] = rcp_fast(m1); im2[i] = rcp_fast(m2);
        } else { imn[i] = 0.0f; im1[i] = 0.0f; im2[i] = 0.0f; }   // lazily, in the first sweep that needs them
        const float vn0 = vz + wx * ay[i] - wy * ax[i];
        const float rest = (vn0 < -c.rest_thr) ? -c.restitution * vn0 : 0.0f;
        tgt[i] = rest + (gap > 0.0f ? -gap * c.inv_dt : -c.erp * gap * c.inv_dt);
        ln[i] = warm ? lam[i * LS] : 0.0f;
        l1[i] = warm ? lam[(5 + i) * LS] : 0.0f;
        l2[i] = warm ? lam[(10 + i) * LS] : 0.0f;
    }
    float lsp = warm ? lam[15 * LS] : 0.0f, lr1 = warm ? lam[16 * LS] : 0.0f, lr2 = warm ? lam[17 * LS] : 0.0f;
    const float iW22 = rcp_fast(W22), iW00 = rcp_fast(W00), iW11 = rcp_fast(W11);
    if (warm) {   // apply the stored impulses at the current contact geometry
#pragma unroll
        for (int i = 0; i < 5; i++) {
            if (i != 0 && ln[i] == 0.0f && l1[i] == 0.0f && l2[i] == 0.0f) continue;   // adds exact zeros
            const float px_ = l1[i], py_ = l2[i], pn_ = ln[i];
            vx += px_ * im; vy += py_ * im; vz += pn_ * im;
            const float tx = ay[i] * pn_ - az[i] * py_, ty = az[i] * px_ - ax[i] * pn_, tz = ax[i] * py_ - ay[i] * px_;
            wx += W00 * tx + W01 * ty + W02 * tz;
            wy += W01 * tx + W11 * ty + W12 * tz;
            wz += W02 * tx + W12 * ty + W22 * tz;
        }
        wx += W02 * lsp + W00 * lr1 + W01 * lr2;
        wy += W12 * lsp + W01 * lr1 + W11 * lr2;
        wz += W22 * lsp + W02 * lr1 + W12 * lr2;
    }
    PH2_CLK(pc1);
    for (int it = 0; it < iters; it++) {
        float lsum = 0.0f;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            // normal row
            const float vn = vz + wx * ay[i] - wy * ax[i];
            if (i != 0) {
                if (!(tgt[i] > vn || ln[i] > 0.0f || l1[i] != 0.0f || l2[i] != 0.0f)) continue;   // exact no-op
                if (imn[i] == 0.0f) {
                    const float mn = im + (W00 * ay[i] * ay[i] - 2.0f * W01 * ax[i] * ay[i] + W11 * ax[i] * ax[i]);
                    const float m1 = im + (W11 * az[i] * az[i] - 2.0f * W12 * az[i] * ay[i] + W22 * ay[i] * ay[i]);
                    const float m2 = im + (W00 * az[i] * az[i] - 2.0f * W02 * az[i] * ax[i] + W22 * ax[i] * ax[i]);
                    imn[i] = rcp_fast(mn); im1[i] = rcp_fast(m1); im2[i] = rcp_fast(m2);
                }
            }
            const float nl = fmaxf(ln[i] + (tgt[i] - vn) * imn[i], 0.0f);
            const float d = nl - ln[i];
            ln[i] = nl;
            lsum += nl;
            vz += d * im;
            wx += (W00 * ay[i] - W01 * ax[i]) * d;
            wy += (W01 * ay[i] - W11 * ax[i]) * d;
            wz += (W02 * ay[i] - W12 * ax[i]) * d;
            // friction disc
            const float vt1 = vx + wy * az[i] - wz * ay[i];
            const float vt2 = vy + wz * ax[i] - wx * az[i];
            float a1 = l1[i] - vt1 * im1[i];
            float a2 = l2[i] - vt2 * im2[i];
            const float lim = c.mu * nl;
            const float mag2 = a1 * a1 + a2 * a2;
            const float sc = mag2 > lim * lim ? lim * rsqrt_normal(fmaxf(mag2, 1e-30f)) : 1.0f;   // branch-free disc projection
            a1 *= sc; a2 *= sc;
            const float d1 = a1 - l1[i], d2 = a2 - l2[i];
            l1[i] = a1; l2[i] = a2;
            vx += d1 * im; vy += d2 * im;
            const float tx = -az[i] * d2, ty = az[i] * d1, tz = ax[i] * d2 - ay[i] * d1;
            wx = fmaf(W00, tx, fmaf(W01, ty, fmaf(W02, tz, wx)));   // three FFMA per component, no separate add
            wy = fmaf(W01, tx, fmaf(W11, ty, fmaf(W12, tz, wy)));
            wz = fmaf(W02, tx, fmaf(W12, ty, fmaf(W22, tz, wz)));
        }
        {   // spinning / rolling friction rows, limited by the total normal impulse
            float lim = c.mu_spin * lsum;
            float nl = clampf(lsp - wz * iW22, -lim, lim);
            float d = nl - lsp; lsp = nl;
            wx += W02 * d; wy += W12 * d; wz += W22 * d;
            lim = c.mu_roll * lsum;
            nl = clampf(lr1 - wx * iW00, -lim, lim);
            d = nl - lr1; lr1 = nl;
            wx += W00 * d; wy += W01 * d; wz += W02 * d;
            nl = clampf(lr2 - wy * iW11, -lim, lim);
            d = nl - lr2; lr2 = nl;
            wx += W01 * d; wy += W11 * d; wz += W12 * d;
        }
    }
#pragma unroll
    for (int i = 0; i < 5; i++) { lam[i * LS] = ln[i]; lam[(5 + i) * LS] = l1[i]; lam[(10 + i) * LS] = l2[i]; }
    lam[15 * LS] = lsp; lam[16 * LS] = lr1; lam[17 * LS] = lr2;
#ifdef TVC_PHASE_PROF2
    { const long long pc2 = clock64(); ph2->setup += (unsigned)(pc1 - pc0); ph2->sweeps += (unsigned)(pc2 - pc1); }
#endif
}

// Shared-memory exchange used to compact ground-contact problems across the CTA: each env that needs
// the solver this substep posts its problem to a slot; the first ceil(M/32) (rotated) warps solve the
// M problems with (nearly) full lanes instead of every warp running the solver for a few lanes.
#define TVC_PROB_FIELDS 15
#ifdef TVC_PHASE_PROF
// diagnostic build only: cycles per phase of integrate(), summed over warp leaders
// [0] free-flight  [1] count barrier  [2] post + barrier  [3] solve (solver warps)  [4] wait for solver (others)
// [5] read-back + pose update  [6] warp-substeps  [7] solver warp-substeps
__device__ unsigned long long g_phase[8];
#define PH_T(x) const long long x = clock64()
#else
#define PH_T(x)
#endif

#define TVC_PROB_FIELDS 15
template <int B>   // B = threads (= envs) per CTA
struct ContactSmemT {
    float f[TVC_PROB_FIELDS][B];   // qx qy qz qw pz vx vy vz wx wy wz inv_mass inv_Ixy inv_Iz cg
    float lam[18][B];              // per ENV THREAD: impulses carried between the substeps of one step
    int owner[B];                  // per slot: posting thread | warm flag << 16
    int cnt[B / 32];
};
typedef ContactSmemT<TVC_BLOCK> ContactSmem;

// Rows B2, B4, B5, B6: K substeps with the world-frame force F and torque T held constant (Q3).
// Block-cooperative: EVERY thread of the CTA must call this (threads without an env pass live=false).
template <int B>
__device__ __forceinline__ void integrate(const DevCfg &c, const BodyP &P, Env &e, float Fx, float Fy, float Fz,
                                          float Tx, float Ty, float Tz, bool live, ContactSmemT<B> &sm) {
#ifdef TVC_PHASE_PROF2
    Ph2 ph2s = {0u, 0u, 0u}; Ph2 *ph2 = &ph2s;
#endif
    const float dt = c.dt;
    const float ax_ = Fx * P.inv_mass, ay_ = Fy * P.inv_mass, az_ = Fz * P.inv_mass;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool have_lam = false;   // cold start at every control step: stored impulses are logically zero
#ifdef TVC_PHASE_PROF
    long long ph0 = 0, ph1 = 0, ph2 = 0, ph3 = 0, ph4 = 0, ph5 = 0, phn = 0, phs = 0;
#endif
    for (int k = 0; k < c.K; k++) {
        PH_T(t0);
        float R[9];
        quat_to_mat(e.qx, e.qy, e.qz, e.qw, R);
        // B5: base-local angular acceleration, damping k(1 + |w|), no gyroscopic term
        float wl0 = R[0] * e.wx + R[3] * e.wy + R[6] * e.wz;
        float wl1 = R[1] * e.wx + R[4] * e.wy + R[7] * e.wz;
        float wl2 = R[2] * e.wx + R[5] * e.wy + R[8] * e.wz;
        float tl0 = R[0] * Tx + R[3] * Ty + R[6] * Tz;
        float tl1 = R[1] * Tx + R[4] * Ty + R[7] * Tz;
        float tl2 = R[2] * Tx + R[5] * Ty + R[8] * Tz;
        float wn2 = wl0 * wl0 + wl1 * wl1 + wl2 * wl2;
        float wn = wn2 > 2.220446049250313e-16f ? sqrt_fast(wn2) : 0.0f;
        float kd = c.ang_damp + c.ang_damp * wn;
        float wd0 = tl0 * P.inv_Ixy - wl0 * kd;
        float wd1 = tl1 * P.inv_Ixy - wl1 * kd;
        float wd2 = tl2 * P.inv_Iz - wl2 * kd;
        float dwx = R[0] * wd0 + R[1] * wd1 + R[2] * wd2;
        float dwy = R[3] * wd0 + R[4] * wd1 + R[5] * wd2;
        float dwz = R[6] * wd0 + R[7] * wd1 + R[8] * wd2;
        float vn2 = e.vx * e.vx + e.vy * e.vy + e.vz * e.vz;
        float vn = vn2 > 2.220446049250313e-16f ? sqrt_fast(vn2) : 0.0f;
        float kl = c.lin_damp + c.lin_damp * vn;
        e.wx = clampf(e.wx + dwx * dt, -100.0f, 100.0f);
        e.wy = clampf(e.wy + dwy * dt, -100.0f, 100.0f);
        e.wz = clampf(e.wz + dwz * dt, -100.0f, 100.0f);
        e.vx = clampf(e.vx + (ax_ - e.vx * kl) * dt, -100.0f, 100.0f);
        e.vy = clampf(e.vy + (ay_ - e.vy * kl) * dt, -100.0f, 100.0f);
        e.vz = clampf(e.vz + (az_ - e.vz * kl) * dt, -100.0f, 100.0f);

        if (c.ground) {
            // B9 (our model): contacts detected at the pre-integration pose, solved on velocities
            const bool need = live && contact_needed(c, P, R, e.pz, e.vz, e.wx, e.wy, e.wz);
            const unsigned bal = __ballot_sync(0xffffffffu, need);
            PH_T(t1);
            if (lane == 0) sm.cnt[warp] = __popc(bal);
            __syncthreads();
            PH_T(t2);
#ifdef TVC_PHASE_PROF
            long long t3 = t2, t4 = t2, t5 = t2; bool solver = false;
#endif
            int total = 0, base = 0;
#pragma unroll
            for (int w = 0; w < B / 32; w++) { const int n = sm.cnt[w]; if (w < warp) base += n; total += n; }
            if (total > 0) {                                   // CTA-uniform
                const int slot = base + __popc(bal & ((1u << lane) - 1u));
                if (need) {
                    sm.f[0][slot] = e.qx; sm.f[1][slot] = e.qy; sm.f[2][slot] = e.qz; sm.f[3][slot] = e.qw;
                    sm.f[4][slot] = e.pz;
                    sm.f[5][slot] = e.vx; sm.f[6][slot] = e.vy; sm.f[7][slot] = e.vz;
                    sm.f[8][slot] = e.wx; sm.f[9][slot] = e.wy; sm.f[10][slot] = e.wz;
                    sm.f[11][slot] = P.inv_mass; sm.f[12][slot] = P.inv_Ixy; sm.f[13][slot] = P.inv_Iz; sm.f[14][slot] = P.cg;
                    sm.owner[slot] = (int)threadIdx.x | (have_lam ? 0x10000 : 0);
                }
                __syncthreads();
#ifdef TVC_PHASE_PROF
                t3 = clock64();
#endif
                // rotate the solver warps over the SMSPs (warp w of every CTA sits on SMSP w % 4)
                // (co-resident CTAs differ by multiples of the SM count in blockIdx, so fold the high bits in)
                const unsigned bx = blockIdx.x;
                const int t = (threadIdx.x + 32 * ((bx + (bx >> 2) + (bx >> 4) + (bx >> 6) + k) & (B / 32 - 1))) & (B - 1);
                if (t < total) {
                    float Rs[9];
                    quat_to_mat(sm.f[0][t], sm.f[1][t], sm.f[2][t], sm.f[3][t], Rs);
                    BodyP Q;
                    Q.inv_mass = sm.f[11][t]; Q.inv_Ixy = sm.f[12][t]; Q.inv_Iz = sm.f[13][t]; Q.cg = sm.f[14][t];
                    float vx = sm.f[5][t], vy = sm.f[6][t], vz = sm.f[7][t];
                    float wx = sm.f[8][t], wy = sm.f[9][t], wz = sm.f[10][t];
                    const int ow = sm.owner[t];
                    solve_contacts<B>(c, Q, Rs, sm.f[4][t], vx, vy, vz, wx, wy, wz, &sm.lam[0][ow & 0xFFFF], (ow >> 16) != 0,
                                   k == 0 ? c.contact_iters : c.warm_iters PH2_PASS);
                    sm.f[5][t] = vx; sm.f[6][t] = vy; sm.f[7][t] = vz;
                    sm.f[8][t] = wx; sm.f[9][t] = wy; sm.f[10][t] = wz;
                }
#ifdef TVC_PHASE_PROF
                t4 = clock64(); solver = __any_sync(0xffffffffu, t < total);
#endif
                __syncthreads();
#ifdef TVC_PHASE_PROF
                t5 = clock64();
#endif
                if (need) {
                    e.vx = sm.f[5][slot]; e.vy = sm.f[6][slot]; e.vz = sm.f[7][slot];
                    e.wx = sm.f[8][slot]; e.wy = sm.f[9][slot]; e.wz = sm.f[10][slot];
                }
            }
            have_lam = need;   // entry rule failed -> stored impulses cleared
#ifdef TVC_PHASE_PROF
            {
                const long long t6 = clock64();
                ph0 += t1 - t0; ph1 += t2 - t1; ph2 += t3 - t2;
                if (solver) { ph3 += t4 - t3; phs += 1; ph4 += t5 - t4; } else ph4 += t5 - t3;
                ph5 += t6 - t5; phn += 1;
            }
#endif
        }

        // B6: semi-implicit Euler + exponential map, q <- dq (x) q, normalise
        e.px += dt * e.vx; e.py += dt * e.vy; e.pz += dt * e.vz;
        float ang = sqrt_fast(e.wx * e.wx + e.wy * e.wy + e.wz * e.wz);
        if (ang * dt > 0.7853981633974483f) ang = 0.7853981633974483f * c.inv_dt;
        // sin(x)/ang and cos(x) with x = ang*dt/2 <= pi/8: polynomials (Bullet switches to its own Taylor form below 1e-3)
        const float hx = 0.5f * ang * dt, hx2 = hx * hx;
        const float sc = 0.5f * dt * (1.0f + hx2 * (-1.6666667e-1f + hx2 * (8.3333333e-3f + hx2 * (-1.9841270e-4f + hx2 * 2.7557319e-6f))));
        const float cw = 1.0f + hx2 * (-0.5f + hx2 * (4.1666667e-2f + hx2 * (-1.3888889e-3f + hx2 * 2.4801587e-5f)));
        float bx = e.wx * sc, by = e.wy * sc, bz = e.wz * sc;
        float nx = cw * e.qx + bx * e.qw + by * e.qz - bz * e.qy;
        float ny = cw * e.qy + by * e.qw + bz * e.qx - bx * e.qz;
        float nz = cw * e.qz + bz * e.qw + bx * e.qy - by * e.qx;
        float nw = cw * e.qw - bx * e.qx - by * e.qy - bz * e.qz;
        float inv = rsqrt_fast(nx * nx + ny * ny + nz * nz + nw * nw);
        e.qx = nx * inv; e.qy = ny * inv; e.qz = nz * inv; e.qw = nw * inv;
    }
#ifdef TVC_PHASE_PROF
    if (lane == 0) {
        atomicAdd(&g_phase[0], (unsigned long long)ph0); atomicAdd(&g_phase[1], (unsigned long long)ph1);
        atomicAdd(&g_phase[2], (unsigned long long)ph2); atomicAdd(&g_phase[3], (unsigned long long)ph3);
        atomicAdd(&g_phase[4], (unsigned long long)ph4); atomicAdd(&g_phase[5], (unsigned long long)ph5);
        atomicAdd(&g_phase[6], (unsigned long long)phn); atomicAdd(&g_phase[7], (unsigned long long)phs);
    }
#endif
}

// Same K substeps for ONE env on its own thread: no shared memory, no CTA barriers, the contact solver inline with
// its 18 carried impulses in registers.  Used by step_kernel_v2, whose warps hold envs of one class (near the ground
// or not) so the `need` branch is nearly warp-uniform.
template <bool LOCKSTEP>   // LOCKSTEP: the CTA's warps re-align at every substep (every thread of the CTA must call this)
__device__ __forceinline__ void integrate_thread(const DevCfg &c, const BodyP &P, Env &e, float Fx, float Fy, float Fz,
                                                 float Tx, float Ty, float Tz PH2_ARG) {
    const float dt = c.dt;
    const float ax_ = Fx * P.inv_mass, ay_ = Fy * P.inv_mass, az_ = Fz * P.inv_mass;
    float lam[18];
#pragma unroll
    for (int j = 0; j < 18; j++) lam[j] = 0.0f;
    bool have_lam = false;   // cold start at every control step
    const float dI = P.inv_Iz - P.inv_Ixy;
    const float far_z = 1.001f * (c.half_len + fabsf(P.cg) + c.radius) + c.margin;
    for (int k = 0; k < c.K; k++) {
#ifdef TVC_PHASE_PROF2
        { const long long b0 = clock64(); if (LOCKSTEP) __syncthreads(); ph2->bar += (unsigned)(clock64() - b0); }
#else
        if (LOCKSTEP) __syncthreads();
#endif
        // B5 with I = diag(a, a, b): R diag(1/a,1/a,1/b) R^T tau = tau/a + (1/b - 1/a)(e.tau) e, e = body axis in the
        // world frame (third column of R); the k(1+|w|) damping is isotropic.  Only e and the third row of R are needed
        // outside the contact solver.
        const float s2 = 2.0f * rcp_fast(e.qx * e.qx + e.qy * e.qy + e.qz * e.qz + e.qw * e.qw);
        const float xs = e.qx * s2, ys = e.qy * s2, zs = e.qz * s2;
        const float e0 = e.qx * zs + e.qw * ys, e1 = e.qy * zs - e.qw * xs, e2 = 1.0f - (e.qx * xs + e.qy * ys);
        const float et = (e0 * Tx + e1 * Ty + e2 * Tz) * dI;
        float wn2 = e.wx * e.wx + e.wy * e.wy + e.wz * e.wz;
        float wn = wn2 > 2.220446049250313e-16f ? sqrt_fast(wn2) : 0.0f;
        float kd = c.ang_damp + c.ang_damp * wn;
        float dwx = Tx * P.inv_Ixy + et * e0 - e.wx * kd;
        float dwy = Ty * P.inv_Ixy + et * e1 - e.wy * kd;
        float dwz = Tz * P.inv_Ixy + et * e2 - e.wz * kd;
        float vn2 = e.vx * e.vx + e.vy * e.vy + e.vz * e.vz;
        float vn = vn2 > 2.220446049250313e-16f ? sqrt_fast(vn2) : 0.0f;
        float kl = c.lin_damp + c.lin_damp * vn;
        e.wx = clampf(e.wx + dwx * dt, -100.0f, 100.0f);
        e.wy = clampf(e.wy + dwy * dt, -100.0f, 100.0f);
        e.wz = clampf(e.wz + dwz * dt, -100.0f, 100.0f);
        e.vx = clampf(e.vx + (ax_ - e.vx * kl) * dt, -100.0f, 100.0f);
        e.vy = clampf(e.vy + (ay_ - e.vy * kl) * dt, -100.0f, 100.0f);
        e.vz = clampf(e.vz + (az_ - e.vz * kl) * dt, -100.0f, 100.0f);
        // the lowest point of the body is never lower than pz - 1.001 (h + |cg| + r): above `far_z` the entry rule fails
        // without evaluating it (same decision, one compare for the airborne envs)
        if (c.ground && e.pz < far_z) {
            const float R31 = e.qx * zs - e.qw * ys, R32 = e.qy * zs + e.qw * xs;   // third row of R (R33 == e2)
            if (contact_needed_row(c, P, R31, R32, e2, e.pz, e.vz, e.wx, e.wy, e.wz)) {
                float R[9];
                quat_to_mat(e.qx, e.qy, e.qz, e.qw, R);
                solve_contacts<1>(c, P, R, e.pz, e.vx, e.vy, e.vz, e.wx, e.wy, e.wz, lam, have_lam,
                                  k == 0 ? c.contact_iters : c.warm_iters PH2_PASS);
                have_lam = true;
            } else have_lam = false;
        } else have_lam = false;
        e.px += dt * e.vx; e.py += dt * e.vy; e.pz += dt * e.vz;
        float ang = sqrt_fast(e.wx * e.wx + e.wy * e.wy + e.wz * e.wz);
        if (ang * dt > 0.7853981633974483f) ang = 0.7853981633974483f * c.inv_dt;
        // sin(x)/ang and cos(x) with x = ang*dt/2 <= pi/8: polynomials (Bullet switches to its own Taylor form below 1e-3)
        const float hx = 0.5f * ang * dt, hx2 = hx * hx;
        const float sc = 0.5f * dt * (1.0f + hx2 * (-1.6666667e-1f + hx2 * (8.3333333e-3f + hx2 * (-1.9841270e-4f + hx2 * 2.7557319e-6f))));
        const float cw = 1.0f + hx2 * (-0.5f + hx2 * (4.1666667e-2f + hx2 * (-1.3888889e-3f + hx2 * 2.4801587e-5f)));
        float bx = e.wx * sc, by = e.wy * sc, bz = e.wz * sc;
        float nx = cw * e.qx + bx * e.qw + by * e.qz - bz * e.qy;
        float ny = cw * e.qy + by * e.qw + bz * e.qx - bx * e.qz;
        float nz = cw * e.qz + bz * e.qw + bx * e.qy - by * e.qx;
        float nw = cw * e.qw - bx * e.qx - by * e.qy - bz * e.qz;
        float inv = rsqrt_fast(nx * nx + ny * ny + nz * nz + nw * nw);
        e.qx = nx * inv; e.qy = ny * inv; e.qz = nz * inv; e.qw = nw * inv;
    }
}

// Contract X per-episode draws (one out-of-line copy): mass scale, thrust scale, wind x/y, cg offset, initial
// tilt quaternion (x, y, w) and angular velocity.
static __device__ __noinline__ void reset_draws(const DevCfg &c, long long gid, int episode, float d[11]) {
    uint4 a = philox(c.seed_lo, c.seed_hi, gid, ST_DR_A, (unsigned)episode, 0u);
    uint4 b = philox(c.seed_lo, c.seed_hi, gid, ST_DR_B, (unsigned)episode, 0u);
    uint4 w = philox(c.seed_lo, c.seed_hi, gid, ST_DR_C, (unsigned)episode, 0u);
    float n0, n1, n2, n3;
    box_muller(a.y, a.z, n0, n1);
    box_muller(b.x, b.y, n2, n3);
    d[0] = 1.0f + c.mass_var * (2.0f * u01(a.x) - 1.0f);
    d[1] = clampf(1.0f + c.thrust_std * n0, c.thrust_lo, c.thrust_hi);
    d[2] = c.wind_std * n1;
    d[3] = c.wind_std * n2;
    d[4] = c.cg_max * (2.0f * u01(a.w) - 1.0f);
    float tx = c.tilt_max * (2.0f * u01(b.z) - 1.0f);
    float ty = c.tilt_max * (2.0f * u01(b.w) - 1.0f);
    float ang = sqrtf(tx * tx + ty * ty);
    float sc = ang < 1e-6f ? 0.5f - ang * ang * (1.0f / 48.0f) : sinf(0.5f * ang) / ang;
    d[5] = tx * sc; d[6] = ty * sc; d[7] = cosf(0.5f * ang);
    d[8] = c.omega_max * (2.0f * u01(w.x) - 1.0f);
    d[9] = c.omega_max * (2.0f * u01(w.y) - 1.0f);
    d[10] = c.omega_max * (2.0f * u01(w.z) - 1.0f);
}

// ref:381-464 reset (rows S13, Q10, Q11, Q15) + Contract X per-episode draws
__device__ __forceinline__ void reset_env(const DevCfg &c, bool X, long long gid, Env &e, bool first_time) {
    e.episode += 1;
    e.px = 0.0f; e.py = 0.0f; e.pz = 1.0f; e.ep_ret = 0.0f;
    e.qx = 0.0f; e.qy = 0.0f; e.qz = 0.0f; e.qw = 1.0f;
    e.vx = e.vy = e.vz = 0.0f; e.step = 0;
    e.wx = e.wy = e.wz = 0.0f;
    e.burn = 0; e.phase = 0; e.success = 0;
    if (first_time || !(c.quirks & 2u)) e.consec = 0;
    if (first_time || !(c.quirks & 4u)) { e.has_prev = 0; e.ap0 = 0.0f; e.ap1 = 0.0f; e.hist_count = 0; e.n_clip = 0; e.n_run = 0; }
    e.mass_scale = 1.0f; e.thrust_scale = 1.0f; e.cg_off = 0.0f; e.wind_x = 0.0f; e.wind_y = 0.0f;
    if (X) {
        float d[11];
        reset_draws(c, gid, e.episode, d);
        e.mass_scale = d[0]; e.thrust_scale = d[1]; e.wind_x = d[2]; e.wind_y = d[3]; e.cg_off = d[4];
        e.qx = d[5]; e.qy = d[6]; e.qz = 0.0f; e.qw = d[7];
        e.wx = d[8]; e.wy = d[9]; e.wz = d[10];
    }
}

// ref:587-606 _get_enhanced_observation (+ Contract X sensor noise on obs[0:7])
__device__ __forceinline__ void build_obs(const DevCfg &c, bool X, long long gid, const Env &e, int phase_for_obs,
                                          float o[10]) {
    reported_quat(e.qx, e.qy, e.qz, e.qw, o[0], o[1], o[2], o[3]);
    o[4] = e.wx; o[5] = e.wy; o[6] = e.wz;
    o[7] = fuel_of(e.burn);
    {   // phase index / 7 (ref:593), correctly rounded constants; progress = min(1, step / max_steps) (ref:596)
        float ph = 0.14285715f * (float)phase_for_obs;          // == fl32(p / 7) except for p = 3 and p = 6
        ph = phase_for_obs == 3 ? 0.42857143f : ph;
        o[8] = phase_for_obs == 6 ? 0.85714287f : ph;
        o[9] = fminf(1.0f, (float)e.step * c.inv_max_steps);
    }
    if (X && c.noise_std > 0.0f) {
        float n[8];
        noise8(c.seed_lo, c.seed_hi, gid, (unsigned)e.episode, (unsigned)e.step, n);
#pragma unroll
        for (int i = 0; i < 7; i++) o[i] += c.noise_std * n[i];
    }
}

__device__ __forceinline__ void load_env(const DevState &st, bool X, long long i, Env &e) {
    float4 a = st.s0[i], b = st.s1[i], c = st.s2[i], d = st.s3[i], f = st.s4[i];
    e.px = a.x; e.py = a.y; e.pz = a.z; e.ep_ret = a.w;
    e.qx = b.x; e.qy = b.y; e.qz = b.z; e.qw = b.w;
    e.vx = c.x; e.vy = c.y; e.vz = c.z; e.step = __float_as_int(c.w);
    e.wx = d.x; e.wy = d.y; e.wz = d.z;
    int fl = __float_as_int(d.w);
    e.burn = fl & 0x7FF; e.phase = (fl >> 11) & 7; e.success = (fl >> 14) & 1; e.has_prev = (fl >> 15) & 1;
    e.consec = (fl >> 16) & 0x7FF;
    e.ap0 = f.x; e.ap1 = f.y; e.hist_count = __float_as_int(f.z);
    int dv = __float_as_int(f.w);
    e.n_clip = dv & 0x3FF; e.n_run = (dv >> 10) & 0x3FF;
    if (X) {
        float4 g = st.d0[i], h = st.d1[i];
        e.mass_scale = g.x; e.thrust_scale = g.y; e.cg_off = g.z; e.wind_x = g.w;
        e.wind_y = h.x; e.episode = __float_as_int(h.y);
    } else {
        e.mass_scale = 1.0f; e.thrust_scale = 1.0f; e.cg_off = 0.0f; e.wind_x = 0.0f; e.wind_y = 0.0f; e.episode = 0;
    }
}

__device__ __forceinline__ void store_env(const DevState &st, bool X, long long i, const Env &e) {
    st.s0[i] = make_float4(e.px, e.py, e.pz, e.ep_ret);
    st.s1[i] = make_float4(e.qx, e.qy, e.qz, e.qw);
    st.s2[i] = make_float4(e.vx, e.vy, e.vz, __int_as_float(e.step));
    int cs = e.consec > 0x7FF ? 0x7FF : e.consec;
    int fl = (e.burn & 0x7FF) | ((e.phase & 7) << 11) | ((e.success & 1) << 14) | ((e.has_prev & 1) << 15) | (cs << 16);
    st.s3[i] = make_float4(e.wx, e.wy, e.wz, __int_as_float(fl));
    int dv = (e.n_clip & 0x3FF) | ((e.n_run & 0x3FF) << 10);
    st.s4[i] = make_float4(e.ap0, e.ap1, __int_as_float(e.hist_count), __int_as_float(dv));
    if (X) {
        st.d0[i] = make_float4(e.mass_scale, e.thrust_scale, e.cg_off, e.wind_x);
        st.d1[i] = make_float4(e.wind_y, __int_as_float(e.episode), 0.0f, 0.0f);
    }
}

// Heuristic work class of an env for the coming step (classify_kernel sorts by it; the rollout kernel sorts its CTA by it):
// 0 = its lowest point touches the ground now, 1 = it may come within reach during the step, 2 = airborne.
#ifndef TVC_NOW_GAP
#define TVC_NOW_GAP 0.008f     // class 0: lower bound of the lowest point's height below this
#endif
#ifndef TVC_MAYBE_GAP
#define TVC_MAYBE_GAP 0.02f    // class 1: that bound minus the first-order travel over the step below this
#endif
__device__ __forceinline__ int class_of(const DevCfg &c, bool X, float pz, float qx, float qy, float qz, float qw, float vz,
                                        float wx, float wy, float wz, float cg_off) {
    if (!c.ground) return 2;
    const float cg = X ? fabsf(cg_off) + fabsf(c.cg_burn) : 0.0f;
    const float R31 = 2.0f * (qx * qz - qw * qy), R32 = 2.0f * (qy * qz + qw * qx);
    const float R33 = 1.0f - 2.0f * (qx * qx + qy * qy);
    const float hh = c.half_len + cg;
    const float gmin = pz - fabsf(R33) * hh - c.radius * sqrtf(R31 * R31 + R32 * R32);
    const float reach = sqrtf(hh * hh + c.radius * c.radius);
    const float travel = (fabsf(vz) + sqrtf(wx * wx + wy * wy + wz * wz) * reach) * (c.dt * (float)c.K);
    return gmin < TVC_NOW_GAP ? 0 : (gmin - travel < TVC_MAYBE_GAP ? 1 : 2);
}

struct StepResult {
    float obs[10];
    float reward;
    int terminated, truncated, reason;
    float alt, tilt, wmag, fuel;
    float comp[12];
    int viol;
};

struct Forces {
    float Fx, Fy, Fz, Tx, Ty, Tz;
    float a0, a1;   // clipped policy action
};

// First half of one env step (ref:466-476): S2 action processing, S3 control forces, S5 aerodynamics.
template <bool X>
__device__ __forceinline__ void env_pre(const DevCfg &c, const DevState &st, long long i, Env &e, float a0, float a1,
                                        BodyP &P, Forces &f) {
    // ---- S2 (ref:470-471) ----
    a0 = clampf(a0, -1.0f, 1.0f); a1 = clampf(a1, -1.0f, 1.0f);
    float p0 = a0, p1 = a1;
    if (X && c.delay > 0) {   // actuator delay: the command issued `delay` control steps ago
        int slot = e.step % c.delay;
        float2 old = st.delay[(long long)slot * st.n + i];
        st.delay[(long long)slot * st.n + i] = make_float2(a0, a1);
        if (e.step >= c.delay) { p0 = old.x; p1 = old.y; } else { p0 = 0.0f; p1 = 0.0f; }
    }
    const float pitch = p0 * c.gimbal_max, yaw = p1 * c.gimbal_max;

    // ---- S3 (ref:520-559) control forces from the pre-step state ----
    float fuel_pre = fuel_of(e.burn);
    P = body_params(c, X, e.mass_scale, e.cg_off, fuel_pre);
    float Fx = 0.0f, Fy = 0.0f, Fz = 0.0f, Tx = 0.0f, Ty = 0.0f, Tz = 0.0f;
    if (c.quirks & 1u) Fz += -9.81f * P.mass;                         // Q1: explicit gravity force
    if (e.burn < 1000) {                                               // Q4: fuel > 0 on entry
        int burn_before = e.burn;
        e.burn += 1;
        float T = c.thrust;
        if (X) T = T * e.thrust_scale * thrust_curve(c.thrust_curve, burn_before);
        float sp, cp, sy, cy;
        sincos_small(pitch, sp, cp); sincos_small(yaw, sy, cy);    // |angle| <= gimbal limit (< 0.8 rad)
        float fl0 = T * sy, fl1 = T * sp, fl2 = T * cp * cy;           // Q2 (ref:539-543)
        float R[9];
        quat_to_mat(e.qx, e.qy, e.qz, e.qw, R);
        float fwx = R[0] * fl0 + R[1] * fl1 + R[2] * fl2;
        float fwy = R[3] * fl0 + R[4] * fl1 + R[5] * fl2;
        float fwz = R[6] * fl0 + R[7] * fl1 + R[8] * fl2;
        float arm = -(c.half_len + P.cg);                              // thrust acts at the base (ref:550)
        float rx = R[2] * arm, ry = R[5] * arm, rz = R[8] * arm;
        Fx += fwx; Fy += fwy; Fz += fwz;
        Tx += ry * fwz - rz * fwy; Ty += rz * fwx - rx * fwz; Tz += rx * fwy - ry * fwx;
    }
    {   // ---- S5 (ref:561-585) aerodynamics ----
        float rho = 1.225f * expf(-e.pz / 8400.0f);
        float vmag = sqrt_fast(e.vx * e.vx + e.vy * e.vy + e.vz * e.vz);
        if (vmag > 0.1f) {                                             // Q5
            float dm = 0.5f * rho * (vmag * vmag) * 0.47f * (3.14159265358979f * 0.0025f);
            float k = -dm * rcp_fast(vmag);
            Fx += k * e.vx; Fy += k * e.vy; Fz += k * e.vz;
        }
        float ad = 0.02f * rho;
        Tx -= ad * e.wx; Ty -= ad * e.wy; Tz -= ad * e.wz;
    }
    if (X) { Fx += e.wind_x; Fy += e.wind_y; }
    Fz += -9.81f * P.mass;                                             // B4: world gravity (ref:338)
    f.Fx = Fx; f.Fy = Fy; f.Fz = Fz; f.Tx = Tx; f.Ty = Ty; f.Tz = Tz; f.a0 = a0; f.a1 = a1;
}

// Second half (ref:478-518), after the K substeps: S7-S12 and the reward R1-R10.
template <bool X, int DIV>
__device__ __forceinline__ void env_post(const DevCfg &c, const DevState &st, long long i, long long gid, Env &e,
                                         float a0, float a1, StepResult &r) {
    e.step += 1;
    // the ten-entry reward ring (R8) is read further down: ask for it now, so that its DRAM round trip runs under the
    // Euler angles, the reward terms and the Philox noise instead of stalling the warp where the values are needed
    float rv[10];
    {
        const float4 *rp = reinterpret_cast<const float4 *>(st.ring + 12 * i);
        const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
        rv[0] = r0.x; rv[1] = r0.y; rv[2] = r0.z; rv[3] = r0.w; rv[4] = r1.x; rv[5] = r1.y; rv[6] = r1.z; rv[7] = r1.w;
        rv[8] = r2.x; rv[9] = r2.y;
    }
    unsigned cw_early = 0u, rw_early = 0u;   // the bit-ring words of this push's slot (fast diversity mode), same reason
    if (DIV == 1) {
        const int wi0 = (e.hist_count % TVC_HIST) >> 5;
        cw_early = st.clipb[(long long)wi0 * st.n + i];
        rw_early = st.runb[(long long)wi0 * st.n + i];
    }

    // ---- S7 (ref:608-633) ----
    float ox, oy, oz, ow, epitch, eyaw;
    reported_quat(e.qx, e.qy, e.qz, e.qw, ox, oy, oz, ow);
    euler_pitch_yaw(ox, oy, oz, ow, epitch, eyaw);
    const float tilt = sqrtf(epitch * epitch + eyaw * eyaw);           // Q7
    const float wmag = sqrtf(e.wx * e.wx + e.wy * e.wy + e.wz * e.wz);
    const float vh = sqrtf(e.vx * e.vx + e.vy * e.vy), vv = fabsf(e.vz);
    const float alt = e.pz;
    const bool crashed = alt < 0.1f;                                   // Q17
    const int phase_pre = e.phase, success_pre = e.success;
    const int burn = e.burn;   // fuel thresholds as integer compares (row S4): <0.8 <=> n>=200, >0.1 <=> n<=899

    // ---- S9 (ref:635-657) ----
    if (e.phase == 0 && burn >= 200) e.phase = 1;
    else if (e.phase == 1 && alt < 5.0f) e.phase = 2;
    else if (e.phase == 2 && alt < 1.0f) e.phase = 3;
    else if (e.phase == 3 && alt < 0.5f) {
        if (tilt < 0.087f && wmag < 0.1f) { e.phase = 5; e.success = 1; }
    }
    // ---- S10 (ref:659-695): deque(100) all-true == 100 consecutive all-met pushes ----
    if (!e.success) {
        bool all_met = (tilt < 0.087f) && (vv < 2.0f && vh < 0.5f) && (0.2f <= alt && alt <= 2.0f) && (wmag < 0.1f);
        e.consec = all_met ? min(e.consec + 1, 0x7FF) : 0;
        if (e.consec >= 100) e.success = 1;
    }
    const bool lag = (c.quirks & 8u) != 0;
    const int phase_r = lag ? phase_pre : e.phase;
    const int success_r = lag ? success_pre : e.success;

    // ---- S8 (ref:587-606) ----
    build_obs(c, X, gid, e, phase_r, r.obs);
    const float fuel = r.obs[7] = fuel_of(e.burn);

    // ---- R1-R10 (ref:86-224) ----
    float comp[12];
#pragma unroll
    for (int k = 0; k < 12; k++) comp[k] = 0.0f;
    comp[0] = success_r ? 100.0f : (phase_r == 2 ? 10.0f : 0.0f);
    {
        float tp = expf(-10.0f * fmaxf(0.0f, tilt - 0.087f));
        float apn = expf(-5.0f * fmaxf(0.0f, wmag - 0.1f));
        float alp = (0.2f <= alt && alt <= 20.0f) ? 1.0f : 0.5f;
        comp[1] = ((tp + apn + alp) / 3.0f) * 50.0f;
    }
    const float ce = sqrtf(a0 * a0 + a1 * a1);
    if (burn <= 899 && ce < 0.5f) comp[2] = (fuel * (1.0f - ce)) * 20.0f;
    comp[3] = (tilt < 0.05f && wmag < 0.1f) ? 10.0f : ((tilt < 0.1f && wmag < 0.2f) ? 5.0f : 0.0f);
    if (e.has_prev) {                                                   // Q11
        float d0 = a0 - e.ap0, d1 = a1 - e.ap1;
        comp[4] = expf(-5.0f * sqrtf(d0 * d0 + d1 * d1)) * 5.0f;
    } else comp[4] = 5.0f;
    e.ap0 = a0; e.ap1 = a1; e.has_prev = 1;
    comp[5] = expf(-2.0f * fabsf(alt - 3.0f)) * 5.0f;
    if (crashed) comp[6] = -1000.0f;
    if (tilt > 0.52f) comp[7] = -500.0f * (tilt - 0.52f);
    if (ce > 0.9f) comp[8] = -50.0f * (ce - 0.9f);

    // R8 (ref:213-218): population variance of the last ten stored totals
    float adj = 0.0f;
    const int hc = e.hist_count;
    const int len = min(hc, TVC_HIST);
    float last_pushed = 0.0f;
    {
        int ls = (hc + 9) % 10;   // slot of push hc-1
#pragma unroll
        for (int k = 0; k < 10; k++) if (k == ls) last_pushed = rv[k];
        if (len > 10) {
            float s = ((rv[0] + rv[1]) + (rv[2] + rv[3])) + ((rv[4] + rv[5]) + (rv[6] + rv[7])) + rv[8] + rv[9];
            float mean = s / 10.0f, q = 0.0f;
#pragma unroll
            for (int k = 0; k < 10; k++) { float d = rv[k] - mean; q += d * d; }
            float var = q / 10.0f;
            if (var > 10000.0f) adj -= c.gp * var;
        }
    }
    // R9 (ref:220-224): diversity bonus, len(set(history)) > 0.8 len  <=>  5 distinct > 4 len
    int distinct = 0;
    if (DIV == 1) distinct = len - e.n_run - e.n_clip + (e.n_clip > 0 ? 1 : 0);
    if (DIV == 2) distinct = e.n_clip;   // exact mode keeps the distinct count in the n_clip field
    const bool div_flag = (DIV != 0) && (5 * distinct > 4 * len);
    if (div_flag) adj += c.db;
    float total = comp[0] + comp[1];
    total += comp[2]; total += comp[3]; total += comp[4]; total += comp[5];
    if (crashed) total += comp[6];
    if (tilt > 0.52f) total += comp[7];
    if (ce > 0.9f) total += comp[8];
    total += adj;
    comp[9] = adj; comp[10] = total; comp[11] = div_flag ? 1.0f : 0.0f;
    const float reward = clampf(total, -1000.0f, 200.0f);

    // ---- push into reward_history (ref:123) ----
    {
        const int slot = hc % TVC_HIST;
        if (DIV == 1) {
            const int wi = slot >> 5;
            const unsigned bit = 1u << (slot & 31);
            unsigned cw = cw_early, rw = rw_early;
            if (hc >= TVC_HIST) {
                if (cw & bit) e.n_clip--;
                if (rw & bit) e.n_run--;
                const int nxt = (slot + 1) % TVC_HIST, wj = nxt >> 5;
                const unsigned nb = 1u << (nxt & 31);
                if (wj == wi) { if (rw & nb) { rw &= ~nb; e.n_run--; } }
                else {
                    unsigned r2 = st.runb[(long long)wj * st.n + i];
                    if (r2 & nb) { st.runb[(long long)wj * st.n + i] = r2 & ~nb; e.n_run--; }
                }
            }
            const bool is_clip = reward == -1000.0f;
            const bool is_run = !is_clip && hc > 0 && last_pushed == reward;
            cw = is_clip ? (cw | bit) : (cw & ~bit);
            rw = is_run ? (rw | bit) : (rw & ~bit);
            e.n_clip += is_clip; e.n_run += is_run;
            st.clipb[(long long)wi * st.n + i] = cw;
            st.runb[(long long)wi * st.n + i] = rw;
        }
        if (DIV == 2) {
            // exact: one pass over the window counting copies of the leaving and the entering value
            const bool full = hc >= TVC_HIST;
            const float leaving = full ? st.hist[(long long)slot * st.n + i] : 0.0f;
            int cx = 0, cy = 0;
            for (int k = 0; k < len; k++) {
                float v = st.hist[(long long)k * st.n + i];
                cx += (v == reward); cy += (v == leaving);
            }
            if (full) {
                if (cy == 1) e.n_clip--;             // the leaving value had no other copy
                if (leaving == reward) cx -= 1;      // do not count the slot being overwritten
            }
            if (cx == 0) e.n_clip++;
            st.hist[(long long)slot * st.n + i] = reward;
        }
        st.ring[12 * i + (hc % 10)] = reward;
        e.hist_count = hc + 1;
        if (e.hist_count >= 2000000000) e.hist_count -= 1000000000;   // keeps % 10, % 1000 and len
    }
    e.ep_ret += reward;

    // ---- S11 (ref:697-721) ----
    int terminated = 0, truncated = 0, reason = 0;
    if (e.success) { terminated = 1; reason = 1; }                     // Q16
    else {
        if (crashed) { terminated = 1; reason = 2; }
        else if (tilt > 0.52f) { terminated = 1; reason = 3; }
        else if (alt > 20.0f) { terminated = 1; reason = 4; }
        else if (e.px * e.px + e.py * e.py > 2500.0f) { terminated = 1; reason = 5; }   // hypot(x,y) > 50
        if (e.step >= c.max_steps) truncated = 1;
    }
    r.reward = reward; r.terminated = terminated; r.truncated = truncated; r.reason = reason;
    r.alt = alt; r.tilt = tilt; r.wmag = wmag; r.fuel = fuel;
#pragma unroll
    for (int k = 0; k < 12; k++) r.comp[k] = comp[k];
    // scripts/train.py:620-641 _check_safety_violation
    r.viol = (tilt * 57.29577951308232f > 0.52f * 57.29577951308232f) || (wmag > 5.0f) || (alt < 0.1f) || (alt > 20.0f);
}

}  // namespace tvc

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

// quirk bits (include/tvc_b200.h TVC_Q_*)
enum : unsigned { Q_DOUBLE_GRAVITY = 1u << 0, Q_KEEP_CRITERIA = 1u << 1, Q_KEEP_REWARD_HIST = 1u << 2, Q_LAGGED_PHASE = 1u << 3,
                  Q_THRUST_VECTOR = 1u << 4, Q_FROZEN_FORCES = 1u << 5, Q_DRAG_CUTOFF = 1u << 6, Q_STACKED_DAMPING = 1u << 7,
                  Q_EULER_TILT = 1u << 8, Q_DIVERSITY_BONUS = 1u << 9, Q_VARIANCE_PENALTY = 1u << 10,
                  Q_SUCCESS_MASKS_TRUNCATION = 1u << 12, Q_CRASH_IS_COM_HEIGHT = 1u << 13 };

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
    int *order;       // [N] the class-ordered work sequence: env ids, all of class 0 (in contact), then 1 (may touch), then 2 (airborne)
    int *ccount;      // [2][3][nchunks] envs of each class per 1,024-env chunk in the NEXT sequence (double-buffered by the parity counter[5])
    int *scount;      // [2][3][nsuper]  the same per super-chunk (256 chunks)
    int *done_list;   // [N] envs whose episode ended in this step (re-initialised by close_kernel's reset CTAs)
    int *totals;      // [3] class totals of the current sequence (diagnostics)
    unsigned *counter;  // CTR_* words (below), one cache line each
    uint8_t *cls;       // [N] class of every env in the next sequence (written by whoever moved the env last)
    int nchunks, nsuper;
    long long n;
};

// words of DevState::counter, one 128-byte line each (the queue pulls of the step kernel do not share a line with anything)
enum { CTR_QUEUE = 0,      // work-queue head of step_kernel_v2
       CTR_SEQ = 32,       // sequence number S: written by close_kernel, read by the kernel that follows it
       CTR_SEQ2 = 64,      // the S the last step_kernel_v2 / classify_state_kernel ran with: read by close_kernel
       CTR_STEPS = 96,     // steps since the statistics were reset
       CTR_DONE = 128,     // [2] length of done_list, double-buffered by S & 1
       CTR_WORDS = 160 };

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

// Code size matters here: the step kernels run many desynchronised warps through straight-line code, and ncu showed
// instruction-fetch stalls (no_instruction) dominating at 90 KB.  So the per-episode draws and the gimbal-lock branch
// are single out-of-line copies and the hot path uses short, slow-path-free math:
//   rcp_fast / sqrt_fast : MUFU.RCP / MUFU.RSQ based, <= 2 ulp, operands are well-scaled positive numbers
//   sincos_small         : degree-9/8 Taylor polynomials, |x| <= 0.8 rad, error < 4e-8
// (an earlier build, then bound by instruction fetch, measured raw `asm volatile` rcp/sqrt/rsqrt.approx.ftz 10 % slower;
//  with the class-ordered sequence the plain-asm single-MUFU rsqrt below is the faster form)
// reciprocal of a well-scaled positive number (masses, inertias, squared norms): one MUFU.RCP.  __fdividef(1, x) carries
// range handling for |x| > 2^126 (two predicated FMULs, two FSELs per call, 42 call sites: 5 % of the step kernel's
// instructions and 11 % of its stall samples in the ncu source view)
#if defined(TVC_RCP_FDIVIDEF) || defined(TVC_HOST_TWIN)   // (host twin: tests/host_twin, g++ has no PTX)
__device__ __forceinline__ float rcp_fast(float x) { return __fdividef(1.0f, x); }
#elif defined(TVC_ACCURATE_MATH)   // diagnostic build: correctly rounded reciprocal / square root instead of the MUFU approximations
__device__ __forceinline__ float rcp_fast(float x) { return __frcp_rn(x); }
#else
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif
// operand known to be a normal number (callers clamp it away from the subnormal range): one MUFU.RSQ, without the
// subnormal pre/post-scaling rsqrtf() carries
#ifdef TVC_HOST_TWIN
__device__ __forceinline__ float rsqrt_normal(float x) { return 1.0f / sqrtf(x); }
#elif defined(TVC_ACCURATE_MATH)
__device__ __forceinline__ float rsqrt_normal(float x) { return __frcp_rn(__fsqrt_rn(x)); }
#else
__device__ __forceinline__ float rsqrt_normal(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif
// squared magnitudes below 1e-30 (|.| < 1e-15) count as zero; callers of rsqrt_fast pass (near-)unit quaternion norms.
// (rsqrtf()'s subnormal pre/post-scaling measured 2.4 % slower end to end with identical trajectories; an L2 prefetch of
// the reward ring / bit-ring lines right after the state load measured 2.7 % slower.)
__device__ __forceinline__ float sqrt_fast(float x) { return x > 1e-30f ? x * rsqrt_normal(x) : 0.0f; }
__device__ __forceinline__ float rsqrt_fast(float x) { return rsqrt_normal(x); }
// e^x for the reward terms and the air density (x <= 0, |x| < 100): ex2.approx of x log2(e), relative error < 3e-7 there --
// expf() adds a range reduction worth five more instructions and the same MUFU
__device__ __forceinline__ float exp_fast(float x) { return __expf(x); }
__device__ __forceinline__ void sincos_small(float x, float &s, float &c) {
    const float x2 = x * x;
    s = x * (1.0f + x2 * (-1.6666667e-1f + x2 * (8.3333333e-3f + x2 * (-1.9841270e-4f + x2 * 2.7557319e-6f))));
    c = 1.0f + x2 * (-0.5f + x2 * (4.1666667e-2f + x2 * (-1.3888889e-3f + x2 * (2.4801587e-5f - x2 * 2.7557319e-7f))));
}

// (inlined since the step kernel shrank to ~55 KB of code: as out-of-line copies the generator and the sensor-noise draw
//  were two serialised calls in the middle of env_post; inlined, their integer rounds interleave with the Euler-angle and
//  reward chains: 0.0821 -> 0.0789 ms per step.  Prefetching the next group's env ids under the statistics: no change.)
static __device__ __forceinline__ uint4 philox(unsigned seed_lo, unsigned seed_hi, long long gid, unsigned stream,
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
// Eight N(0,1) draws for one (env, episode, step) from ONE Philox block (sensor noise, Contract X; same bits in the oracle):
// each 32-bit word gives two 16-bit uniforms (k + 0.5) / 2^16 -- the radius from the low half, the angle from the high half --
// i.e. four Box-Muller pairs with |n| <= 4.9 sigma.  The noise is scaled by 0.02 before it meets a 1e-5 tolerance, so the
// logarithm and the sine / cosine are the single-MUFU forms (|error| < 1e-6 on a unit normal).
static __device__ __forceinline__ void noise8(unsigned seed_lo, unsigned seed_hi, long long gid, unsigned episode, unsigned step,
                                    float n[8]) {
    const uint4 a = philox(seed_lo, seed_hi, gid, ST_NOISE_A, episode, step);
    const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float u0 = ((float)(w[k] & 0xFFFFu) + 0.5f) * (1.0f / 65536.0f), u1 = ((float)(w[k] >> 16) + 0.5f) * (1.0f / 65536.0f);
        const float r = sqrt_fast(-2.0f * __logf(u0));
        float sn, cs;
        __sincosf(6.283185307179586f * u1, &sn, &cs);
        n[2 * k] = r * cs; n[2 * k + 1] = r * sn;
    }
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

// what env_pre hands to the K substeps
struct Forces {
    float Fx, Fy, Fz, Tx, Ty, Tz;   // world frame, thrust included (as evaluated from the pre-step state)
    float a0, a1;   // clipped policy action
    float fl0, fl1, fl2, arm;   // thrust in the body frame and its lever arm along body z (quirk Q3 cleared: thrust follows the body)
};

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
// Entered when the lowest candidate is within `margin` and some row can bind (contact_needed).  Normal targets:
// speculative vn >= -gap/dt, Baumgarte for gap < 0, restitution on the approach speed beyond the threshold.  The solve
// runs in the BODY frame (diagonal inverse inertia: 1/Iz = 400 only ever multiplies the body-z component of a row's
// angular Jacobian, small by geometry, not by cancellation) as a block Gauss-Seidel: a point that can bind gets its three
// rows solved together with its 3x3 Delassus matrix (stick / slide along the stick impulse with the normal row re-solved /
// release -- every branch meets its neighbour continuously); point 0 carries the axial spin row in its block (rim
// friction and axial spin are coupled through 1/Iz); torsional friction acts along the body's principal axes, where its
// three rows are decoupled from each other.  A solve that follows free flight runs contact_iters passes from a cold
// start, one that follows a solve runs warm_iters passes warm-started with the previous substep's 18 impulses.
// Bullet's own manifold/solver (row B9) is not reproducible without its source; this model is shared with the oracle
// (oracle/tvc_oracle.c solve_contacts) by specification only.
// ------------------------------------------------------------------------------------------
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
#ifdef TVC_SOLVE_STATS
// diagnostic build only (tools/solve_stats.py): which substeps of a step ran the contact solve, per lane
// g_ss[class by position][substep][0 = warps with a live lane, 1 = warps in which a lane solved, 2 = lanes that solved]
__device__ unsigned long long g_ss[3][16][3];
#define SS_ARG , unsigned &ss_mask
#define SS_PASS , ss_mask
#else
#define SS_ARG
#define SS_PASS
#endif

// The attitude in double, carried INSIDE a step by the envs that are within `margin` of the ground.  The normal targets divide a
// candidate's gap by dt (x500 at K = 10), and the gap sees the attitude through an arm of 0.5-0.6 m along the body axis: the 1e-7
// a float quaternion drifts by over the substeps of a step and the 6e-8 rounding of R33 - 1 are each 1e-5..1e-4 m/s in a target
// and, through rim friction and 1 / I, most of the tail an fp32 evaluation of this model shows against the fp64 oracle (the
// oracle's own gain on such steps is up to 2e4 rad/s per metre of height).  So near the ground the
// quaternion product ALSO runs in double, and the two O(0.5 m) terms of a cap centre's height, (pz + z_cap) + z_cap (R33 - 1),
// are combined in double before the result (a few cm) is rounded.  Everything else is computed from the float quaternion, and
// the state a step stores is float (the double attitude rounded once).
// Cost on B200 (tools/micro/fp64_rate.cu): DFMA at half the FFMA rate with a 10-cycle dependent latency; the float <-> double
// conversions are the expensive part (15 lanes/clk/SM), hence 8 of them per substep.  A/B of the step path: 0.0789 vs 0.0733 ms
// (DESIGN.md section 6 lists the cheaper-looking forms that measured worse: out of line in local memory, heights formed a
// substep ahead, tracking from far_z on).
struct HpAtt { double x, y, z, w; };
__device__ __forceinline__ void hp_start(HpAtt &a, float qx, float qy, float qz, float qw) {
    a.x = (double)qx; a.y = (double)qy; a.z = (double)qz; a.w = (double)qw;
}
// heights of the bottom / top cap centre above the plane: (pz + z_cap) + pzc + z_cap (R33 - 1), rounded once
__device__ __forceinline__ void hp_heights(const HpAtt &a, float pz, float pzc, double zbd, double ztd, float &Hb, float &Ht) {
    const double nz1 = -2.0 * (a.x * a.x + a.y * a.y);
    const double pzd = (double)pz, pcd = (double)pzc;
    Hb = (float)(((pzd + zbd) + pcd) + zbd * nz1); Ht = (float)(((pzd + ztd) + pcd) + ztd * nz1);
}
// q <- dq (x) q with the substep's float increment dq = (bx, by, bz, cw): its relative rounding scales a rotation of <= 0.01 rad,
// and a common factor leaves with the normalisation (first-order 1/sqrt, exact to 1e-13 for |n|^2 = 1 + O(1e-7))
__device__ __forceinline__ void hp_rotate(HpAtt &a, float cw, float bx, float by, float bz) {
    const double cb = (double)cw, b0 = (double)bx, b1 = (double)by, b2 = (double)bz;
    const double nx = cb * a.x + b0 * a.w + b1 * a.z - b2 * a.y;
    const double ny = cb * a.y + b1 * a.w + b2 * a.x - b0 * a.z;
    const double nz = cb * a.z + b2 * a.w + b0 * a.y - b1 * a.x;
    const double nw = cb * a.w - b0 * a.x - b1 * a.y - b2 * a.z;
    const double inv = 1.5 - 0.5 * (nx * nx + ny * ny + nz * nz + nw * nw);
    a.x = nx * inv; a.y = ny * inv; a.z = nz * inv; a.w = nw * inv;
}

// One point block: the point's Delassus matrix A (symmetric), the wanted change of the contact velocity (ex, ey, en) and
// the old impulses -> new impulses.  Stick solution p* = p + A^-1 e, accepted when p*_n > 0 and |p*_t| <= mu p*_n;
// otherwise the friction impulse keeps the direction of p*_t at magnitude mu p_n and the normal row is re-solved with
// that coupling; p_n <= 0 releases the point.
__device__ __forceinline__ void point_block(float Axx, float Axy, float Axn, float Ayy, float Ayn, float Ann, float ex, float ey,
                                            float en, float mu, float l1, float l2, float ln, float &px, float &py, float &pn) {
    const float c00 = Ayy * Ann - Ayn * Ayn, c01 = Axn * Ayn - Axy * Ann, c02 = Axy * Ayn - Axn * Ayy;
    const float c11 = Axx * Ann - Axn * Axn, c12 = Axy * Axn - Axx * Ayn, c22 = Axx * Ayy - Axy * Axy;
    const float idet = rcp_fast(Axx * c00 + Axy * c01 + Axn * c02);
    px = l1 + (c00 * ex + c01 * ey + c02 * en) * idet;
    py = l2 + (c01 * ex + c11 * ey + c12 * en) * idet;
    pn = ln + (c02 * ex + c12 * ey + c22 * en) * idet;
    const float mag2 = px * px + py * py, lim = mu * pn;
    if (pn > 0.0f && mag2 <= lim * lim) return;                       // stick
    if (mag2 > 0.0f) {                                                // slide (or a stick solution that pulls)
        const float is = rsqrt_normal(fmaxf(mag2, 1e-30f)), tx = px * is, ty = py * is;
        const float den = fmaxf(Ann + mu * (Axn * tx + Ayn * ty), 0.25f * Ann);
        const float rhs = en + Ann * ln + Axn * l1 + Ayn * l2;
        pn = fmaxf(rhs * rcp_fast(den), 0.0f);
        px = mu * pn * tx; py = mu * pn * ty;
    } else { px = 0.0f; py = 0.0f; pn = 0.0f; }                       // release
}

// What the contact solve carries between the substeps of one control step (cold at every step).  Point 0's impulses and the
// torsional impulses live in registers; the impulses of points 1-4 -- the top cap's lowest point and the three body-fixed
// rim points, which carry load in a few per cent of the solves -- live in a small per-thread array that only the threads
// that visit such a point ever touch.
struct ContactCarry {
    float l1, l2, ln;       // point 0: tangent x, tangent y, normal
    float lt0, lt1, lt2;    // torsional impulses about the body axes x, y (roll) and z (spin)
    unsigned xmask;         // bit i (1..4): point i holds a stored impulse in lamx[3 (i - 1) ..]
    bool have;              // the previous substep ran the solve: warm start
};

// body-frame arm of contact point i (1..4): i == 1 the top cap's lowest rim point, 2..4 the body-fixed rim points of the bottom cap
__device__ __forceinline__ void extra_point_arm(int i, float cx0, float cy0, float r, float zb, float zt, float &cx, float &cy, float &cz) {
    cx = i == 1 ? cx0 : (i == 2 ? r : -0.5f * r);
    cy = i == 1 ? cy0 : (i == 2 ? 0.0f : (i == 3 ? 0.8660254037844386f * r : -0.8660254037844386f * r));
    cz = i == 1 ? zt : zb;
}

// The contact solve of one substep (model: DESIGN.md section 4; oracle/tvc_oracle.c solve_contacts is the same algorithm in
// fp64).  Inputs beyond the state: the rotation matrix R (row-major, body -> world), Hb / Ht = the heights of the two cap
// centres above the plane, (pz + z_cap) + pzc + z_cap (R33 - 1), formed in double by integrate_thread (a candidate's gap is that
// plus its radial part, arm r = 0.05 m), u = the direction of the lowest rim point, gthr = the reach bound of the entry rule.
// All angular quantities are body-frame inside (w0, w1, w2).
__device__ __forceinline__ void solve_contacts(const DevCfg &c, const BodyP &P, const float R[9], float Hb, float Ht,
                                               float ux, float uy, float gthr, float &vx, float &vy, float &vz, float &wx, float &wy,
                                               float &wz, ContactCarry &cc, float *lamx PH2_ARG) {
    PH2_CLK(pc0);
    const float r = c.radius;
    const float zb = -c.half_len - P.cg, zt = c.half_len - P.cg;
    const float ia = P.inv_Ixy, ib = P.inv_Iz, im = P.inv_mass, mu = c.mu;
    const bool warm = cc.have;
    const int iters = warm ? c.warm_iters : c.contact_iters;
    // body-frame angular velocity wb0 = R^T w; (w0, w1, w2) carries the running value and w += R (wt - wb0) at the end, so that
    // a solve in which nothing binds leaves omega bit-identical
    const float wb0x = R[0] * wx + R[3] * wy + R[6] * wz, wb0y = R[1] * wx + R[4] * wy + R[7] * wz;
    const float wb0z = R[2] * wx + R[5] * wy + R[8] * wz;
    float w0 = wb0x, w1 = wb0y, w2 = wb0z;
    // ---- point 0: angular Jacobians c x (row of R) of the world-axis rows z (normal), x, y ----
    const float cx0 = r * ux, cy0 = r * uy;
    const float Jn0 = cy0 * R[8] - zb * R[7], Jn1 = zb * R[6] - cx0 * R[8], Jn2 = cx0 * R[7] - cy0 * R[6];
    const float Jx0 = cy0 * R[2] - zb * R[1], Jx1 = zb * R[0] - cx0 * R[2], Jx2 = cx0 * R[1] - cy0 * R[0];
    const float Jy0 = cy0 * R[5] - zb * R[4], Jy1 = zb * R[3] - cx0 * R[5], Jy2 = cx0 * R[4] - cy0 * R[3];
    float tgt0;
    bool act0;   // point 0 is in the substep's manifold: within reach or still holding an impulse (same rule as points 1-4)
    {
        const float gap = Hb + (R[6] * cx0 + R[7] * cy0);
        act0 = gap < gthr || cc.ln != 0.0f || cc.l1 != 0.0f || cc.l2 != 0.0f;
        const float vn0 = vz + (wb0x * Jn0 + wb0y * Jn1 + wb0z * Jn2);
        const float rest = (vn0 < -c.rest_thr) ? c.restitution * (-vn0 - c.rest_thr) : 0.0f;
        tgt0 = rest + (gap > 0.0f ? -gap * c.inv_dt : -c.erp * gap * c.inv_dt);
    }
    // ---- points 1-4: the substep's manifold holds those within reach (gap below the entry rule's own bound) or still
    // holding an impulse; their targets are fixed here, before the warm start (same rule and order in the oracle) ----
    unsigned amask = 0u;
    float tgtx[4];
#pragma unroll
    for (int i = 1; i < 5; i++) {
        float cx, cy, cz;
        extra_point_arm(i, cx0, cy0, r, zb, zt, cx, cy, cz);
        const float gap = (i == 1 ? Ht : Hb) + (R[6] * cx + R[7] * cy);
        tgtx[i - 1] = 0.0f;
        if (gap < gthr || (cc.xmask >> i) & 1u) {
            const float jn0 = cy * R[8] - cz * R[7], jn1 = cz * R[6] - cx * R[8], jn2 = cx * R[7] - cy * R[6];
            const float vn0 = vz + (wb0x * jn0 + wb0y * jn1 + wb0z * jn2);
            const float rest = (vn0 < -c.rest_thr) ? c.restitution * (-vn0 - c.rest_thr) : 0.0f;
            tgtx[i - 1] = rest + (gap > 0.0f ? -gap * c.inv_dt : -c.erp * gap * c.inv_dt);
            amask |= 1u << i;
        }
    }
    float l1 = cc.l1, l2 = cc.l2, ln = cc.ln, lt0 = cc.lt0, lt1 = cc.lt1, lt2 = cc.lt2;
    float lx_sum = 0.0f;    // normal impulses held by points 1-4
    if (warm) {   // apply the stored impulses at the current contact geometry
        vx += l1 * im; vy += l2 * im; vz += ln * im;
        w0 += ia * (Jx0 * l1 + Jy0 * l2 + Jn0 * ln + lt0);
        w1 += ia * (Jx1 * l1 + Jy1 * l2 + Jn1 * ln + lt1);
        w2 += ib * (Jx2 * l1 + Jy2 * l2 + Jn2 * ln + lt2);
        for (unsigned m = cc.xmask; m; m &= m - 1u) {
            const int i = __ffs(m) - 1;
            float cx, cy, cz;
            extra_point_arm(i, cx0, cy0, r, zb, zt, cx, cy, cz);
            const float p1 = lamx[3 * (i - 1)], p2 = lamx[3 * (i - 1) + 1], pn = lamx[3 * (i - 1) + 2];
            vx += p1 * im; vy += p2 * im; vz += pn * im;
            w0 += ia * ((cy * R[2] - cz * R[1]) * p1 + (cy * R[5] - cz * R[4]) * p2 + (cy * R[8] - cz * R[7]) * pn);
            w1 += ia * ((cz * R[0] - cx * R[2]) * p1 + (cz * R[3] - cx * R[5]) * p2 + (cz * R[6] - cx * R[8]) * pn);
            w2 += ib * ((cx * R[1] - cy * R[0]) * p1 + (cx * R[4] - cy * R[3]) * p2 + (cx * R[7] - cy * R[6]) * pn);
            lx_sum += pn;
        }
    }
    // point 0's Delassus matrix with the axial row folded in: body axis z is left out of the angular dynamics
    const float Axx = im + ia * (Jx0 * Jx0 + Jx1 * Jx1), Axy = ia * (Jx0 * Jy0 + Jx1 * Jy1), Axn = ia * (Jx0 * Jn0 + Jx1 * Jn1);
    const float Ayy = im + ia * (Jy0 * Jy0 + Jy1 * Jy1), Ayn = ia * (Jy0 * Jn0 + Jy1 * Jn1);
    const float Ann = im + ia * (Jn0 * Jn0 + Jn1 * Jn1);
    PH2_CLK(pc1);
    for (int it = 0; it < iters; it++) {
        float lsum = 0.0f;
        bool spin_done = false;
        {   // ---- point 0 (rim friction and axial spin are coupled through 1/Iz: the block is solved with the axial row
            //      sticking; if the axial impulse that needs exceeds its limit it goes to the limit and the point is solved again
            //      with the full angular dynamics -- both solutions coincide at the limit) ----
            const float un = vz + (w0 * Jn0 + w1 * Jn1);
            const float unz = un + w2 * Jn2;                             // true normal velocity
            if (act0 && (tgt0 > unz || ln > 0.0f || l1 != 0.0f || l2 != 0.0f)) {
                const float ux_ = vx + (w0 * Jx0 + w1 * Jx1), uy_ = vy + (w0 * Jy0 + w1 * Jy1);
                float px, py, pn;
                point_block(Axx, Axy, Axn, Ayy, Ayn, Ann, -ux_, -uy_, tgt0 - un, mu, l1, l2, ln, px, py, pn);
                float dx = px - l1, dy = py - l2, dn = pn - ln;
                const float slim = c.mu_spin * (pn + lx_sum);
                const float cand = lt2 - (w2 * P.Iz + (Jx2 * dx + Jy2 * dy + Jn2 * dn));
                if (cand >= -slim && cand <= slim) { w2 = 0.0f; lt2 = cand; }
                else {
                    const float slim0 = c.mu_spin * (ln + lx_sum);
                    const float nl = clampf(cand, -slim0, slim0);
                    w2 += ib * (nl - lt2);
                    lt2 = nl;
                    const float un3 = un + w2 * Jn2, ux3 = ux_ + w2 * Jx2, uy3 = uy_ + w2 * Jy2;
                    point_block(Axx + ib * Jx2 * Jx2, Axy + ib * Jx2 * Jy2, Axn + ib * Jx2 * Jn2, Ayy + ib * Jy2 * Jy2,
                                Ayn + ib * Jy2 * Jn2, Ann + ib * Jn2 * Jn2, -ux3, -uy3, tgt0 - un3, mu, l1, l2, ln, px, py, pn);
                    dx = px - l1; dy = py - l2; dn = pn - ln;
                    w2 += ib * (Jx2 * dx + Jy2 * dy + Jn2 * dn);
                }
                spin_done = true;
                vx += dx * im; vy += dy * im; vz += dn * im;
                w0 += ia * (Jx0 * dx + Jy0 * dy + Jn0 * dn);
                w1 += ia * (Jx1 * dx + Jy1 * dy + Jn1 * dn);
                l1 = px; l2 = py; ln = pn;
                lsum += pn;
            }
        }
        lx_sum = 0.0f;
        for (unsigned m = amask; m; m &= m - 1u) {   // ---- points 1-4 of this substep's manifold ----
            const int i = __ffs(m) - 1;
            float cx, cy, cz;
            extra_point_arm(i, cx0, cy0, r, zb, zt, cx, cy, cz);
            const float jn0 = cy * R[8] - cz * R[7], jn1 = cz * R[6] - cx * R[8], jn2 = cx * R[7] - cy * R[6];
            const float un = vz + (w0 * jn0 + w1 * jn1 + w2 * jn2);
            const bool held = (cc.xmask >> i) & 1u;
            float q1 = 0.0f, q2 = 0.0f, qn = 0.0f;
            if (held) { q1 = lamx[3 * (i - 1)]; q2 = lamx[3 * (i - 1) + 1]; qn = lamx[3 * (i - 1) + 2]; }
            const float tg = i == 1 ? tgtx[0] : (i == 2 ? tgtx[1] : (i == 3 ? tgtx[2] : tgtx[3]));
            if (!(tg > un || qn > 0.0f || q1 != 0.0f || q2 != 0.0f)) continue;   // exact no-op
            const float jx0 = cy * R[2] - cz * R[1], jx1 = cz * R[0] - cx * R[2], jx2 = cx * R[1] - cy * R[0];
            const float jy0 = cy * R[5] - cz * R[4], jy1 = cz * R[3] - cx * R[5], jy2 = cx * R[4] - cy * R[3];
            const float ux_ = vx + (w0 * jx0 + w1 * jx1 + w2 * jx2), uy_ = vy + (w0 * jy0 + w1 * jy1 + w2 * jy2);
            const float Bxx = im + (ia * (jx0 * jx0 + jx1 * jx1) + ib * jx2 * jx2);
            const float Bxy = ia * (jx0 * jy0 + jx1 * jy1) + ib * jx2 * jy2;
            const float Bxn = ia * (jx0 * jn0 + jx1 * jn1) + ib * jx2 * jn2;
            const float Byy = im + (ia * (jy0 * jy0 + jy1 * jy1) + ib * jy2 * jy2);
            const float Byn = ia * (jy0 * jn0 + jy1 * jn1) + ib * jy2 * jn2;
            const float Bnn = im + (ia * (jn0 * jn0 + jn1 * jn1) + ib * jn2 * jn2);
            float px, py, pn;
            point_block(Bxx, Bxy, Bxn, Byy, Byn, Bnn, -ux_, -uy_, tg - un, mu, q1, q2, qn, px, py, pn);
            const float dx = px - q1, dy = py - q2, dn = pn - qn;
            vx += dx * im; vy += dy * im; vz += dn * im;
            w0 += ia * (jx0 * dx + jy0 * dy + jn0 * dn);
            w1 += ia * (jx1 * dx + jy1 * dy + jn1 * dn);
            w2 += ib * (jx2 * dx + jy2 * dy + jn2 * dn);
            lamx[3 * (i - 1)] = px; lamx[3 * (i - 1) + 1] = py; lamx[3 * (i - 1) + 2] = pn;
            cc.xmask = (px != 0.0f || py != 0.0f || pn != 0.0f) ? (cc.xmask | (1u << i)) : (cc.xmask & ~(1u << i));
            lsum += pn;
            lx_sum += pn;
        }
        {   // torsional rows about the body axes x, y (and z when point 0 did not carry it): the impulse that zeroes the
            // component, limited by mu_k * (total normal impulse)
            const float lr = c.mu_roll * lsum;
            float nl = clampf(lt0 - w0 * P.Ixy, -lr, lr);
            w0 += ia * (nl - lt0); lt0 = nl;
            nl = clampf(lt1 - w1 * P.Ixy, -lr, lr);
            w1 += ia * (nl - lt1); lt1 = nl;
            if (!spin_done) {
                const float ls = c.mu_spin * lsum;
                nl = clampf(lt2 - w2 * P.Iz, -ls, ls);
                w2 += ib * (nl - lt2); lt2 = nl;
            }
        }
    }
    cc.l1 = l1; cc.l2 = l2; cc.ln = ln; cc.lt0 = lt0; cc.lt1 = lt1; cc.lt2 = lt2; cc.have = true;
    {   // back to the world frame: w += R (wt - wb0)
        const float d0 = w0 - wb0x, d1 = w1 - wb0y, d2 = w2 - wb0z;
        wx += R[0] * d0 + R[1] * d1 + R[2] * d2;
        wy += R[3] * d0 + R[4] * d1 + R[5] * d2;
        wz += R[6] * d0 + R[7] * d1 + R[8] * d2;
    }
#ifdef TVC_PHASE_PROF2
    { const long long pc2 = clock64(); ph2->setup += (unsigned)(pc1 - pc0); ph2->sweeps += (unsigned)(pc2 - pc1); }
#endif
}

// Rows B2, B4, B5, B6: K substeps for ONE env on its own thread with the world-frame force F and torque T held constant
// (Q3): no shared memory, no CTA barriers, the contact solver inline.  Used by step_kernel_v2 and the rollout kernel, whose
// warps hold envs of one class (near the ground or not), so the contact branch is nearly warp-uniform.
// LOCKSTEP: the CTA's warps re-align at every substep (every thread of the CTA must call this).
// FOLLOW (quirk Q3 cleared): the thrust force and torque are body-fixed and re-evaluated at every substep's attitude instead
// of being held constant in the world frame; a separate instantiation, so that the reference path carries none of it.
template <bool LOCKSTEP, bool FOLLOW>
__device__ __forceinline__ void integrate_thread(const DevCfg &c, const BodyP &P, Env &e, const Forces &f PH2_ARG SS_ARG) {
    const float dt = c.dt;
    float Fx = f.Fx, Fy = f.Fy, Fz = f.Fz, Tx = f.Tx, Ty = f.Ty, Tz = f.Tz;
    float ax_ = Fx * P.inv_mass, ay_ = Fy * P.inv_mass, az_ = Fz * P.inv_mass;
    float Fox = 0.0f, Foy = 0.0f, Foz = 0.0f, Tox = 0.0f, Toy = 0.0f, Toz = 0.0f;
    const float tl0 = -f.arm * f.fl1, tl1 = f.arm * f.fl0;   // (0, 0, arm) x thrust, body frame
    if (FOLLOW) {   // everything but the thrust: subtract the thrust as env_pre evaluated it (same attitude)
        float R[9];
        quat_to_mat(e.qx, e.qy, e.qz, e.qw, R);
        Fox = Fx - (R[0] * f.fl0 + R[1] * f.fl1 + R[2] * f.fl2); Foy = Fy - (R[3] * f.fl0 + R[4] * f.fl1 + R[5] * f.fl2);
        Foz = Fz - (R[6] * f.fl0 + R[7] * f.fl1 + R[8] * f.fl2);
        Tox = Tx - (R[0] * tl0 + R[1] * tl1); Toy = Ty - (R[3] * tl0 + R[4] * tl1); Toz = Tz - (R[6] * tl0 + R[7] * tl1);
    }
    ContactCarry cc;
    cc.l1 = cc.l2 = cc.ln = cc.lt0 = cc.lt1 = cc.lt2 = 0.0f; cc.xmask = 0u; cc.have = false;   // cold start at every control step
    float lamx[12];          // impulses of points 1-4, valid where cc.xmask says so
    float pzc = 0.0f;        // running compensation of the height update: true height = e.pz + pzc
    // Within `margin` of the ground the attitude is also carried in double INSIDE the step (HpAtt, above), valid while hp
    HpAtt att;
    att.x = att.y = att.z = att.w = 0.0;
    bool hp = false;
    const float dI = P.inv_Iz - P.inv_Ixy;
    const float hh = c.half_len + fabsf(P.cg);
    const float far_z = 1.001f * (hh + c.radius) + c.margin;
    const float reach = sqrt_fast(hh * hh + c.radius * c.radius);
    const float zb = -c.half_len - P.cg, zt = c.half_len - P.cg;
    float wmag = sqrt_fast(e.wx * e.wx + e.wy * e.wy + e.wz * e.wz);   // carried from substep to substep
    for (int k = 0; k < c.K; k++) {
#ifdef TVC_PHASE_PROF2
        { const long long b0 = clock64(); if (LOCKSTEP) __syncthreads(); ph2->bar += (unsigned)(clock64() - b0); }
#else
        if (LOCKSTEP) __syncthreads();
#endif
        if (FOLLOW && k > 0) {
            float R[9];
            quat_to_mat(e.qx, e.qy, e.qz, e.qw, R);
            Fx = Fox + (R[0] * f.fl0 + R[1] * f.fl1 + R[2] * f.fl2); Fy = Foy + (R[3] * f.fl0 + R[4] * f.fl1 + R[5] * f.fl2);
            Fz = Foz + (R[6] * f.fl0 + R[7] * f.fl1 + R[8] * f.fl2);
            Tx = Tox + (R[0] * tl0 + R[1] * tl1); Ty = Toy + (R[3] * tl0 + R[4] * tl1); Tz = Toz + (R[6] * tl0 + R[7] * tl1);
            ax_ = Fx * P.inv_mass; ay_ = Fy * P.inv_mass; az_ = Fz * P.inv_mass;
        }
        // B5 with I = diag(a, a, b): R diag(1/a,1/a,1/b) R^T tau = tau/a + (1/b - 1/a)(e.tau) e, e = body axis in the
        // world frame (third column of R); the k(1+|w|) damping is isotropic.  Only e and the third row of R are needed
        // outside the contact solver.  The quaternion is unit to rounding (normalised at the end of every substep, on
        // import and at reset: | |q|^2 - 1 | < 2e-7), so btMatrix3x3::setRotation's 2 / |q|^2 is 2 to the last bit or two.
        const float xs = e.qx + e.qx, ys = e.qy + e.qy, zs = e.qz + e.qz;
        const float nz1 = -(e.qx * xs + e.qy * ys);          // R33 - 1, without the cancellation
        const float e0 = e.qx * zs + e.qw * ys, e1 = e.qy * zs - e.qw * xs, e2 = 1.0f + nz1;
        const float et = (e0 * Tx + e1 * Ty + e2 * Tz) * dI;
        // |w|: the value the previous substep left (btVector3::safeNorm's cut-off at 1.5e-8 is below an ulp of k (1 + |w|))
        float kd = c.ang_damp + c.ang_damp * wmag;
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
        // B9 (our model): contacts detected at the pre-integration pose, solved on velocities.  Entry rule (same in the
        // oracle): the lowest candidate is within `margin` AND within reach, gap_min < g_reach = (1+e)(|vz| + |w| reach) dt
        // + 1e-4 -- a row binds only if (1+e) * approach speed * dt exceeds its gap.  The lowest point of the body is never
        // lower than pz - 1.001 (h + |cg| + r): above `far_z` the rule fails without evaluating it (one compare for the
        // airborne envs).  When the rule fails the stored impulses are cleared.
        bool solved = false;
        bool nearg = false;      // this substep sees the ground within `margin`: it works on the double attitude
        if (c.ground && e.pz < far_z) {
            const float R31 = e.qx * zs - e.qw * ys, R32 = e.qy * zs + e.qw * xs;   // third row of R (R33 == e2)
            const float rr = R31 * R31 + R32 * R32;
            const float rho = sqrt_fast(rr);
            const float inv = rcp_fast(fmaxf(rho, 1e-3f));
            const float low = -c.radius * rr * inv;
            const float hbf = (e.pz + zb) + pzc, htf = (e.pz + zt) + pzc;
            if (fminf(hbf + (zb * nz1 + low), htf + (zt * nz1 + low)) < c.margin) {
                nearg = true;
                if (!hp) { hp_start(att, e.qx, e.qy, e.qz, e.qw); hp = true; }
                float Hb, Ht;
                hp_heights(att, e.pz, pzc, (double)zb, (double)zt, Hb, Ht);
                const float gmin = fminf(Hb + low, Ht + low);
                const float vmax = fabsf(e.vz) + sqrt_fast(e.wx * e.wx + e.wy * e.wy + e.wz * e.wz) * reach;
                const float gthr = (1.0f + c.restitution) * vmax * dt + 1e-4f;
                if (gmin < gthr) {
                    float R[9];
                    R[0] = 1.0f - (e.qy * ys + e.qz * zs); R[1] = e.qx * ys - e.qw * zs; R[2] = e0;
                    R[3] = e.qx * ys + e.qw * zs; R[4] = 1.0f - (e.qx * xs + e.qz * zs); R[5] = e1;
                    R[6] = R31; R[7] = R32; R[8] = e2;
                    solve_contacts(c, P, R, Hb, Ht, -R31 * inv, -R32 * inv, gthr, e.vx, e.vy, e.vz, e.wx, e.wy, e.wz, cc, lamx PH2_PASS);
                    solved = true;
                }
            }
        }
        if (!solved && cc.have) { cc.l1 = cc.l2 = cc.ln = cc.lt0 = cc.lt1 = cc.lt2 = 0.0f; cc.xmask = 0u; cc.have = false; }
#ifdef TVC_SOLVE_STATS
        if (solved) ss_mask |= 1u << k;
#endif
        // B6: semi-implicit Euler; near the ground the height carries a running compensation (the contact targets divide the
        // gap by dt, so the 3e-8 rounding of pz + dt vz per substep would otherwise show up as 1.5e-5 m/s at dt = 0.002)
        e.px += dt * e.vx; e.py += dt * e.vy;
        {
            const float y = __fadd_rn(__fmul_rn(dt, e.vz), pzc), t = __fadd_rn(e.pz, y);
            pzc = __fsub_rn(y, __fsub_rn(t, e.pz));
            e.pz = t;
        }
        float ang = sqrt_fast(e.wx * e.wx + e.wy * e.wy + e.wz * e.wz);
        wmag = ang;
        if (ang * dt > 0.7853981633974483f) ang = 0.7853981633974483f * c.inv_dt;
        // sin(x)/ang and cos(x) with x = ang*dt/2 <= pi/8: polynomials (Bullet switches to its own Taylor form below 1e-3)
        const float hx = 0.5f * ang * dt, hx2 = hx * hx;
        // (hx <= pi/8: the x^8 terms are 1.5e-9 and 1.4e-8, below half an ulp)
        const float sc = 0.5f * dt * (1.0f + hx2 * (-1.6666667e-1f + hx2 * (8.3333333e-3f + hx2 * -1.9841270e-4f)));
        const float cw = 1.0f + hx2 * (-0.5f + hx2 * (4.1666667e-2f + hx2 * -1.3888889e-3f));
        float bx = e.wx * sc, by = e.wy * sc, bz = e.wz * sc;
        {
            float nx = cw * e.qx + bx * e.qw + by * e.qz - bz * e.qy;
            float ny = cw * e.qy + by * e.qw + bz * e.qx - bx * e.qz;
            float nz = cw * e.qz + bz * e.qw + bx * e.qy - by * e.qx;
            float nw = cw * e.qw - bx * e.qx - by * e.qy - bz * e.qz;
            // |dq (x) q|^2 = 1 + O(1e-7) (both factors are unit to rounding): 1/sqrt by its first-order expansion, exact to 1e-13
            float inv = 1.5f - 0.5f * (nx * nx + ny * ny + nz * nz + nw * nw);
            e.qx = nx * inv; e.qy = ny * inv; e.qz = nz * inv; e.qw = nw * inv;
        }
        // the same rotation on the double attitude, BESIDE the float one: the next substep's dynamics start from the float
        // quaternion (short chain); the double one is only read for the cap heights at the solver entry
        if (nearg) hp_rotate(att, cw, bx, by, bz);
        else hp = false;
    }
    // a tracked attitude is what the step hands on (rounded once); the float one beside it has drifted by a few 1e-8
    if (hp) { e.qx = (float)att.x; e.qy = (float)att.y; e.qz = (float)att.z; e.qw = (float)att.w; }
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
    if (first_time || !(c.quirks & Q_KEEP_CRITERIA)) e.consec = 0;
    if (first_time || !(c.quirks & Q_KEEP_REWARD_HIST)) { e.has_prev = 0; e.ap0 = 0.0f; e.ap1 = 0.0f; e.hist_count = 0; e.n_clip = 0; e.n_run = 0; }
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

// CG: L1-bypassing loads (ld.global.cg) for an env another SM stored during the same launch (reset workers of step_kernel_v2)
template <bool CG = false>
__device__ __forceinline__ void load_env(const DevState &st, bool X, long long i, Env &e) {
    float4 a, b, c, d, f;
    if (CG) { a = __ldcg(st.s0 + i); b = __ldcg(st.s1 + i); c = __ldcg(st.s2 + i); d = __ldcg(st.s3 + i); f = __ldcg(st.s4 + i); }
    else { a = st.s0[i]; b = st.s1[i]; c = st.s2[i]; d = st.s3[i]; f = st.s4[i]; }
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
        float4 g, h;
        if (CG) { g = __ldcg(st.d0 + i); h = __ldcg(st.d1 + i); } else { g = st.d0[i]; h = st.d1[i]; }
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
    const float gmin = pz - fabsf(R33) * hh - c.radius * sqrt_fast(R31 * R31 + R32 * R32);
    const float reach = sqrt_fast(hh * hh + c.radius * c.radius);
    const float travel = (fabsf(vz) + sqrt_fast(wx * wx + wy * wy + wz * wz) * reach) * (c.dt * (float)c.K);
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
    if (c.quirks & Q_DOUBLE_GRAVITY) Fz += -9.81f * P.mass;           // Q1: explicit gravity force
    f.fl0 = 0.0f; f.fl1 = 0.0f; f.fl2 = 0.0f; f.arm = 0.0f;
    if (e.burn < 1000) {                                               // Q4: fuel > 0 on entry
        int burn_before = e.burn;
        e.burn += 1;
        float T = c.thrust;
        if (X) T = T * e.thrust_scale * thrust_curve(c.thrust_curve, burn_before);
        float sp, cp, sy, cy;
        sincos_small(pitch, sp, cp); sincos_small(yaw, sy, cy);    // |angle| <= gimbal limit (< 0.8 rad)
        float fl0 = T * sy, fl1 = T * sp, fl2 = T * cp * cy;           // Q2 (ref:539-543)
        if (!(c.quirks & Q_THRUST_VECTOR)) {                           // Q2 cleared: the same direction, |F| = T
            const float sc = T * rsqrt_normal(fmaxf(fl0 * fl0 + fl1 * fl1 + fl2 * fl2, 1e-30f));
            fl0 *= sc; fl1 *= sc; fl2 *= sc;
        }
        float R[9];
        quat_to_mat(e.qx, e.qy, e.qz, e.qw, R);
        float fwx = R[0] * fl0 + R[1] * fl1 + R[2] * fl2;
        float fwy = R[3] * fl0 + R[4] * fl1 + R[5] * fl2;
        float fwz = R[6] * fl0 + R[7] * fl1 + R[8] * fl2;
        float arm = -(c.half_len + P.cg);                              // thrust acts at the base (ref:550)
        float rx = R[2] * arm, ry = R[5] * arm, rz = R[8] * arm;
        Fx += fwx; Fy += fwy; Fz += fwz;
        Tx += ry * fwz - rz * fwy; Ty += rz * fwx - rx * fwz; Tz += rx * fwy - ry * fwx;
        f.fl0 = fl0; f.fl1 = fl1; f.fl2 = fl2; f.arm = arm;
    }
    {   // ---- S5 (ref:561-585) aerodynamics ----
        float rho = 1.225f * exp_fast(e.pz * (-1.0f / 8400.0f));
        float vmag = sqrt_fast(e.vx * e.vx + e.vy * e.vy + e.vz * e.vz);
        if (vmag > 0.1f || (!(c.quirks & Q_DRAG_CUTOFF) && vmag > 0.0f)) {   // Q5
            float dm = 0.5f * rho * (vmag * vmag) * 0.47f * (3.14159265358979f * 0.0025f);
            float k = -dm * rcp_fast(vmag);
            Fx += k * e.vx; Fy += k * e.vy; Fz += k * e.vz;
        }
        if (c.quirks & Q_STACKED_DAMPING) {                            // Q6
            float ad = 0.02f * rho;
            Tx -= ad * e.wx; Ty -= ad * e.wy; Tz -= ad * e.wz;
        }
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
    float tilt = sqrt_fast(epitch * epitch + eyaw * eyaw);             // Q7
    if (!(c.quirks & Q_EULER_TILT)) {                                  // Q7 cleared: angle between the body axis and the vertical
        const float sxy = 2.0f * sqrtf((ox * ox + oy * oy) * (oz * oz + ow * ow));   // |sin|, reported quaternion is unit norm
        tilt = atan2f(sxy, 1.0f - 2.0f * (ox * ox + oy * oy));
    }
    const float wmag = sqrt_fast(e.wx * e.wx + e.wy * e.wy + e.wz * e.wz);
    const float vh = sqrt_fast(e.vx * e.vx + e.vy * e.vy), vv = fabsf(e.vz);
    const float alt = e.pz;
    bool crashed = alt < 0.1f;                                         // Q17
    if (!(c.quirks & Q_CRASH_IS_COM_HEIGHT)) {                         // Q17 cleared: a hard or tilted touchdown
        const float cgx = X ? e.cg_off + c.cg_burn * (1.0f - fuel_of(e.burn)) : 0.0f;
        const float R31 = 2.0f * (ox * oz - ow * oy), R32 = 2.0f * (oy * oz + ow * ox), R33 = 1.0f - 2.0f * (ox * ox + oy * oy);
        const float low = alt + fminf(R33 * (-c.half_len - cgx), R33 * (c.half_len - cgx)) - c.radius * sqrtf(R31 * R31 + R32 * R32);
        crashed = low < 0.01f && (e.vz < -2.0f || tilt > 0.52f);
    }
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
    const bool lag = (c.quirks & Q_LAGGED_PHASE) != 0;
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
        float tp = exp_fast(-10.0f * fmaxf(0.0f, tilt - 0.087f));
        float apn = exp_fast(-5.0f * fmaxf(0.0f, wmag - 0.1f));
        float alp = (0.2f <= alt && alt <= 20.0f) ? 1.0f : 0.5f;
        comp[1] = ((tp + apn + alp) * (1.0f / 3.0f)) * 50.0f;
    }
    const float ce = sqrt_fast(a0 * a0 + a1 * a1);
    if (burn <= 899 && ce < 0.5f) comp[2] = (fuel * (1.0f - ce)) * 20.0f;
    comp[3] = (tilt < 0.05f && wmag < 0.1f) ? 10.0f : ((tilt < 0.1f && wmag < 0.2f) ? 5.0f : 0.0f);
    if (e.has_prev) {                                                   // Q11
        float d0 = a0 - e.ap0, d1 = a1 - e.ap1;
        comp[4] = exp_fast(-5.0f * sqrt_fast(d0 * d0 + d1 * d1)) * 5.0f;
    } else comp[4] = 5.0f;
    e.ap0 = a0; e.ap1 = a1; e.has_prev = 1;
    comp[5] = exp_fast(-2.0f * fabsf(alt - 3.0f)) * 5.0f;
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
            float mean = s * 0.1f, q = 0.0f;
#pragma unroll
            for (int k = 0; k < 10; k++) { float d = rv[k] - mean; q += d * d; }
            float var = q * 0.1f;
            if (var > 10000.0f && (c.quirks & Q_VARIANCE_PENALTY)) adj -= c.gp * var;
        }
    }
    // R9 (ref:220-224): diversity bonus, len(set(history)) > 0.8 len  <=>  5 distinct > 4 len
    int distinct = 0;
    if (DIV == 1) distinct = len - e.n_run - e.n_clip + (e.n_clip > 0 ? 1 : 0);
    if (DIV == 2) distinct = e.n_clip;   // exact mode keeps the distinct count in the n_clip field
    const bool div_flag = (DIV != 0) && (c.quirks & Q_DIVERSITY_BONUS) && (5 * distinct > 4 * len);
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
    if (e.success) {                                                   // Q16
        terminated = 1; reason = 1;
        if (!(c.quirks & Q_SUCCESS_MASKS_TRUNCATION) && e.step >= c.max_steps) truncated = 1;
    } else {
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

/*
 * tvc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See tvc_oracle.h.
 *
 * Every function cites the reference lines (relative to /root/reference/) or the SURVEY.md
 * section-8(a) row it restates.  "ref:" = env/enhanced_rocket_tvc_env.py.
 *
 * Arithmetic: IEEE double, evaluated in the reference's operation order so that the env
 * layer reproduces the reference's own Python code bit-for-bit wherever that code works in
 * float64.  Where NumPy 2 keeps float32 intermediates (np.linalg.norm of the float32
 * action, ref:152,173,203) the same float32 roundings are emulated.
 * Build with -ffp-contract=off (no FMA contraction).
 */
#include "tvc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI_D 3.14159265358979323846

/* ------------------------------------------------------------------ */
/* small vector helpers                                                */
/* ------------------------------------------------------------------ */
static inline void cross3(const double a[3], const double b[3], double o[3]) {
    double x = a[1] * b[2] - a[2] * b[1];
    double y = a[2] * b[0] - a[0] * b[2];
    double z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline double dot3(const double a[3], const double b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
/* row-major 3x3 times vector */
static inline void matvec(const double m[9], const double v[3], double o[3]) {
    double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
    double y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
    double z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline void matTvec(const double m[9], const double v[3], double o[3]) {
    double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
    double y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
    double z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ------------------------------------------------------------------ */
/* Bullet helper functions (row B8, B7)                                */
/* ------------------------------------------------------------------ */

/* btMatrix3x3::setRotation, as used by pybullet getMatrixFromQuaternion (ref:546). Row B8. */
void orc_matrix_from_quat(const double q[4], double m[9]) {
    double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    double s = 2.0 / d;
    double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
    double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
    double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
    double yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
    m[0] = 1.0 - (yy + zz); m[1] = xy - wz;         m[2] = xz + wy;
    m[3] = xy + wz;         m[4] = 1.0 - (xx + zz); m[5] = yz - wx;
    m[6] = xz - wy;         m[7] = yz + wx;         m[8] = 1.0 - (xx + yy);
}

/* btMatrix3x3::getRotation (matrix -> quaternion). */
static void quat_from_matrix(const double m[9], double q[4]) {
    double trace = m[0] + m[4] + m[8];
    double t[4];
    if (trace > 0.0) {
        double s = sqrt(trace + 1.0);
        t[3] = s * 0.5;
        s = 0.5 / s;
        t[0] = (m[7] - m[5]) * s;
        t[1] = (m[2] - m[6]) * s;
        t[2] = (m[3] - m[1]) * s;
    } else {
        int i = m[0] < m[4] ? (m[4] < m[8] ? 2 : 1) : (m[0] < m[8] ? 2 : 0);
        int j = (i + 1) % 3, k = (i + 2) % 3;
        double s = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
        t[i] = s * 0.5;
        s = 0.5 / s;
        t[3] = (m[k * 3 + j] - m[j * 3 + k]) * s;
        t[j] = (m[j * 3 + i] + m[i * 3 + j]) * s;
        t[k] = (m[k * 3 + i] + m[i * 3 + k]) * s;
    }
    q[0] = t[0]; q[1] = t[1]; q[2] = t[2]; q[3] = t[3];
}

/* Row B7: getBasePositionAndOrientation reports btTransform(setRotation(q)).getRotation():
 * quaternion -> matrix -> quaternion, which canonicalises the sign. */
void orc_reported_quat(const double q[4], double out[4]) {
    double m[9];
    orc_matrix_from_quat(q, m);
    quat_from_matrix(m, out);
}

/* pybullet.c getEulerFromQuaternion (ref:614, ref:728). Row B8. */
void orc_euler_from_quat(const double q[4], double rpy[3]) {
    double sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], squ = q[3] * q[3];
    double sarg = -2 * (q[0] * q[2] - q[3] * q[1]);
    if (sarg <= -0.99999) {
        rpy[0] = 0; rpy[1] = -0.5 * PI_D; rpy[2] = 2 * atan2(q[0], -q[1]);
    } else if (sarg >= 0.99999) {
        rpy[0] = 0; rpy[1] = 0.5 * PI_D; rpy[2] = 2 * atan2(-q[0], q[1]);
    } else {
        rpy[0] = atan2(2 * (q[1] * q[2] + q[3] * q[0]), squ - sqx - sqy + sqz);
        rpy[1] = asin(sarg);
        rpy[2] = atan2(2 * (q[0] * q[1] + q[3] * q[2]), squ + sqx - sqy - sqz);
    }
}

/* ------------------------------------------------------------------ */
/* Physics layer                                                       */
/* ------------------------------------------------------------------ */

void orc_body_params_default(orc_body_params *p) {
    /* ref:412-432 mass/inertia; ref:451-458 damping + contact material; ref:338-345 world */
    double mass = 2.0, length = 1.0, radius = 0.05;
    memset(p, 0, sizeof(*p));
    p->mass = mass;
    p->inertia[0] = p->inertia[1] = (1.0 / 12.0) * mass * (3 * radius * radius + length * length);
    p->inertia[2] = (1.0 / 2.0) * mass * radius * radius;
    p->lin_damp = 0.01; p->ang_damp = 0.02;
    p->use_gyro = 0;
    p->substeps = 4;
    p->gravity[0] = 0; p->gravity[1] = 0; p->gravity[2] = -9.81;
    p->dt_step = 0.02;
    p->max_vel = 100.0;
    p->ground = 1;
    p->contact_iters = 2;
    p->warm_iters = 1;
    p->radius = radius; p->half_len = 0.5 * length; p->cg = 0.0;
    p->mu = 0.3 * 0.8;                       /* ref:456 x ref:350, Bullet combines by product */
    p->mu_spin = 0.1 * 0.8 + 0.1 * 0.3;      /* ref:351,457: Bullet combined torsional friction */
    p->mu_roll = 0.05 * 0.8 + 0.05 * 0.3;    /* ref:352,458 */
    p->restitution = 0.1 * 0.0 + 0.1;        /* ref:455; plane restitution 0 -> we keep the body's */
    p->rest_threshold = 0.2;                 /* Bullet restitutionVelocityThreshold default */
    p->erp = 0.2;                            /* Bullet m_erp2 default */
    p->margin = 0.05;
}

void orc_body_init(orc_body *b, const double pos[3], const double quat[4]) {
    memset(b, 0, sizeof(*b));
    memcpy(b->pos, pos, sizeof(double) * 3);
    memcpy(b->quat, quat, sizeof(double) * 4);
}

/* Row B3: applyExternalForce(..., pos, WORLD_FRAME) on the base: addBaseForce(F);
 * addBaseTorque((pos - base_origin) x F), evaluated once at call time. */
void orc_apply_external_force(orc_body *b, const double f[3], const double pos_world[3]) {
    double rel[3] = {pos_world[0] - b->pos[0], pos_world[1] - b->pos[1], pos_world[2] - b->pos[2]};
    double t[3];
    cross3(rel, f, t);
    for (int i = 0; i < 3; i++) { b->force[i] += f[i]; b->torque[i] += t[i]; }
}
void orc_apply_external_torque(orc_body *b, const double t[3]) {
    for (int i = 0; i < 3; i++) b->torque[i] += t[i];
}

/*
 * `real` is double for the oracle proper.  Compiling with -DORC_PHYS_FLOAT makes the physics
 * layer (contact + integrator) evaluate in float32, which emulates the device arithmetic on the
 * CPU; tests use that second build only to *measure* fp32 sensitivity, never as the reference.
 */
#ifdef ORC_PHYS_FLOAT
typedef float real;
#define R_SQRT sqrtf
#define R_SIN sinf
#define R_COS cosf
#define R_EPS2 2.220446049250313e-16f
#else
typedef double real;
#define R_SQRT sqrt
#define R_SIN sin
#define R_COS cos
#define R_EPS2 2.220446049250313e-16
#endif

/* -DORC_GEO_DOUBLE (with -DORC_PHYS_FLOAT only): the float build with the CUDA kernel's precision split (tvc_device.cuh
 * HpAtt) -- the attitude and the two O(0.5 m) terms of a contact candidate's height stay double, everything else is float.
 * tests/test_oracle.py uses it to show on the CPU which part of an fp32 evaluation makes the contact tail. */
#if defined(ORC_GEO_DOUBLE) && defined(ORC_PHYS_FLOAT)
typedef double greal;
#define ORC_SPLIT 1
#else
typedef real greal;
#define ORC_SPLIT 0
#endif
static inline real r_clamp(real x, real lo, real hi) { return x < lo ? lo : (x > hi ? hi : x); }

static greal r_matrix_from_quat(const greal qg[4], real m[9]) {
    const real q[4] = {(real)qg[0], (real)qg[1], (real)qg[2], (real)qg[3]};   /* (the matrix itself is `real` in every build) */
    real d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    real s = (real)2.0 / d;
    real xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
    real wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
    real xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
    real yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
    m[0] = (real)1.0 - (yy + zz); m[1] = xy - wz;                 m[2] = xz + wy;
    m[3] = xy + wz;               m[4] = (real)1.0 - (xx + zz);   m[5] = yz - wx;
    m[6] = xz - wy;               m[7] = yz + wx;                 m[8] = (real)1.0 - (xx + yy);
#if ORC_SPLIT
    return -2.0 * (qg[0] * qg[0] + qg[1] * qg[1]) / (qg[0] * qg[0] + qg[1] * qg[1] + qg[2] * qg[2] + qg[3] * qg[3]);
#else
    return -(xx + yy);   /* R33 - 1 without the cancellation */
#endif
}

/*
 * Ground contact -- OUR documented model (option (a)); Bullet's
 * GJK manifold + btMultiBodyConstraintSolver cannot be restated without its source (row B9).
 * Same model, constants and row order in the CUDA kernel (tvc_device.cuh solve_contacts).
 * Designed to be CONTINUOUS in the state so that fp32 and fp64 evaluations stay close:
 *
 *  - plane z = 0, normal +z; cylinder caps at local z = -+half_len - cg
 *  - stateless 5-point manifold per substep, always in this order:
 *      p0  lowest rim point of the bottom cap,  p1  lowest rim point of the top cap
 *          (direction -(R31,R32)/max(rho, 1e-3): slides to the cap centre as the axis becomes vertical)
 *      f0,f1,f2  three body-fixed rim points of the bottom cap at 0, 120, 240 degrees
 *  - entered only when the lowest candidate is closer than `margin` AND some row can bind at all:
 *    gap_min < g_reach = (1+e) (|vz| + |w| reach) dt + 1e-4  (otherwise the stored impulses are cleared);
 *    the substep's manifold holds the candidates with gap < g_reach and those that still carry an impulse
 *  - normal target per point: vn >= -gap/dt (gap >= 0, speculative) or vn >= -erp*gap/dt (gap < 0,
 *    Baumgarte), plus restitution e on the approach speed beyond the threshold (continuous at the threshold).
 *    The gap is formed as (pz + cz) + cz (nbz - 1) + cx nbx + cy nby so that the two O(0.5) terms cancel first
 *    and the height update carries a running compensation: the targets divide the gap by dt (x500 at K = 10).
 *  - BODY-FRAME rows: the angular velocity is carried in the body frame during the solve, so the inverse
 *    inertia stays diag(1/Ixy, 1/Ixy, 1/Iz) and the large 1/Iz (400 kg^-1 m^-2) only ever multiplies the
 *    body-z component of a row's angular Jacobian c x d_b (|c_x|, |c_y| <= r), which is small by geometry
 *    instead of by cancellation.  A solve in which nothing binds leaves omega bit-identical.
 *  - torsional friction (Bullet's spinning / rolling friction, ref:351-352, :457-458) acts along the body's
 *    principal axes: axial spin limited by mu_spin * (total normal impulse), the two transverse axes by
 *    mu_roll * (total normal impulse).  Along principal axes the three rows are decoupled from each other
 *    (a world-axis formulation couples them through 1/Iz, and a scalar Gauss-Seidel between them needs
 *    6-8 passes at an impact).
 *  - block Gauss-Seidel: a point that can bind (target above its normal velocity, or a stored impulse) gets
 *    its three rows (normal z, tangents x and y, world axes) solved TOGETHER with the point's 3x3 Delassus
 *    matrix A = (1/m) I + J I^-1 J^T: stick solution p* = p + A^-1 e, accepted when p*_n > 0 and
 *    |p*_t| <= mu p*_n; otherwise slide: the friction impulse keeps the direction of p*_t at magnitude
 *    mu p_n and the normal row is re-solved with that coupling (p_n = rhs / (A_nn + mu A_nt . t)); p_n <= 0
 *    releases the point.  Every branch meets its neighbour continuously.
 *  - point 0 (the manifold's first point, the one Bullet attaches its torsional rows to) carries the axial spin
 *    row in its block: rim friction and axial spin are coupled through 1/Iz (a scalar pass between them
 *    contracts by ~0.67 only).  The block is first solved with the axial row sticking (axis z removed from the
 *    angular dynamics, contact velocity evaluated at w_z = 0); if the axial impulse that needs stays inside
 *    its limit the solution is exact, otherwise the axial impulse goes to the limit and the point is solved
 *    again without the fold (both solutions coincide at the limit).
 *  - after the five points: the transverse torsional rows (and the axial one if point 0 was not visited).
 *  - a solve that follows free flight starts cold and runs contact_iters passes; a substep that follows a solve
 *    is warm-started with its 18 impulses (applied before the passes) and runs warm_iters passes.  Impulses do
 *    not persist across steps.
 */
/* diagnostic counters (not thread-safe; meaningful for serial runs only) */
long long orc_dbg_substeps = 0, orc_dbg_entered = 0, orc_dbg_canbind = 0;
long long orc_dbg_blocks[4] = {0, 0, 0, 0};   /* release, stick, slide, points 1-4 visited */

/* Diagnostic for the parity reports (orc_step_out.contact_margin): the manifold has DISCRETE decisions -- the entry rule
 * (gmin < margin, gmin < gthr) and each point's reach rule (gap < gthr).  An evaluation in another arithmetic that lands on
 * the other side of one of them solves a different manifold in that substep (an impact a substep earlier or later).  This is
 * the distance in metres between the deciding quantity and its threshold, smallest over the decisions of the current step;
 * 1e30 when the step made none.  Thread-local: orc_step runs envs on OpenMP threads. */
static _Thread_local double orc_tl_margin = 1e30;
static inline void note_margin(double d) { if (d < 0) d = -d; if (d < orc_tl_margin) orc_tl_margin = d; }

typedef struct { real Jx[3], Jy[3], Jn[3]; } orc_rows;

/* angular Jacobians (body frame) of the three world-axis rows at body-frame arm c: c x d_b */
static inline void point_rows(const real c[3], const real xb[3], const real yb[3], const real nb[3], orc_rows *J) {
    J->Jx[0] = c[1] * xb[2] - c[2] * xb[1]; J->Jx[1] = c[2] * xb[0] - c[0] * xb[2]; J->Jx[2] = c[0] * xb[1] - c[1] * xb[0];
    J->Jy[0] = c[1] * yb[2] - c[2] * yb[1]; J->Jy[1] = c[2] * yb[0] - c[0] * yb[2]; J->Jy[2] = c[0] * yb[1] - c[1] * yb[0];
    J->Jn[0] = c[1] * nb[2] - c[2] * nb[1]; J->Jn[1] = c[2] * nb[0] - c[0] * nb[2]; J->Jn[2] = c[0] * nb[1] - c[1] * nb[0];
}

typedef struct { real Axx, Axy, Axn, Ayy, Ayn, Ann; } orc_sym3;

/* one point block: stick / slide / release for the Delassus matrix A and the wanted velocity change e at old impulses
 * (l1, l2, ln); returns the new impulses */
static inline void point_block(const orc_sym3 *A, real ex, real ey, real en, real mu, real l1, real l2, real ln,
                               real *opx, real *opy, real *opn) {
    const real Axx = A->Axx, Axy = A->Axy, Axn = A->Axn, Ayy = A->Ayy, Ayn = A->Ayn, Ann = A->Ann;
    const real c00 = Ayy * Ann - Ayn * Ayn, c01 = Axn * Ayn - Axy * Ann, c02 = Axy * Ayn - Axn * Ayy;
    const real c11 = Axx * Ann - Axn * Axn, c12 = Axy * Axn - Axx * Ayn, c22 = Axx * Ayy - Axy * Axy;
    const real idet = (real)1.0 / (Axx * c00 + Axy * c01 + Axn * c02);
    real px = l1 + (c00 * ex + c01 * ey + c02 * en) * idet;
    real py = l2 + (c01 * ex + c11 * ey + c12 * en) * idet;
    real pn = ln + (c02 * ex + c12 * ey + c22 * en) * idet;
    const real mag2 = px * px + py * py, lim = mu * pn;
    if (pn > 0 && mag2 <= lim * lim) orc_dbg_blocks[1]++;                 /* stick */
    else if (mag2 > 0) {
        /* sliding (or a stick solution that pulls): the friction impulse keeps the direction of the stick
         * impulse at magnitude mu p_n, and the normal row is re-solved with that coupling; p_n <= 0 releases */
        const real is = (real)1.0 / R_SQRT(mag2), tx = px * is, ty = py * is;
        real den = Ann + mu * (Axn * tx + Ayn * ty);
        if (den < (real)0.25 * Ann) den = (real)0.25 * Ann;
        const real rhs = en + Ann * ln + Axn * l1 + Ayn * l2;
        pn = rhs / den;
        if (pn < 0) pn = 0;
        px = mu * pn * tx; py = mu * pn * ty;
        orc_dbg_blocks[pn > 0 ? 2 : 0]++;
    } else { px = 0; py = 0; pn = 0; orc_dbg_blocks[0]++; }              /* release */
    *opx = px; *opy = py; *opn = pn;
}

/* lam[18]: impulses carried between the substeps of a step: normal(5), tangent-x(5), tangent-y(5), torsional about the
 * body axes z (spin), x, y (roll) */
static void solve_contacts(const orc_body_params *p, real dt, const real R[9], greal nz1, greal pz, greal pzc, real v[3], real w[3],
                           real lam[18], int *have_lam) {
    const real r = (real)p->radius, h = (real)p->half_len, cg = (real)p->cg, margin = (real)p->margin;
    orc_dbg_substeps++;
    const real R31 = R[6], R32 = R[7];
    const real rho = R_SQRT(R31 * R31 + R32 * R32);
    const real inv = (real)1.0 / (rho > (real)1e-3 ? rho : (real)1e-3);
    const real ux = -R31 * inv, uy = -R32 * inv;
    const real zb = -h - cg, zt = h - cg;
    const real low = r * (R31 * ux + R32 * uy);          /* = -r*rho outside the regularised zone */
#if ORC_SPLIT   /* the kernel's association: cap-centre height in double, rounded, plus the float radial part */
    const real Hb = (real)(((pz + (greal)zb) + pzc) + (greal)zb * nz1), Ht = (real)(((pz + (greal)zt) + pzc) + (greal)zt * nz1);
    const real gb = Hb + low, gt = Ht + low;
#else
    const real gb = ((pz + zb) + pzc) + (zb * nz1 + low), gt = ((pz + zt) + pzc) + (zt * nz1 + low);
#endif
    const real gmin = gb < gt ? gb : gt;
    int enter = gmin < margin;
    if (gmin < (real)4.0 * margin) note_margin((double)gmin - (double)margin);
    real gthr = 0;     /* a point further than this from the plane cannot be reached within the substep at the entry speeds */
    if (enter) {
        orc_dbg_entered++;
        real hh = h + (cg < 0 ? -cg : cg), reach = R_SQRT(hh * hh + r * r);
        real vmax = (v[2] < 0 ? -v[2] : v[2]) + R_SQRT(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]) * reach;
        gthr = ((real)1.0 + (real)p->restitution) * vmax * dt + (real)1e-4;
        enter = gmin < gthr;
        note_margin((double)gmin - (double)gthr);
        if (enter) orc_dbg_canbind++;
    }
    if (!enter) { for (int i = 0; i < 18; i++) lam[i] = 0; *have_lam = 0; return; }
    const int iters = *have_lam ? p->warm_iters : p->contact_iters;
    *have_lam = 1;

    const real c[5][3] = {
        {r * ux, r * uy, zb},
        {r * ux, r * uy, zt},
        {r, (real)0.0, zb},
        {(real)-0.5 * r, (real)0.8660254037844386 * r, zb},
        {(real)-0.5 * r, (real)-0.8660254037844386 * r, zb},
    };
    /* world axes in body coordinates = rows of R; body-frame inverse inertia stays diagonal */
    const real xb[3] = {R[0], R[1], R[2]}, yb[3] = {R[3], R[4], R[5]}, nb[3] = {R[6], R[7], R[8]};
    const real Ii[3] = {(real)(1.0 / p->inertia[0]), (real)(1.0 / p->inertia[0]), (real)(1.0 / p->inertia[2])};
    const real Im[3] = {(real)p->inertia[0], (real)p->inertia[0], (real)p->inertia[2]};
    const real im = (real)(1.0 / p->mass);
    const real mu = (real)p->mu;
    const real mut[3] = {(real)p->mu_roll, (real)p->mu_roll, (real)p->mu_spin};
    const real inv_dt = (real)1.0 / dt;
    /* body-frame angular velocity wb0 = R^T w; wt carries the running value, w += R (wt - wb0) at the end */
    const real wb0[3] = {R[0] * w[0] + R[3] * w[1] + R[6] * w[2], R[1] * w[0] + R[4] * w[1] + R[7] * w[2],
                         R[2] * w[0] + R[5] * w[1] + R[8] * w[2]};
    real wt[3] = {wb0[0], wb0[1], wb0[2]};

    orc_rows J[5];
    real tgt[5], ln[5], l1[5], l2[5];
    int active[5];
    for (int i = 0; i < 5; i++) {
        point_rows(c[i], xb, yb, nb, &J[i]);
        /* height of the point above the plane: pz + c.nb = (pz + cz) + cz (nbz - 1) + cx nbx + cy nby */
#if ORC_SPLIT
        real gap = (i == 1 ? Ht : Hb) + (nb[0] * c[i][0] + nb[1] * c[i][1]);
#else
        real gap = ((pz + c[i][2]) + pzc) + (c[i][2] * nz1 + (nb[0] * c[i][0] + nb[1] * c[i][1]));
#endif
        /* the substep's manifold: points within reach (the entry rule's own bound) or still holding an impulse */
        active[i] = gap < gthr || lam[i] != 0 || lam[5 + i] != 0 || lam[10 + i] != 0;
        if (lam[i] == 0 && lam[5 + i] == 0 && lam[10 + i] == 0) note_margin((double)gap - (double)gthr);
        real vn0 = v[2] + (wb0[0] * J[i].Jn[0] + wb0[1] * J[i].Jn[1] + wb0[2] * J[i].Jn[2]);
        real rest = (vn0 < -(real)p->rest_threshold) ? (real)p->restitution * (-vn0 - (real)p->rest_threshold) : (real)0.0;
        tgt[i] = rest + (gap > 0 ? -gap * inv_dt : -(real)p->erp * gap * inv_dt);
        ln[i] = lam[i]; l1[i] = lam[5 + i]; l2[i] = lam[10 + i];
    }
    real lt[3] = {lam[16], lam[17], lam[15]};    /* torsional impulses about body x, y, z */
    /* warm start: apply the stored impulses at the current contact geometry (the targets above use the
     * velocities before this) */
    for (int i = 0; i < 5; i++) {
        if (ln[i] == 0 && l1[i] == 0 && l2[i] == 0) continue;
        v[0] += l1[i] * im; v[1] += l2[i] * im; v[2] += ln[i] * im;
        for (int k = 0; k < 3; k++) wt[k] += Ii[k] * (J[i].Jx[k] * l1[i] + J[i].Jy[k] * l2[i] + J[i].Jn[k] * ln[i]);
    }
    for (int k = 0; k < 3; k++) wt[k] += Ii[k] * lt[k];

    for (int it = 0; it < iters; it++) {
        real lsum = 0;
        int spin_done = 0;
        for (int i = 0; i < 5; i++) {
            const orc_rows *Ji = &J[i];
            /* point 0 is solved with the axial row folded in: body axis z (k = 2) is left out of the angular dynamics */
            const int nk = (i == 0) ? 2 : 3;
            real un = v[2], ux_ = v[0], uy_ = v[1];
            for (int k = 0; k < nk; k++) { un += wt[k] * Ji->Jn[k]; ux_ += wt[k] * Ji->Jx[k]; uy_ += wt[k] * Ji->Jy[k]; }
            const real unz = un + (nk == 2 ? wt[2] * Ji->Jn[2] : (real)0.0);      /* true normal velocity */
            /* a point whose normal row cannot bind and that holds no impulse is an exact no-op */
            if (!active[i] || !(tgt[i] > unz || ln[i] > 0 || l1[i] != 0 || l2[i] != 0)) continue;
            if (i) orc_dbg_blocks[3]++;
            /* the point's Delassus matrix (symmetric): A_ab = (1/m) delta_ab + sum_k Ii_k Ja_k Jb_k */
            orc_sym3 A = {im, 0, 0, im, 0, im};
            for (int k = 0; k < nk; k++) {
                const real ix = Ii[k] * Ji->Jx[k], iy = Ii[k] * Ji->Jy[k], in_ = Ii[k] * Ji->Jn[k];
                A.Axx += ix * Ji->Jx[k]; A.Axy += ix * Ji->Jy[k]; A.Axn += ix * Ji->Jn[k];
                A.Ayy += iy * Ji->Jy[k]; A.Ayn += iy * Ji->Jn[k]; A.Ann += in_ * Ji->Jn[k];
            }
            real px, py, pn;
            point_block(&A, -ux_, -uy_, tgt[i] - un, mu, l1[i], l2[i], ln[i], &px, &py, &pn);
            real dx = px - l1[i], dy = py - l2[i], dn = pn - ln[i];
            if (i == 0) {
                /* the axial impulse that keeps w_z at zero through this block, limited by mu_spin * (this point's new
                 * normal impulse + what the other points hold) */
                const real slim = mut[2] * (pn + ln[1] + ln[2] + ln[3] + ln[4]);
                const real cand = lt[2] - (wt[2] * Im[2] + (Ji->Jx[2] * dx + Ji->Jy[2] * dy + Ji->Jn[2] * dn));
                if (cand >= -slim && cand <= slim) {
                    wt[2] = 0;                       /* (exactly: the axial row sticks) */
                    lt[2] = cand;
                } else {
                    /* the axial row slips: its impulse goes to the limit of the OLD normal impulses (what a scalar row
                     * would see), and the point is solved again with the full angular dynamics */
                    const real slim0 = mut[2] * (ln[0] + ln[1] + ln[2] + ln[3] + ln[4]);
                    const real nl = r_clamp(cand, -slim0, slim0);
                    wt[2] += Ii[2] * (nl - lt[2]);
                    lt[2] = nl;
                    const real un3 = un + wt[2] * Ji->Jn[2], ux3 = ux_ + wt[2] * Ji->Jx[2], uy3 = uy_ + wt[2] * Ji->Jy[2];
                    const real ix = Ii[2] * Ji->Jx[2], iy = Ii[2] * Ji->Jy[2], in_ = Ii[2] * Ji->Jn[2];
                    A.Axx += ix * Ji->Jx[2]; A.Axy += ix * Ji->Jy[2]; A.Axn += ix * Ji->Jn[2];
                    A.Ayy += iy * Ji->Jy[2]; A.Ayn += iy * Ji->Jn[2]; A.Ann += in_ * Ji->Jn[2];
                    point_block(&A, -ux3, -uy3, tgt[i] - un3, mu, l1[i], l2[i], ln[i], &px, &py, &pn);
                    dx = px - l1[i]; dy = py - l2[i]; dn = pn - ln[i];
                    wt[2] += Ii[2] * (Ji->Jx[2] * dx + Ji->Jy[2] * dy + Ji->Jn[2] * dn);
                }
                spin_done = 1;
            }
            v[0] += dx * im; v[1] += dy * im; v[2] += dn * im;
            for (int k = 0; k < nk; k++) wt[k] += Ii[k] * (Ji->Jx[k] * dx + Ji->Jy[k] * dy + Ji->Jn[k] * dn);
            l1[i] = px; l2[i] = py; ln[i] = pn;
            lsum += pn;
        }
        /* torsional rows about the body axes x, y (and z when point 0 did not carry it): the impulse that zeroes the
         * component, limited by mu_k * (total normal impulse) */
        for (int k = 0; k < (spin_done ? 2 : 3); k++) {
            const real lim = mut[k] * lsum;
            const real nl = r_clamp(lt[k] - wt[k] * Im[k], -lim, lim);
            wt[k] += Ii[k] * (nl - lt[k]);
            lt[k] = nl;
        }
    }
    for (int i = 0; i < 5; i++) { lam[i] = ln[i]; lam[5 + i] = l1[i]; lam[10 + i] = l2[i]; }
    lam[15] = lt[2]; lam[16] = lt[0]; lam[17] = lt[1];
    /* back to the world frame: w += R (wt - wb0) */
    const real d0 = wt[0] - wb0[0], d1 = wt[1] - wb0[1], d2 = wt[2] - wb0[2];
    w[0] += R[0] * d0 + R[1] * d1 + R[2] * d2;
    w[1] += R[3] * d0 + R[4] * d1 + R[5] * d2;
    w[2] += R[6] * d0 + R[7] * d1 + R[8] * d2;
}

/* Rows B2, B4, B5, B6: one p.stepSimulation() (ref:477). */
void orc_step_simulation(const orc_body_params *p, orc_body *b, double *trace) {
    const int K = p->substeps;
    const real dt = (real)(p->dt_step / K);           /* B2: 0.02/4 = 0.005 exactly */
    const real maxv = (real)p->max_vel;
    const real Ia = (real)p->inertia[0], Ib = (real)p->inertia[1], Ic = (real)p->inertia[2];
    /* B4: applyGravity() once before the substep loop; forces constant over the K substeps */
    real F[3], T[3], pos[3], v[3], w[3];
    greal q[4], pzg = (greal)b->pos[2];       /* (pzg is pos[2] itself unless ORC_SPLIT) */
    for (int i = 0; i < 3; i++) {
        F[i] = (real)(b->force[i] + p->gravity[i] * p->mass); T[i] = (real)b->torque[i];
        pos[i] = (real)b->pos[i]; v[i] = (real)b->vel[i]; w[i] = (real)b->omega[i];
    }
    for (int i = 0; i < 4; i++) q[i] = (greal)b->quat[i];
    real lam[18];
    for (int i = 0; i < 18; i++) lam[i] = 0;   /* cold start at every control step */
    int have_lam = 0;                          /* the previous substep ran the solve (warm start) */
    greal pzc = 0;                             /* compensation of the height update, true height = pos[2] + pzc */

    const int follow = b->thrust_local[0] != 0 || b->thrust_local[1] != 0 || b->thrust_local[2] != 0;
    const real F0[3] = {F[0], F[1], F[2]}, T0[3] = {T[0], T[1], T[2]};
    for (int k = 0; k < K; k++) {
        real R[9];
        const greal nz1 = r_matrix_from_quat(q, R);
        if (follow) {   /* quirk Q3 cleared: the body-fixed thrust and its torque at this substep's attitude */
            const real fl[3] = {(real)b->thrust_local[0], (real)b->thrust_local[1], (real)b->thrust_local[2]};
            const real tl[3] = {-(real)b->thrust_arm * fl[1], (real)b->thrust_arm * fl[0], (real)0.0};
            for (int i = 0; i < 3; i++) {
                F[i] = F0[i] + (R[3 * i] * fl[0] + R[3 * i + 1] * fl[1] + R[3 * i + 2] * fl[2]);
                T[i] = T0[i] + (R[3 * i] * tl[0] + R[3 * i + 1] * tl[1]);
            }
        }
        /* B5: ABA for a lone floating base, in base-local coordinates */
        real wl[3], tl[3], wdl[3], wd[3];
        wl[0] = R[0] * w[0] + R[3] * w[1] + R[6] * w[2];
        wl[1] = R[1] * w[0] + R[4] * w[1] + R[7] * w[2];
        wl[2] = R[2] * w[0] + R[5] * w[1] + R[8] * w[2];
        tl[0] = R[0] * T[0] + R[3] * T[1] + R[6] * T[2];
        tl[1] = R[1] * T[0] + R[4] * T[1] + R[7] * T[2];
        tl[2] = R[2] * T[0] + R[5] * T[1] + R[8] * T[2];
        real wn2 = wl[0] * wl[0] + wl[1] * wl[1] + wl[2] * wl[2];
        real wn = wn2 > R_EPS2 ? R_SQRT(wn2) : (real)0.0;      /* btVector3::safeNorm */
        real kd = (real)p->ang_damp + (real)p->ang_damp * wn;
        real zacc[3] = {-tl[0] + Ia * wl[0] * kd, -tl[1] + Ib * wl[1] * kd, -tl[2] + Ic * wl[2] * kd};
        if (p->use_gyro) {
            real Iw[3] = {Ia * wl[0], Ib * wl[1], Ic * wl[2]};
            zacc[0] += wl[1] * Iw[2] - wl[2] * Iw[1];
            zacc[1] += wl[2] * Iw[0] - wl[0] * Iw[2];
            zacc[2] += wl[0] * Iw[1] - wl[1] * Iw[0];
        }
        wdl[0] = -(zacc[0] / Ia); wdl[1] = -(zacc[1] / Ib); wdl[2] = -(zacc[2] / Ic);
        wd[0] = R[0] * wdl[0] + R[1] * wdl[1] + R[2] * wdl[2];
        wd[1] = R[3] * wdl[0] + R[4] * wdl[1] + R[5] * wdl[2];
        wd[2] = R[6] * wdl[0] + R[7] * wdl[1] + R[8] * wdl[2];
        real vn2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        real vn = vn2 > R_EPS2 ? R_SQRT(vn2) : (real)0.0;
        real kl = (real)p->lin_damp + (real)p->lin_damp * vn;
        for (int i = 0; i < 3; i++)      /* applyDeltaVeeMultiDof with the +-100 clamp */
            w[i] = r_clamp(w[i] + wd[i] * dt, -maxv, maxv);
        for (int i = 0; i < 3; i++) {
            real vd = F[i] / (real)p->mass - v[i] * kl;
            v[i] = r_clamp(v[i] + vd * dt, -maxv, maxv);
        }
        /* B9 (our model): contacts detected at the pre-integration pose, solved on velocities */
        if (p->ground) solve_contacts(p, dt, R, nz1, pzg, pzc, v, w, lam, &have_lam);

        /* B6: stepPositionsMultiDof -- semi-implicit Euler, exponential map */
        pos[0] += dt * v[0]; pos[1] += dt * v[1];
        {   /* height with a running compensation (Kahan): the contact targets divide the gap by dt, so the 3e-8
             * rounding of pz + dt vz per substep would otherwise show up as 1.5e-5 m/s in fp32 at dt = 0.002 */
#if ORC_SPLIT   /* (the kernel keeps the float height + float compensation and adds them in double: same value) */
            const real pc = (real)pzc, y = dt * v[2] + pc, t = pos[2] + y;
            pzc = (greal)(real)(y - (t - pos[2]));
            pos[2] = t; pzg = (greal)t;
#else
            const real y = dt * v[2] + pzc, t = pos[2] + y;
            pzc = y - (t - pos[2]);
            pos[2] = t; pzg = t;
#endif
        }
        real ang = R_SQRT(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        if (ang * dt > (real)(0.25 * PI_D)) ang = (real)(0.5 * (0.5 * PI_D)) / dt;   /* ANGULAR_MOTION_THRESHOLD */
        real sc;
        if (ang < (real)0.001) sc = (real)0.5 * dt - (dt * dt * dt) * (real)0.020833333333 * ang * ang;
        else sc = R_SIN((real)0.5 * ang * dt) / ang;
        real ax0 = w[0] * sc, ax1 = w[1] * sc, ax2 = w[2] * sc;
        real cw = R_COS(ang * dt * (real)0.5);
        /* q <- dq (x) q  with dq = (ax, cw) */
        /* (greal == real unless ORC_SPLIT: then the product runs in double on the float increment, as in the kernel) */
        greal nx = (greal)cw * q[0] + (greal)ax0 * q[3] + (greal)ax1 * q[2] - (greal)ax2 * q[1];
        greal ny = (greal)cw * q[1] + (greal)ax1 * q[3] + (greal)ax2 * q[0] - (greal)ax0 * q[2];
        greal nz = (greal)cw * q[2] + (greal)ax2 * q[3] + (greal)ax0 * q[1] - (greal)ax1 * q[0];
        greal nw = (greal)cw * q[3] - (greal)ax0 * q[0] - (greal)ax1 * q[1] - (greal)ax2 * q[2];
#if ORC_SPLIT
        greal inv = 1.0 / sqrt(nx * nx + ny * ny + nz * nz + nw * nw);
#else
        real inv = (real)1.0 / R_SQRT(nx * nx + ny * ny + nz * nz + nw * nw);
#endif
        q[0] = nx * inv; q[1] = ny * inv; q[2] = nz * inv; q[3] = nw * inv;

        if (trace) {
            double *t = trace + 13 * k;
            for (int i = 0; i < 3; i++) { t[i] = pos[i]; t[7 + i] = v[i]; t[10 + i] = w[i]; }
            for (int i = 0; i < 4; i++) t[3 + i] = q[i];
        }
    }
    for (int i = 0; i < 3; i++) { b->pos[i] = pos[i]; b->vel[i] = v[i]; b->omega[i] = w[i]; }
    for (int i = 0; i < 4; i++) b->quat[i] = q[i];
    /* B4: clearForces() after the loop */
    for (int i = 0; i < 3; i++) { b->force[i] = 0; b->torque[i] = 0; b->thrust_local[i] = 0; }
    b->thrust_arm = 0;
}

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., Random123) -- counter-based RNG       */
/* ------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { ST_DR_A = 1, ST_DR_B = 2, ST_NOISE_A = 3, ST_NOISE_B = 4, ST_ACTION = 5, ST_ACTOR = 6, ST_DR_C = 7 };

static void draw4(uint64_t seed, int64_t gid, uint32_t stream, uint32_t a, uint32_t b, uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(((uint64_t)gid >> 32) & 0xFFFFu) | (stream << 16), a, b};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(ctr, key, out);
}
/* 23-bit uniform in (0,1), (k + 0.5) / 2^23: exact in float32 and float64 */
static inline double u01(uint32_t x) { return ((double)(x >> 9) + 0.5) * (1.0 / 8388608.0); }
static inline void box_muller(uint32_t x0, uint32_t x1, double *n0, double *n1) {
    double r = sqrt(-2.0 * log(u01(x0)));
    double th = 2.0 * PI_D * u01(x1);
    *n0 = r * cos(th); *n1 = r * sin(th);
}

/* ------------------------------------------------------------------ */
/* Env layer                                                           */
/* ------------------------------------------------------------------ */
struct orc_sim {
    orc_config cfg;
    int64_t n;
    orc_env *envs;
    double stats[ORC_NSTATS];
    double *trace;  /* K x 13, env 0 */
    double fuel_tab[1002];
};

/* Row S4 (ref:530-533): fuel after n decrements, by repeated fp64 subtraction. */
static void build_fuel_table(double *t) {
    double f = 1.0;
    t[0] = f;
    for (int n = 1; n <= 1001; n++) { f = f - 0.001; if (f < 0) f = 0; t[n] = f; }
}
double orc_fuel_table(int n) {
    static double tab[1002]; static int init = 0;
    if (!init) { build_fuel_table(tab); init = 1; }
    if (n < 0) n = 0;
    if (n > 1001) n = 1001;
    return tab[n];
}

/* Contract X thrust curve (our definition; the reference's thrust is the constant ref:463).
 * Multiplier as a function of burn fraction u = burn/1000: ignition spike to 1.3 at 5 %,
 * settling to 1.0 at 15 %, flat, linear tail-off to 0.5 over the last 10 %. */
double orc_thrust_curve(int mode, int burn) {
    if (mode == 0) return 1.0;
    double u = burn * 0.001;
    if (u < 0.05) return 1.0 + 6.0 * u;
    if (u < 0.15) return 1.3 - 3.0 * (u - 0.05);
    if (u < 0.9) return 1.0;
    return 1.0 - 5.0 * (u - 0.9);
}

void orc_config_default(orc_config *c, int contract) {
    memset(c, 0, sizeof(*c));
    c->contract = contract;
    c->substeps = contract == ORC_CONTRACT_R ? 4 : 10;
    c->max_episode_steps = 1000;
    c->autoreset = 0;
    c->quirks = contract == ORC_CONTRACT_R ? ORC_Q_ALL_REFERENCE : ORC_Q_CONTRACT_X;
    c->diversity_mode = contract == ORC_CONTRACT_R ? ORC_DIV_EXACT : ORC_DIV_FAST;
    c->contact_iters = 2;
    c->contact_warm_iters = 1;
    c->ground = 1;
    c->dt_step = 0.02;
    c->gradient_penalty = 0.1; c->diversity_bonus = 0.05;      /* ref:83-84 defaults */
    c->mass = 2.0; c->radius = 0.05; c->length = 1.0; c->thrust = 35.0;   /* ref:412-414, 463 */
    c->gimbal_max_rad = 18.0 * (PI_D / 180.0);                 /* ref:471 np.radians(18.0) */
    c->lin_damp = 0.01; c->ang_damp = 0.02;                    /* ref:453-454 */
    if (contract == ORC_CONTRACT_X) {
        /* config/config.yaml:340-349 */
        c->mass_variation = 0.3;
        c->thrust_std = 0.2; c->thrust_lo = 0.4; c->thrust_hi = 1.6;
        c->cg_offset_max = 0.1;
        c->wind_std = 3.0;
        c->sensor_noise_std = 0.02;
        c->init_tilt_max = 0.0; c->init_omega_max = 0.0;
        c->propellant_fraction = 0.0; c->cg_burn_shift = 0.0;
        c->delay_steps = 0; c->thrust_curve = 0;
    }
    c->thrust_lo = c->thrust_lo ? c->thrust_lo : 0.4;
    c->thrust_hi = c->thrust_hi ? c->thrust_hi : 1.6;
    c->seed = 42;
    c->env_id_base = 0;
    {   /* contact material: ref:349-352 (plane) x ref:455-458 (rocket), Bullet's combination rules */
        orc_body_params p;
        orc_body_params_default(&p);
        c->contact_mu = p.mu; c->contact_mu_spin = p.mu_spin; c->contact_mu_roll = p.mu_roll;
        c->contact_restitution = p.restitution; c->contact_rest_threshold = p.rest_threshold;
        c->contact_erp = p.erp; c->contact_margin = p.margin;
    }
}

static void env_params(const orc_config *c, const orc_env *e, orc_body_params *p) {
    orc_body_params_default(p);
    double ms = (c->contract == ORC_CONTRACT_X) ? e->mass_scale : 1.0;
    double burnt = 1.0 - e->fuel;
    double m = c->mass * ms * (1.0 - c->propellant_fraction * burnt);
    double cg = (c->contract == ORC_CONTRACT_X) ? e->cg_offset + c->cg_burn_shift * burnt : 0.0;
    p->mass = m;
    p->radius = c->radius; p->half_len = 0.5 * c->length;
    p->inertia[0] = p->inertia[1] = (1.0 / 12.0) * m * (3 * c->radius * c->radius + c->length * c->length) + m * cg * cg;
    p->inertia[2] = (1.0 / 2.0) * m * c->radius * c->radius;
    p->cg = cg;
    p->lin_damp = c->lin_damp; p->ang_damp = c->ang_damp;
    p->substeps = c->substeps; p->dt_step = c->dt_step;
    p->ground = c->ground; p->contact_iters = c->contact_iters; p->warm_iters = c->contact_warm_iters;
    p->mu = c->contact_mu; p->mu_spin = c->contact_mu_spin; p->mu_roll = c->contact_mu_roll;
    p->restitution = c->contact_restitution; p->rest_threshold = c->contact_rest_threshold;
    p->erp = c->contact_erp; p->margin = c->contact_margin;
}

/* ref:587-606 _get_enhanced_observation */
static void build_obs(const orc_config *c, const orc_env *e, int64_t gid, int phase_for_obs, float obs[10]) {
    double orn[4];
    orc_reported_quat(e->body.quat, orn);
    double o[10] = {orn[0], orn[1], orn[2], orn[3], e->body.omega[0], e->body.omega[1], e->body.omega[2],
                    e->fuel, (double)phase_for_obs / 7.0,
                    fmin(1.0, (double)e->step / (double)c->max_episode_steps)};
    if (c->contract == ORC_CONTRACT_X && c->sensor_noise_std > 0) {
        /* eight N(0,1) draws from ONE Philox block: each 32-bit word gives two 16-bit uniforms (k + 0.5) / 2^16 (radius from
         * the low half, angle from the high half) -> four Box-Muller pairs, |n| <= 4.9 sigma (sensor noise, not a tail study) */
        uint32_t a[4]; double n[8];
        draw4(c->seed, gid, ST_NOISE_A, (uint32_t)e->episode, (uint32_t)e->step, a);
        for (int k = 0; k < 4; k++) {
            double u0 = ((double)(a[k] & 0xFFFFu) + 0.5) * (1.0 / 65536.0), u1 = ((double)(a[k] >> 16) + 0.5) * (1.0 / 65536.0);
            double r = sqrt(-2.0 * log(u0)), th = 2.0 * PI_D * u1;
            n[2 * k] = r * cos(th); n[2 * k + 1] = r * sin(th);
        }
        for (int i = 0; i < 7; i++) o[i] += c->sensor_noise_std * n[i];
    }
    for (int i = 0; i < 10; i++) obs[i] = (float)o[i];
}

/* ref:381-407 reset + ref:409-464 _create_enhanced_rocket (row S13, quirks Q10, Q11, Q15) */
static void env_reset(const orc_config *c, orc_env *e, int64_t gid, int first_time) {
    double pos[3] = {0, 0, 1.0}, quat[4] = {0, 0, 0, 1};
    e->episode += 1;
    orc_body_init(&e->body, pos, quat);
    e->fuel = 1.0; e->burn = 0;
    e->step = 0; e->phase = 0; e->success = 0;
    e->ep_return = 0;
    if (first_time || !(c->quirks & ORC_Q_KEEP_CRITERIA)) { e->consec = 0; e->crit_pushes = 0; }
    if (first_time || !(c->quirks & ORC_Q_KEEP_REWARD_HIST)) {
        e->has_prev = 0; e->prev_action[0] = e->prev_action[1] = 0;
        e->hist_count = 0; e->n_clip = 0; e->n_run = 0;
        memset(e->clip_bits, 0, sizeof(e->clip_bits)); memset(e->run_bits, 0, sizeof(e->run_bits));
    }
    e->mass_scale = 1; e->thrust_scale = 1; e->cg_offset = 0; e->wind[0] = e->wind[1] = 0;
    memset(e->delay_ring, 0, sizeof(e->delay_ring));
    if (c->contract == ORC_CONTRACT_X) {
        uint32_t a[4], b[4], d[4]; double n0, n1, n2, n3;
        draw4(c->seed, gid, ST_DR_A, (uint32_t)e->episode, 0, a);
        draw4(c->seed, gid, ST_DR_B, (uint32_t)e->episode, 0, b);
        draw4(c->seed, gid, ST_DR_C, (uint32_t)e->episode, 0, d);
        box_muller(a[1], a[2], &n0, &n1);
        box_muller(b[0], b[1], &n2, &n3);
        e->mass_scale = 1.0 + c->mass_variation * (2.0 * u01(a[0]) - 1.0);
        e->thrust_scale = clampd(1.0 + c->thrust_std * n0, c->thrust_lo, c->thrust_hi);
        e->wind[0] = c->wind_std * n1;
        e->wind[1] = c->wind_std * n2;
        e->cg_offset = c->cg_offset_max * (2.0 * u01(a[3]) - 1.0);
        double tx = c->init_tilt_max * (2.0 * u01(b[2]) - 1.0);
        double ty = c->init_tilt_max * (2.0 * u01(b[3]) - 1.0);
        double ang = sqrt(tx * tx + ty * ty);
        double sc = ang < 1e-6 ? 0.5 - ang * ang / 48.0 : sin(0.5 * ang) / ang;
        e->body.quat[0] = tx * sc; e->body.quat[1] = ty * sc; e->body.quat[2] = 0; e->body.quat[3] = cos(0.5 * ang);
        for (int i = 0; i < 3; i++) e->body.omega[i] = c->init_omega_max * (2.0 * u01(d[i]) - 1.0);
    }
}

/* numpy pairwise sum of exactly 10 doubles (np.add.reduce, n=10: 8-lane block + 2 tail) */
static double np_sum10(const double *a) {
    double r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    r += a[8]; r += a[9];
    return r;
}

static inline float clipf1(float x) { return x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x); }

/* number of distinct values among the stored reward history (ref:221 len(set(history))) */
static int distinct_exact(const orc_env *e) {
    int len = e->hist_count < ORC_HIST ? (int)e->hist_count : ORC_HIST;
    int distinct = 0;
    for (int i = 0; i < len; i++) {
        int dup = 0;
        for (int j = 0; j < i; j++) if (e->hist[j] == e->hist[i]) { dup = 1; break; }
        distinct += !dup;
    }
    return distinct;
}

static inline int bit_get(const uint32_t *w, int pos) { return (w[pos >> 5] >> (pos & 31)) & 1u; }
static inline void bit_set(uint32_t *w, int pos, int v) {
    if (v) w[pos >> 5] |= (1u << (pos & 31)); else w[pos >> 5] &= ~(1u << (pos & 31));
}

/* push into reward_history (ref:123) with the fast duplicate bookkeeping:
 *   clip bit : value == -1000 (the only clip value that occurs; +200 is unreachable)
 *   run bit  : value == immediately preceding value (and not a clip value)
 * distinct = len - n_run - n_clip + (n_clip > 0).  Exact for clip duplicates and runs;
 * differs from len(set()) only for non-adjacent coincidences. */
static void hist_push(orc_env *e, double v) {
    int slot = (int)(e->hist_count % ORC_HIST);
    if (e->hist_count >= ORC_HIST) {
        /* the entry in `slot` leaves the window */
        if (bit_get(e->clip_bits, slot)) e->n_clip--;
        if (bit_get(e->run_bits, slot)) e->n_run--;
        /* the new oldest entry can no longer be a run-duplicate of its predecessor */
        int nxt = (slot + 1) % ORC_HIST;
        if (bit_get(e->run_bits, nxt)) { bit_set(e->run_bits, nxt, 0); e->n_run--; }
    }
    int is_clip = (v == -1000.0);
    int is_run = 0;
    if (!is_clip && e->hist_count > 0) {
        int prev = (int)((e->hist_count - 1) % ORC_HIST);
        is_run = (e->hist[prev] == v);
    }
    /* a window of length 1 about to be overwritten has no predecessor inside the window */
    bit_set(e->clip_bits, slot, is_clip); bit_set(e->run_bits, slot, is_run);
    e->n_clip += is_clip; e->n_run += is_run;
    e->hist[slot] = v;
    e->hist_count++;
}

static void env_step(orc_sim *s, orc_env *e, int64_t gid, const float *act, orc_step_out *o, double *trace,
                     double *stats) {
    const orc_config *c = &s->cfg;
    const int X = c->contract == ORC_CONTRACT_X;
    orc_step_out local;
    if (!o) o = &local;
    memset(o, 0, sizeof(*o));

    /* ---- S2 (ref:470-471): clip, scale to gimbal angle ---- */
    float a[2] = {clipf1(act[0]), clipf1(act[1])};
    float ap[2] = {a[0], a[1]};
    if (X && c->delay_steps > 0) {               /* actuator delay ring (Contract X) */
        int D = c->delay_steps;
        ap[0] = e->delay_ring[0][0]; ap[1] = e->delay_ring[0][1];
        for (int i = 0; i + 1 < D; i++) { e->delay_ring[i][0] = e->delay_ring[i + 1][0]; e->delay_ring[i][1] = e->delay_ring[i + 1][1]; }
        e->delay_ring[D - 1][0] = a[0]; e->delay_ring[D - 1][1] = a[1];
    }
    double pitch = (double)ap[0] * c->gimbal_max_rad;
    double yaw = (double)ap[1] * c->gimbal_max_rad;

    /* ---- S3 (ref:520-559) _apply_enhanced_control ---- */
    orc_body_params P;
    env_params(c, e, &P);            /* mass/inertia/cg from the pre-step fuel */
    orc_body *b = &e->body;
    double orn[4], R[9];
    orc_reported_quat(b->quat, orn);
    if (c->quirks & ORC_Q_DOUBLE_GRAVITY) {       /* Q1 (ref:525-527) */
        double g[3] = {0, 0, -9.81 * P.mass};
        orc_apply_external_force(b, g, b->pos);
    }
    if (e->fuel > 0) {                            /* Q4 (ref:530-533) */
        int burn_before = e->burn;
        e->burn += 1;
        e->fuel = s->fuel_tab[e->burn > 1001 ? 1001 : e->burn];
        double T = c->thrust;
        if (X) T = T * e->thrust_scale * orc_thrust_curve(c->thrust_curve, burn_before);
        double fl[3] = {T * sin(yaw), T * sin(pitch), T * cos(pitch) * cos(yaw)};   /* Q2 (ref:539-543) */
        if (!(c->quirks & ORC_Q_THRUST_VECTOR)) {                                   /* Q2 cleared: same direction, |F| = T */
            double nrm = sqrt(fl[0] * fl[0] + fl[1] * fl[1] + fl[2] * fl[2]);
            for (int i = 0; i < 3; i++) fl[i] = fl[i] * (T / nrm);
        }
        double off[3] = {0, 0, -(P.half_len + P.cg)};
        if (c->quirks & ORC_Q_FROZEN_FORCES) {                                      /* Q3: evaluated once, world frame */
            orc_matrix_from_quat(orn, R);
            double fw[3], rel[3], tp[3];
            matvec(R, fl, fw);
            matvec(R, off, rel);
            for (int i = 0; i < 3; i++) tp[i] = b->pos[i] + rel[i];                 /* ref:550 */
            orc_apply_external_force(b, fw, tp);
        } else {                                                                    /* Q3 cleared: the thrust follows the body */
            for (int i = 0; i < 3; i++) b->thrust_local[i] = fl[i];
            b->thrust_arm = off[2];
        }
    }
    /* ---- S5 (ref:561-585) _apply_aerodynamics ---- */
    {
        double rho = 1.225 * exp(-b->pos[2] / 8400);
        double vmag = sqrt(b->vel[0] * b->vel[0] + b->vel[1] * b->vel[1] + b->vel[2] * b->vel[2]);
        if (vmag > 0.1 || (!(c->quirks & ORC_Q_DRAG_CUTOFF) && vmag > 0)) {   /* Q5 */
            double area = PI_D * (0.05 * 0.05);
            double dm = 0.5 * rho * (vmag * vmag) * 0.47 * area;
            double df[3];
            for (int i = 0; i < 3; i++) df[i] = dm * (-b->vel[i] / vmag);
            orc_apply_external_force(b, df, b->pos);
        }
        if (c->quirks & ORC_Q_STACKED_DAMPING) {   /* Q6 */
            double ad = 0.02 * rho;
            double dt3[3] = {-ad * b->omega[0], -ad * b->omega[1], -ad * b->omega[2]};
            orc_apply_external_torque(b, dt3);
        }
    }
    if (X) {                                       /* wind: constant world-frame force per episode */
        double w[3] = {e->wind[0], e->wind[1], 0};
        orc_apply_external_force(b, w, b->pos);
    }

    /* ---- S6 (ref:477) ---- */
    orc_tl_margin = 1e30;
    orc_step_simulation(&P, b, trace);
    o->contact_margin = orc_tl_margin;
    e->step += 1;                                  /* ref:478 */

    /* ---- S7 (ref:608-633) _get_state_dict ---- */
    double rpy[3];
    orc_reported_quat(b->quat, orn);
    orc_euler_from_quat(orn, rpy);
    double tilt = sqrt(rpy[1] * rpy[1] + rpy[2] * rpy[2]);                 /* Q7 */
    if (!(c->quirks & ORC_Q_EULER_TILT)) {                                 /* Q7 cleared: angle between body axis and vertical */
        double sxy = 2.0 * sqrt((orn[0] * orn[0] + orn[1] * orn[1]) * (orn[2] * orn[2] + orn[3] * orn[3]));
        tilt = atan2(sxy, 1.0 - 2.0 * (orn[0] * orn[0] + orn[1] * orn[1]));
    }
    double wmag = sqrt(b->omega[0] * b->omega[0] + b->omega[1] * b->omega[1] + b->omega[2] * b->omega[2]);
    double vh = sqrt(b->vel[0] * b->vel[0] + b->vel[1] * b->vel[1]);
    double vv = fabs(b->vel[2]);
    double alt = b->pos[2];
    int crashed = alt < 0.1;                                               /* Q17 */
    if (!(c->quirks & ORC_Q_CRASH_IS_COM_HEIGHT)) {                        /* Q17 cleared: a hard or tilted touchdown */
        double cgx = X ? e->cg_offset + c->cg_burn_shift * (1.0 - e->fuel) : 0.0, hl = 0.5 * c->length;
        double R31 = 2.0 * (orn[0] * orn[2] - orn[3] * orn[1]), R32 = 2.0 * (orn[1] * orn[2] + orn[3] * orn[0]);
        double R33 = 1.0 - 2.0 * (orn[0] * orn[0] + orn[1] * orn[1]);
        double low = alt + fmin(R33 * (-hl - cgx), R33 * (hl - cgx)) - c->radius * sqrt(R31 * R31 + R32 * R32);
        crashed = low < 0.01 && (b->vel[2] < -2.0 || tilt > 0.52);
    }
    int phase_pre = e->phase, success_pre = e->success;                    /* Q8, Q9 */

    /* ---- S9 (ref:635-657) _update_mission_phase ---- */
    if (e->phase == 0 && e->fuel < 0.8) e->phase = 1;
    else if (e->phase == 1 && alt < 5.0) e->phase = 2;
    else if (e->phase == 2 && alt < 1.0) e->phase = 3;
    else if (e->phase == 3 && alt < 0.5) {
        if (tilt < 0.087 && wmag < 0.1) { e->phase = 5; e->success = 1; }
    }
    /* ---- S10 (ref:659-695) _check_mission_success ---- */
    if (!e->success) {
        int all_met = (tilt < 0.087) && (vv < 2.0 && vh < 0.5) && (0.2 <= alt && alt <= 2.0) && (wmag < 0.1);
        e->consec = all_met ? (e->consec < 1000000 ? e->consec + 1 : e->consec) : 0;
        e->crit_pushes += e->crit_pushes < 1000000;
        if (e->consec >= 100) e->success = 1;      /* deque(maxlen=100) full and all true */
    }
    int phase_r = (c->quirks & ORC_Q_LAGGED_PHASE) ? phase_pre : e->phase;
    int success_r = (c->quirks & ORC_Q_LAGGED_PHASE) ? success_pre : e->success;

    /* ---- S8 (ref:587-606): obs uses the lagged phase (Q8) ---- */
    build_obs(c, e, gid, phase_r, o->obs);

    /* ---- R1-R10 (ref:86-224) MultiObjectiveReward.compute_reward ---- */
    double comp[ORC_NCOMP] = {0};
    comp[0] = (success_r ? 1.0 : (phase_r == 2 ? 0.1 : 0.0)) * 100.0;                        /* R1 */
    {
        double tp = exp(-10 * fmax(0.0, tilt - 0.087));
        double apn = exp(-5 * fmax(0.0, wmag - 0.1));
        double alp = (0.2 <= alt && alt <= 20.0) ? 1.0 : 0.5;
        comp[1] = ((tp + apn + alp) / 3.0) * 50.0;                                           /* R2 */
    }
    float ce = sqrtf(a[0] * a[0] + a[1] * a[1]);            /* np.linalg.norm(float32[2]) -> float32 */
    if (e->fuel > 0.1 && ce < 0.5f) {
        float fe = (float)e->fuel * (1.0f - ce);             /* NumPy 2 weak-scalar promotion: float32 */
        comp[2] = (double)(fe * 20.0f);                                                      /* R3 */
    } else comp[2] = 0.0 * 20.0;
    comp[3] = ((tilt < 0.05 && wmag < 0.1) ? 1.0 : ((tilt < 0.1 && wmag < 0.2) ? 0.5 : 0.0)) * 10.0;   /* R4 */
    if (e->has_prev) {                                                                        /* R5, Q11 */
        float d0 = a[0] - e->prev_action[0], d1 = a[1] - e->prev_action[1];
        float ad = sqrtf(d0 * d0 + d1 * d1);
        float sm = expf(-5.0f * ad);
        comp[4] = (double)(sm * 5.0f);
    } else comp[4] = 1.0 * 5.0;
    e->prev_action[0] = a[0]; e->prev_action[1] = a[1]; e->has_prev = 1;
    comp[5] = exp(-2 * fabs(alt - 3.0)) * 5.0;                                                /* R6 */
    int has_crash = crashed, has_tilt = tilt > 0.52, has_sat = ce > 0.9f;                     /* R7 */
    if (has_crash) comp[6] = -1000.0;
    if (has_tilt) comp[7] = -500.0 * (tilt - 0.52);
    if (has_sat) comp[8] = (double)(-50.0f * (ce - 0.9f));
    /* R8, R9 (ref:209-224) */
    double adj = 0.0;
    int64_t hc = e->hist_count;
    int len = hc < ORC_HIST ? (int)hc : ORC_HIST;
    if (len > 10) {
        double r10[10], d10[10];
        for (int i = 0; i < 10; i++) r10[i] = e->hist[(hc - 10 + i) % ORC_HIST];
        double mean = np_sum10(r10) / 10.0;
        for (int i = 0; i < 10; i++) { double d = r10[i] - mean; d10[i] = d * d; }
        double var = np_sum10(d10) / 10.0;
        if (var > 10000 && (c->quirks & ORC_Q_VARIANCE_PENALTY)) adj -= c->gradient_penalty * var;
    }
    int div_flag = 0;
    if (c->diversity_mode == ORC_DIV_EXACT) div_flag = (double)distinct_exact(e) > len * 0.8;
    else if (c->diversity_mode == ORC_DIV_FAST) {
        int distinct = len - e->n_run - e->n_clip + (e->n_clip > 0);
        div_flag = (double)distinct > len * 0.8;
    }
    if (!(c->quirks & ORC_Q_DIVERSITY_BONUS)) div_flag = 0;
    if (div_flag) adj += c->diversity_bonus;
    /* R10: Python sum() over the dict in insertion order, then + adjustment, clip, append */
    double total = 0;
    total = total + comp[0]; total = total + comp[1]; total = total + comp[2];
    total = total + comp[3]; total = total + comp[4]; total = total + comp[5];
    if (has_crash) total = total + comp[6];
    if (has_tilt) total = total + comp[7];
    if (has_sat) total = total + comp[8];
    total = total + adj;
    comp[9] = adj; comp[10] = total; comp[11] = div_flag;
    double reward = clampd(total, -1000.0, 200.0);
    hist_push(e, reward);
    memcpy(o->comp, comp, sizeof(comp));
    o->reward = reward;
    e->ep_return += reward;

    /* ---- S11 (ref:697-721) _check_termination ---- */
    int terminated = 0, truncated = 0, reason = 0;
    if (e->success) {                                                     /* Q16 */
        terminated = 1; reason = 1;
        if (!(c->quirks & ORC_Q_SUCCESS_MASKS_TRUNCATION) && e->step >= c->max_episode_steps) truncated = 1;
    } else {
        if (crashed) { terminated = 1; reason = 2; }
        else if (tilt > 0.52) { terminated = 1; reason = 3; }
        else if (alt > 20.0) { terminated = 1; reason = 4; }
        else if (sqrt(b->pos[0] * b->pos[0] + b->pos[1] * b->pos[1]) > 50.0) { terminated = 1; reason = 5; }
        if (e->step >= c->max_episode_steps) truncated = 1;
    }
    o->terminated = terminated; o->truncated = truncated; o->term_reason = reason;

    /* ---- S12 (ref:723-742) info ---- */
    o->altitude = alt; o->tilt = tilt; o->omega_mag = wmag; o->fuel = e->fuel; o->vh = vh; o->vv = vv;
    memcpy(o->position, b->pos, 24);
    o->phase = e->phase; o->success = e->success; o->step = e->step;
    o->criteria_met = (e->crit_pushes >= 10) && (e->consec >= 10);        /* Q18 */

    /* ---- episode statistics ---- */
    stats[14] += 1;
    {
        double tilt_deg = tilt * (180.0 / PI_D);
        int viol = (tilt_deg > 0.52 * (180.0 / PI_D)) || (wmag > 5.0) || (alt < 0.1) || (alt > 20.0);
        stats[10] += viol;
    }
    if (terminated || truncated) {
        stats[0] += 1; stats[1] += e->ep_return; stats[2] += e->ep_return * e->ep_return; stats[3] += e->step;
        stats[4] += e->success;
        stats[5] += reason == 2; stats[6] += reason == 3; stats[7] += reason == 4; stats[8] += reason == 5;
        stats[9] += truncated;
        stats[11] += alt; stats[12] += tilt; stats[13] += e->fuel;
        if (c->autoreset) {
            memcpy(o->final_obs, o->obs, sizeof(o->obs));
            env_reset(c, e, gid, 0);
            build_obs(c, e, gid, 0, o->obs);
        }
    }
}

orc_sim *orc_create(const orc_config *c, int64_t n) {
    orc_sim *s = (orc_sim *)calloc(1, sizeof(orc_sim));
    s->cfg = *c; s->n = n;
    s->envs = (orc_env *)calloc((size_t)n, sizeof(orc_env));
    s->trace = (double *)calloc((size_t)(c->substeps > 0 ? c->substeps : 1) * 13, sizeof(double));
    build_fuel_table(s->fuel_tab);
    for (int64_t i = 0; i < n; i++) { s->envs[i].episode = -1; env_reset(&s->cfg, &s->envs[i], c->env_id_base + i, 1); }
    return s;
}
void orc_destroy(orc_sim *s) { if (!s) return; free(s->envs); free(s->trace); free(s); }
int64_t orc_num_envs(const orc_sim *s) { return s->n; }
orc_env *orc_env_ptr(orc_sim *s, int64_t i) { return &s->envs[i]; }
orc_config *orc_config_ptr(orc_sim *s) { return &s->cfg; }
const double *orc_last_trace(const orc_sim *s) { return s->trace; }

void orc_reset(orc_sim *s, const uint8_t *mask, float *obs_out) {
    for (int64_t i = 0; i < s->n; i++) {
        if (mask && !mask[i]) continue;
        int64_t gid = s->cfg.env_id_base + i;
        env_reset(&s->cfg, &s->envs[i], gid, 0);
        if (obs_out) build_obs(&s->cfg, &s->envs[i], gid, 0, obs_out + 10 * i);
    }
}

void orc_step(orc_sim *s, const float *actions, orc_step_out *outs, int threads) {
    const int64_t n = s->n;
    if (threads <= 1 || n < 2) {
        for (int64_t i = 0; i < n; i++)
            env_step(s, &s->envs[i], s->cfg.env_id_base + i, actions + 2 * i, outs ? outs + i : NULL,
                     i == 0 ? s->trace : NULL, s->stats);
        return;
    }
#ifdef _OPENMP
    double acc[ORC_NSTATS] = {0};
#pragma omp parallel num_threads(threads)
    {
        double loc[ORC_NSTATS] = {0};
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; i++)
            env_step(s, &s->envs[i], s->cfg.env_id_base + i, actions + 2 * i, outs ? outs + i : NULL,
                     i == 0 ? s->trace : NULL, loc);
#pragma omp critical
        for (int k = 0; k < ORC_NSTATS; k++) acc[k] += loc[k];
    }
    for (int k = 0; k < ORC_NSTATS; k++) s->stats[k] += acc[k];
#else
    for (int64_t i = 0; i < n; i++)
        env_step(s, &s->envs[i], s->cfg.env_id_base + i, actions + 2 * i, outs ? outs + i : NULL,
                 i == 0 ? s->trace : NULL, s->stats);
#endif
}

void orc_step_arrays(orc_sim *s, const float *actions, float *obs, double *reward, uint8_t *terminated,
                     uint8_t *truncated, float *final_obs, int threads) {
    const int64_t n = s->n;
    orc_step_out *outs = (orc_step_out *)malloc(sizeof(orc_step_out) * (size_t)n);
    orc_step(s, actions, outs, threads);
    for (int64_t i = 0; i < n; i++) {
        if (obs) memcpy(obs + 10 * i, outs[i].obs, 40);
        if (reward) reward[i] = outs[i].reward;
        if (terminated) terminated[i] = (uint8_t)outs[i].terminated;
        if (truncated) truncated[i] = (uint8_t)outs[i].truncated;
        if (final_obs && (outs[i].terminated || outs[i].truncated)) memcpy(final_obs + 10 * i, outs[i].final_obs, 40);
    }
    free(outs);
}

void orc_random_actions(const orc_sim *s, int64_t t, float *out) {
    for (int64_t i = 0; i < s->n; i++) {
        uint32_t r[4];
        draw4(s->cfg.seed, s->cfg.env_id_base + i, ST_ACTION, (uint32_t)t, (uint32_t)((uint64_t)t >> 32), r);
        out[2 * i + 0] = (float)(2.0 * u01(r[0]) - 1.0);
        out[2 * i + 1] = (float)(2.0 * u01(r[1]) - 1.0);
    }
}

void orc_stats(orc_sim *s, double out[ORC_NSTATS], int reset_after) {
    memcpy(out, s->stats, sizeof(s->stats));
    if (reset_after) memset(s->stats, 0, sizeof(s->stats));
}

// Fused state provider + OSC step for the DualUR5 (sm_100a): one THREAD per robot instance.
//
// What the reference pulls out of MuJoCo every timestep - mj_fullM (robot.py:68-72), the EE body
// Jacobians (device.py:115-133), qfrc_bias (osc.py:191), EE poses (device.py:93-95) and the F/T
// site frame (device.py:135-143) - is recomputed here from (q, dq) and a rigid-body description of
// the robot, inside the same kernel as the control law (osc.py:41-68, 120-210).  Only q, dq and the
// targets (~0.6 KB per instance) are read from HBM; M and J are never formed.
//
// Formulation.  Every spatial quantity is expressed in WORLD coordinates about the WORLD ORIGIN:
// motion vectors (omega, v_O), force vectors (n_O, f), hinge j with world axis a_j through c_j has
// the motion subspace s_j = (a_j, c_j x a_j).  Then
//     M[i][j]   = s_j . I^c_i s_i          (composite inertia of the subtree of i, j an ancestor)
//     J[r][j]   = s_j . e_r                (e_r = unit task force of row r at the EE: (p x x^, x^) or (x^, 0))
//     (M dq)_i  = s_i . sum_{b in subtree(i)} I_b v_b
//     bias_i    = s_i . sum_{b in subtree(i)} (I_b a_b + v_b x* I_b v_b)          (RNEA, a_world = -g)
// and eliminating joint i leaves-first from the augmented matrix [[M, J^T], [J, 0]] (the same
// elimination osc_tree.cuh performs on matrix entries) is the articulated-body recursion
//     f = I^A s_i,  d = s_i . f,   I^A <- I^A - f f^T / d,
//     e_r <- e_r - f (s_i . e_r) / d,   A[r][r'] += (s_i . e_r)(s_i . e_r') / d,
// on a fixed-size state (21 + 6 KD + KD(KD+1)/2 doubles per arm).  Each joint is the same code on
// different constants, so the kernel is a handful of compact loops (small instruction footprint,
// which is what limited osc_tree.cuh) with all state in registers, and every lane does useful work.
//
// Per instance: downward pass per arm (FK, velocities, accelerations, momenta; prefix sums give the
// subtree sums without a second sweep), the two gripper halves (3 joints each, leaves), upward
// articulated sweep over arm joints 6..1 with the arm's task rows, then the stand joint couples the
// arms; dense k x k LDL^T solve in registers; the joint-space assembly of osc.py:184-200 in its
// collapsed form u = c_j (M dq)_j + bias_j - (J^T w)_j.
// Instances whose task-space inverse needs the eigen-decomposition (pinv with rcond, osc.py:55)
// are pushed to a queue and finished by a warp-cooperative fix-up kernel (tiled::eigen_solve).
//
// The per-instance function is __host__ __device__ so that tests/host_fused can run exactly this
// code on the CPU against the oracle (test infrastructure only; the product never does).
#pragma once
#include <cmath>
#include "irlosc_device.cuh"
#include "osc_fused_types.h"

namespace irlosc {
namespace fused {

// Strided per-thread scratch: element i of this thread lives at base[i * stride]
// (device: shared memory, stride = threads per CTA -> conflict-free; host: stride 1).
struct Scratch {
    double *base;
    int stride;
    IRLOSC_HD double &operator()(int i) const { return base[(size_t)i * stride]; }
};
constexpr int kBodyScratch = 16;                   // s[6], r[3], Iw[6], pre
constexpr int kScratchDoubles = 6 * kBodyScratch;  // one arm's chain

struct Body {
    double R[9], o[3], v[6], a[6];
};

// ---------------------------------------------------------------- small helpers
IRLOSC_HD void mat3_vec(const double *A, const double *x, double *y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = fma(A[3 * i + 2], x[2], fma(A[3 * i + 1], x[1], A[3 * i] * x[0]));
}
IRLOSC_HD void mat3_mul(const double *A, const double *B, double *C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = fma(A[3 * i + 2], B[6 + j], fma(A[3 * i + 1], B[3 + j], A[3 * i] * B[j]));
}
IRLOSC_HD void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
IRLOSC_HD void cross3_add(const double *a, const double *b, double *c) {
    c[0] += a[1] * b[2] - a[2] * b[1];
    c[1] += a[2] * b[0] - a[0] * b[2];
    c[2] += a[0] * b[1] - a[1] * b[0];
}
IRLOSC_HD double dot6(const double *a, const double *b) {
    return fma(a[5], b[5], fma(a[4], b[4], fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])))));
}
IRLOSC_HD double rcp64(double d) {
#ifdef __CUDA_ARCH__
    return fast_rcp(d);
#else
    return 1.0 / d;
#endif
}
IRLOSC_HD void sincos64(double x, double *s, double *c) {
#ifdef __CUDA_ARCH__
    sincos(x, s, c);
#else
    *s = sin(x);
    *c = cos(x);
#endif
}

// ---------------------------------------------------------------- kinematics / momenta of one joint
// Child body of `p` through hinge `jm` at angle q, rate dq.  s: motion subspace, r: world COM,
// Iw: world inertia about the COM (xx yy zz xy xz yz).
IRLOSC_HD void joint_down(const KJoint &jm, const Body &p, double q, double dq, Body &c, double *s, double *r,
                          double *Iw) {
    double sn, cs;
    sincos64(q, &sn, &cs);
    double Rl[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Rl[e] = fma(cs, jm.Q0[e], fma(sn, jm.P1[e], jm.P2[e]));
    mat3_mul(p.R, Rl, c.R);
    double t[3];
    mat3_vec(p.R, jm.pos, t);
#pragma unroll
    for (int i = 0; i < 3; ++i) c.o[i] = p.o[i] + t[i];
    mat3_vec(p.R, jm.axp, s);                 // world hinge axis
    cross3(c.o, s, s + 3);                    // velocity of the world origin per unit rate
    // sdot = v x^ s (the parent's velocity; s x^ s = 0)
    double sd[6];
    cross3(p.v, s, sd);
    cross3(p.v, s + 3, sd + 3);
    cross3_add(p.v + 3, s, sd + 3);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        c.v[i] = fma(s[i], dq, p.v[i]);
        c.a[i] = fma(sd[i], dq, p.a[i]);
    }
    mat3_vec(c.R, jm.com, t);
#pragma unroll
    for (int i = 0; i < 3; ++i) r[i] = c.o[i] + t[i];
    // Iw = R Ic R^T
    const double ic[9] = {jm.ic[0], jm.ic[3], jm.ic[4], jm.ic[3], jm.ic[1], jm.ic[5], jm.ic[4], jm.ic[5], jm.ic[2]};
    double T[9];
    mat3_mul(c.R, ic, T);
    auto rt = [&](int i, int j) { return fma(T[3 * i + 2], c.R[3 * j + 2], fma(T[3 * i + 1], c.R[3 * j + 1], T[3 * i] * c.R[3 * j])); };
    Iw[0] = rt(0, 0); Iw[1] = rt(1, 1); Iw[2] = rt(2, 2);
    Iw[3] = rt(1, 0); Iw[4] = rt(2, 0); Iw[5] = rt(2, 1);
}

IRLOSC_HD void sym3_vec(const double *S, const double *x, double *y) {      // S: xx yy zz xy xz yz
    y[0] = fma(S[4], x[2], fma(S[3], x[1], S[0] * x[0]));
    y[1] = fma(S[5], x[2], fma(S[1], x[1], S[3] * x[0]));
    y[2] = fma(S[2], x[2], fma(S[5], x[1], S[4] * x[0]));
}

// h = I v (spatial momentum) and fb = I a + v x* h (bias wrench) of one rigid body.
IRLOSC_HD void body_wrench(double m, const double *r, const double *Iw, const double *v, const double *a,
                           double *h, double *fb) {
    double t[3];
    cross3(v, r, t);                                        // omega x r
    double pl[3] = {m * (v[3] + t[0]), m * (v[4] + t[1]), m * (v[5] + t[2])};
    sym3_vec(Iw, v, h);
    cross3_add(r, pl, h);
    h[3] = pl[0]; h[4] = pl[1]; h[5] = pl[2];
    cross3(a, r, t);                                        // alpha x r
    double pa[3] = {m * (a[3] + t[0]), m * (a[4] + t[1]), m * (a[5] + t[2])};
    sym3_vec(Iw, a, fb);
    cross3_add(r, pa, fb);
    cross3_add(v, h, fb);                                   // omega x L
    cross3_add(v + 3, pl, fb);                              // v_O x p
    cross3(v, pl, t);                                       // omega x p
    fb[3] = pa[0] + t[0]; fb[4] = pa[1] + t[1]; fb[5] = pa[2] + t[2];
}

// ---------------------------------------------------------------- symmetric 6x6 (articulated inertia)
IRLOSC_HD constexpr int sidx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

IRLOSC_HD void add_rigid(double *IA, double m, const double *r, const double *Iw) {
    const double mx = m * r[0], my = m * r[1], mz = m * r[2];
    IA[sidx(0, 0)] += Iw[0] + fma(my, r[1], mz * r[2]);
    IA[sidx(1, 1)] += Iw[1] + fma(mx, r[0], mz * r[2]);
    IA[sidx(2, 2)] += Iw[2] + fma(mx, r[0], my * r[1]);
    IA[sidx(1, 0)] += Iw[3] - mx * r[1];
    IA[sidx(2, 0)] += Iw[4] - mx * r[2];
    IA[sidx(2, 1)] += Iw[5] - my * r[2];
    IA[sidx(3, 1)] += mz;  IA[sidx(3, 2)] -= my;
    IA[sidx(4, 0)] -= mz;  IA[sidx(4, 2)] += mx;
    IA[sidx(5, 0)] += my;  IA[sidx(5, 1)] -= mx;
    IA[sidx(3, 3)] += m;   IA[sidx(4, 4)] += m;   IA[sidx(5, 5)] += m;
}

// Eliminate hinge s from the articulated inertia: f = IA s, inv = 1 / (s . f), IA -= f f^T inv.
// Returns false when the pivot is not positive (M not positive definite).
IRLOSC_HD bool joint_up(double *IA, const double *s, double *f, double *inv_out) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) acc = fma(IA[sidx(i, j)], s[j], acc);
        f[i] = acc;
    }
    const double d = dot6(s, f);
    const double inv = rcp64(d);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double fi = f[i] * inv;
#pragma unroll
        for (int j = 0; j <= i; ++j) IA[sidx(i, j)] = fma(-fi, f[j], IA[sidx(i, j)]);
    }
    *inv_out = inv;
    return d > 0.0;
}

// Unit task force of component `comp` (0..2 xyz, 3..5 abg) at world point p.
IRLOSC_HD void task_force(int comp, const double *p, double *e) {
#pragma unroll
    for (int i = 0; i < 6; ++i) e[i] = 0.0;
    if (comp >= 3) { e[comp - 3] = 1.0; return; }
    e[3 + comp] = 1.0;
    if (comp == 0) { e[1] = p[2]; e[2] = -p[1]; }
    else if (comp == 1) { e[0] = -p[2]; e[2] = p[0]; }
    else { e[0] = p[1]; e[1] = -p[0]; }
}

// Rotation matrix -> unit quaternion, w >= 0 (sign is irrelevant to osc.py:101-118).
IRLOSC_HD void mat_to_quat(const double *R, double *q) {
    const double m00 = R[0], m11 = R[4], m22 = R[8];
    const double tr = m00 + m11 + m22;
    if (tr > 0.0) {
        const double S = 2.0 * fast_sqrt(tr + 1.0), iS = rcp64(S);
        q[0] = 0.25 * S; q[1] = (R[7] - R[5]) * iS; q[2] = (R[2] - R[6]) * iS; q[3] = (R[3] - R[1]) * iS;
    } else if (m00 > m11 && m00 > m22) {
        const double S = 2.0 * fast_sqrt(1.0 + m00 - m11 - m22), iS = rcp64(S);
        q[0] = (R[7] - R[5]) * iS; q[1] = 0.25 * S; q[2] = (R[1] + R[3]) * iS; q[3] = (R[2] + R[6]) * iS;
    } else if (m11 > m22) {
        const double S = 2.0 * fast_sqrt(1.0 + m11 - m00 - m22), iS = rcp64(S);
        q[0] = (R[2] - R[6]) * iS; q[1] = (R[1] + R[3]) * iS; q[2] = 0.25 * S; q[3] = (R[5] + R[7]) * iS;
    } else {
        const double S = 2.0 * fast_sqrt(1.0 + m22 - m00 - m11), iS = rcp64(S);
        q[0] = (R[3] - R[1]) * iS; q[1] = (R[2] + R[6]) * iS; q[2] = (R[5] + R[7]) * iS; q[3] = 0.25 * S;
    }
    if (q[0] < 0.0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
}

// Coefficient of (M dq)_j in u_j: the velocity term of the device that owns joint j when it took
// the zero-target-velocity branch (osc.py:174, last device wins) plus the null-space term
// (osc.py:195-200, collapsed form).
IRLOSC_HD double coef_uv(const KParams &P, const int *vel_zero, int j) {
    double c = 0.0;
#pragma unroll
    for (int d = 0; d < IRLOSC_MAX_DEVICES; ++d)
        if (d < P.D && vel_zero[d] && ((P.dev[d].joint_mask >> j) & 1u)) c = -1.0 * P.dev[d].kv;
    if (P.has_nullspace) c -= P.nullspace_kv;
    return c;
}

// Optional taps for the host test harness (all pointers may be null).
struct Debug {
    double *A;      // K x K
    double *g;      // K
    double *uv;     // n: M dq
    double *bias;   // n
    double *dx;     // K
    double *J;      // K x n
};

// ---------------------------------------------------------------- one instance
// Returns true when the instance was queued for the eigen fix-up (outputs of the chain joints are
// then written by osc_fused_fixup / fixup_finish).
template <int KD, bool HAS_BASE>
IRLOSC_HD bool fused_instance(const KParams &P, const KModel &Mdl, const FRoles &R, const FIo &io, int64_t inst,
                              const Scratch &scr, double *hard_rec, const Debug *dbg) {
    constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    constexpr int N = kN;
    using RC = Rec<KD, HAS_BASE>;
    const int D = P.D;
    const double *q = io.q + inst * N;
    const double *dq = io.dq + inst * N;
    const double gb = P.use_g ? 1.0 : 0.0;
    int flags = 0;

    // which devices track a target velocity (osc.py:172-177): known from the inputs alone
    int vel_zero[IRLOSC_MAX_DEVICES];
#pragma unroll
    for (int d = 0; d < IRLOSC_MAX_DEVICES; ++d) {
        bool tracking = false;
        if (d < D && io.target_vel != nullptr) {
            tracking = true;
            for (int i = 0; i < 6; ++i) tracking = tracking && (io.target_vel[(inst * D + d) * 6 + i] != 0.0);
        }
        vel_zero[d] = tracking ? 0 : 1;
    }

    double u[N];                        // joint-space signal (osc.py:152-200)
    double As[K][K], j0[K], dxr[K];     // blocks of A = J M^-1 J^T before the stand joint couples them
    double jarm[2][6][KD], base_arm[2][6], jst[K];
    double ee_p[IRLOSC_MAX_DEVICES][3], ee_q[IRLOSC_MAX_DEVICES][4], ft_R[IRLOSC_MAX_DEVICES][9];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) As[i][j] = 0.0;

    // ------------------------------------------------------------ stand joint (world -> joint 0)
    Body W;
#pragma unroll
    for (int e = 0; e < 9; ++e) W.R[e] = (e % 4 == 0) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { W.o[i] = 0.0; W.v[i] = W.v[3 + i] = 0.0; W.a[i] = 0.0; W.a[3 + i] = -Mdl.gravity[i]; }
    Body S0;
    double s0[6], r0[3], Iw0[6], h0[6], fb0[6];
    joint_down(Mdl.stand, W, q[0], dq[0], S0, s0, r0, Iw0);
    body_wrench(Mdl.stand.mass, r0, Iw0, S0.v, S0.a, h0, fb0);
    double IA0[21], Htot[6], FBtot[6];
#pragma unroll
    for (int e = 0; e < 21; ++e) IA0[e] = 0.0;
    add_rigid(IA0, Mdl.stand.mass, r0, Iw0);
#pragma unroll
    for (int i = 0; i < 6; ++i) { Htot[i] = h0[i]; FBtot[i] = fb0[i]; }
    bool m_ok = true;

    if (HAS_BASE) {
        const int d = R.dev_base;
        const KFrame &F = Mdl.ee[d];
        double t[3], Re[9];
        mat3_vec(S0.R, F.pos, t);
        for (int i = 0; i < 3; ++i) ee_p[d][i] = S0.o[i] + t[i];
        mat3_mul(S0.R, F.R, Re);
        mat_to_quat(Re, ee_q[d]);
        for (int e = 0; e < 9; ++e) ft_R[d][e] = (e % 4 == 0) ? 1.0 : 0.0;
        double e6[6];
        task_force(P.row_comp[R.row_base], ee_p[d], e6);
        const double jb0 = dot6(s0, e6);
        j0[R.row_base] = jb0;
        jst[R.row_base] = jb0;
        dxr[R.row_base] = dot6(e6, S0.v);
    }

    // ------------------------------------------------------------ the two arms
#pragma unroll 1
    for (int arm = 0; arm < 2; ++arm) {
        const int jb = 1 + 12 * arm;
        const int dev = R.dev_arm[arm];
        const int row_a = R.row_arm[arm];
        // ---- downward pass over arm joints 1..6: FK, velocity, acceleration, momenta, prefix sums
        Body Bc = S0;
        double Hpre[6], FBpre[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { Hpre[i] = 0.0; FBpre[i] = 0.0; }
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {
            const KJoint &jm = Mdl.arm[arm][i];
            Body Bn;
            double s[6], r[3], Iw[6], h[6], fb[6];
            joint_down(jm, Bc, q[jb + i], dq[jb + i], Bn, s, r, Iw);
            body_wrench(jm.mass, r, Iw, Bn.v, Bn.a, h, fb);
            const double cu = coef_uv(P, vel_zero, jb + i);
            double x[6];
#pragma unroll
            for (int e = 0; e < 6; ++e) x[e] = fma(cu, Hpre[e], gb * FBpre[e]);
            const int o = i * kBodyScratch;
#pragma unroll
            for (int e = 0; e < 6; ++e) scr(o + e) = s[e];
#pragma unroll
            for (int e = 0; e < 3; ++e) scr(o + 6 + e) = r[e];
#pragma unroll
            for (int e = 0; e < 6; ++e) scr(o + 9 + e) = Iw[e];
            scr(o + 15) = dot6(s, x);
            if (dbg && dbg->uv) { dbg->uv[jb + i] = -dot6(s, Hpre); dbg->bias[jb + i] = -dot6(s, FBpre); }
#pragma unroll
            for (int e = 0; e < 6; ++e) { Hpre[e] += h[e]; FBpre[e] += fb[e]; }
            Bc = Bn;
        }
        // Bc = arm link 6; Hpre / FBpre = sums over the arm links (grippers are added below)
        // ---- EE pose, F/T frame, task forces (device.py:93-95,125-143)
        double E[KD][6];
        {
            const KFrame &F = Mdl.ee[dev];
            double t[3], Re[9];
            mat3_vec(Bc.R, F.pos, t);
            for (int i = 0; i < 3; ++i) ee_p[dev][i] = Bc.o[i] + t[i];
            mat3_mul(Bc.R, F.R, Re);
            mat_to_quat(Re, ee_q[dev]);
            const KFrame &T = Mdl.ft[dev];
            if (T.has) mat3_mul(Bc.R, T.R, ft_R[dev]);
            else
                for (int e = 0; e < 9; ++e) ft_R[dev][e] = 0.0;
#pragma unroll
            for (int cr = 0; cr < KD; ++cr) {
                task_force(P.row_comp[row_a + cr], ee_p[dev], E[cr]);
                dxr[row_a + cr] = dot6(E[cr], Bc.v);           // dx = J dq (osc.py:150)
                jst[row_a + cr] = dot6(E[cr], s0);             // J[r][stand]
            }
        }
        // ---- gripper halves: leaves g1 -> g0 and g2, eliminated into link 6's articulated inertia
        double IA[21];
#pragma unroll
        for (int e = 0; e < 21; ++e) IA[e] = 0.0;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int gj = jb + 6 + 3 * half;
            double IAg[21], f[6], inv;
            Body B0;
            double sa[6], ra[3], Iwa[6], ha[6], fba[6];
            joint_down(Mdl.grip[arm][half][0], Bc, q[gj], dq[gj], B0, sa, ra, Iwa);
            body_wrench(Mdl.grip[arm][half][0].mass, ra, Iwa, B0.v, B0.a, ha, fba);
            {
                Body B1;
                double sb[6], rb[3], Iwb[6], hb[6], fbb[6];
                joint_down(Mdl.grip[arm][half][1], B0, q[gj + 1], dq[gj + 1], B1, sb, rb, Iwb);
                body_wrench(Mdl.grip[arm][half][1].mass, rb, Iwb, B1.v, B1.a, hb, fbb);
                const double cu1 = coef_uv(P, vel_zero, gj + 1);
                const double uv1 = dot6(sb, hb), b1 = dot6(sb, fbb);
                u[gj + 1] = fma(cu1, uv1, gb * b1);
                if (dbg && dbg->uv) { dbg->uv[gj + 1] = uv1; dbg->bias[gj + 1] = b1; }
#pragma unroll
                for (int e = 0; e < 6; ++e) { ha[e] += hb[e]; fba[e] += fbb[e]; }     // subtree sums of g0
#pragma unroll
                for (int e = 0; e < 21; ++e) IAg[e] = 0.0;
                add_rigid(IAg, Mdl.grip[arm][half][1].mass, rb, Iwb);
                m_ok = joint_up(IAg, sb, f, &inv) && m_ok;
            }
            const double cu0 = coef_uv(P, vel_zero, gj);
            const double uv0 = dot6(sa, ha), b0 = dot6(sa, fba);
            u[gj] = fma(cu0, uv0, gb * b0);
            if (dbg && dbg->uv) { dbg->uv[gj] = uv0; dbg->bias[gj] = b0; }
#pragma unroll
            for (int e = 0; e < 6; ++e) { Hpre[e] += ha[e]; FBpre[e] += fba[e]; }
            add_rigid(IAg, Mdl.grip[arm][half][0].mass, ra, Iwa);
            m_ok = joint_up(IAg, sa, f, &inv) && m_ok;
#pragma unroll
            for (int e = 0; e < 21; ++e) IA[e] += IAg[e];
            {   // g2
                Body B2;
                double sc[6], rc[3], Iwc[6], hc[6], fbc[6];
                joint_down(Mdl.grip[arm][half][2], Bc, q[gj + 2], dq[gj + 2], B2, sc, rc, Iwc);
                body_wrench(Mdl.grip[arm][half][2].mass, rc, Iwc, B2.v, B2.a, hc, fbc);
                const double cu2 = coef_uv(P, vel_zero, gj + 2);
                const double uv2 = dot6(sc, hc), b2 = dot6(sc, fbc);
                u[gj + 2] = fma(cu2, uv2, gb * b2);
                if (dbg && dbg->uv) { dbg->uv[gj + 2] = uv2; dbg->bias[gj + 2] = b2; }
#pragma unroll
                for (int e = 0; e < 6; ++e) { Hpre[e] += hc[e]; FBpre[e] += fbc[e]; }
#pragma unroll
                for (int e = 0; e < 21; ++e) IAg[e] = 0.0;
                add_rigid(IAg, Mdl.grip[arm][half][2].mass, rc, Iwc);
                m_ok = joint_up(IAg, sc, f, &inv) && m_ok;
#pragma unroll
                for (int e = 0; e < 21; ++e) IA[e] += IAg[e];
            }
        }
        // Hpre / FBpre now hold the sums over the whole arm subtree
#pragma unroll
        for (int e = 0; e < 6; ++e) { Htot[e] += Hpre[e]; FBtot[e] += FBpre[e]; }
        // ---- upward articulated sweep over arm joints 6..1 with the arm's task rows
        double ak[KD * (KD + 1) / 2];
#pragma unroll
        for (int e = 0; e < KD * (KD + 1) / 2; ++e) ak[e] = 0.0;
#pragma unroll 1
        for (int i = 5; i >= 0; --i) {
            const int o = i * kBodyScratch;
            double s[6], r[3], Iw[6];
#pragma unroll
            for (int e = 0; e < 6; ++e) s[e] = scr(o + e);
#pragma unroll
            for (int e = 0; e < 3; ++e) r[e] = scr(o + 6 + e);
#pragma unroll
            for (int e = 0; e < 6; ++e) Iw[e] = scr(o + 9 + e);
            // (M dq)_j and bias_j through the subtree sums: s . (X_total - X_prefix)
            const double cu = coef_uv(P, vel_zero, jb + i);
            double x[6];
#pragma unroll
            for (int e = 0; e < 6; ++e) x[e] = fma(cu, Hpre[e], gb * FBpre[e]);
            base_arm[arm][i] = dot6(s, x) - scr(o + 15);
            if (dbg && dbg->uv) { dbg->uv[jb + i] += dot6(s, Hpre); dbg->bias[jb + i] += dot6(s, FBpre); }
#pragma unroll
            for (int cr = 0; cr < KD; ++cr) {
                double e0[6];
                task_force(P.row_comp[row_a + cr], ee_p[dev], e0);
                jarm[arm][i][cr] = dot6(s, e0);                 // original J[r][joint] for J^T w
            }
            add_rigid(IA, Mdl.arm[arm][i].mass, r, Iw);
            double f[6], inv;
            m_ok = joint_up(IA, s, f, &inv) && m_ok;
            double jk[KD];
#pragma unroll
            for (int cr = 0; cr < KD; ++cr) jk[cr] = dot6(s, E[cr]);
#pragma unroll
            for (int cr = 0; cr < KD; ++cr) {
                const double tc = jk[cr] * inv;
#pragma unroll
                for (int c2 = cr; c2 < KD; ++c2) ak[c2 * (c2 + 1) / 2 + cr] = fma(jk[c2], tc, ak[c2 * (c2 + 1) / 2 + cr]);
#pragma unroll
                for (int e = 0; e < 6; ++e) E[cr][e] = fma(-tc, f[e], E[cr][e]);
            }
        }
        // ---- what the stand joint needs from this arm
#pragma unroll
        for (int e = 0; e < 21; ++e) IA0[e] += IA[e];
#pragma unroll
        for (int cr = 0; cr < KD; ++cr) {
            j0[row_a + cr] = dot6(s0, E[cr]);
#pragma unroll
            for (int c2 = 0; c2 <= cr; ++c2) {
                As[row_a + cr][row_a + c2] = ak[cr * (cr + 1) / 2 + c2];
                As[row_a + c2][row_a + cr] = ak[cr * (cr + 1) / 2 + c2];
            }
        }
    }

    // ------------------------------------------------------------ stand joint couples the arms
    double f0[6], inv0;
    m_ok = joint_up(IA0, s0, f0, &inv0) && m_ok;
    const double cu_st = coef_uv(P, vel_zero, 0);
    const double uv_st = dot6(s0, Htot), b_st = dot6(s0, FBtot);
    const double base_st = fma(cu_st, uv_st, gb * b_st);
    if (dbg && dbg->uv) { dbg->uv[0] = uv_st; dbg->bias[0] = b_st; }

    // ------------------------------------------------------------ per-device task signal (osc.py:156-181)
    double g[K];
#pragma unroll 1
    for (int d = 0; d < D; ++d) {
        const KDevice &dv = P.dev[d];
        double mv[2] = {dv.max_vel[0], dv.max_vel[1]};
        if (io.max_vel) { mv[0] = io.max_vel[(inst * D + d) * 2]; mv[1] = io.max_vel[(inst * D + d) * 2 + 1]; }
        double tv[6], u6[6], txyz[3], tquat[4];
        if (io.target_vel)
            for (int i = 0; i < 6; ++i) tv[i] = io.target_vel[(inst * D + d) * 6 + i];
        for (int i = 0; i < 3; ++i) txyz[i] = io.target_xyz[(inst * D + d) * 3 + i];
        for (int i = 0; i < 4; ++i) tquat[i] = io.target_quat[(inst * D + d) * 4 + i];
        bool oob = false;
        const bool tracking = device_task_signal(dv, ee_p[d], ee_q[d], txyz, tquat, io.target_vel ? tv : nullptr, mv,
                                                 dxr, K, u6, &oob);
        double ft[6] = {0, 0, 0, 0, 0, 0};
        if (P.admittance) {
            double raw[6];
            for (int i = 0; i < 6; ++i) raw[i] = io.ft_raw[(inst * D + d) * 6 + i];
            rotate_wrench(ft_R[d], raw, ft);
        }
        int r = dv.row0;
        const double kvn = P.has_nullspace ? P.nullspace_kv : 0.0;
        for (int i = 0; i < 6; ++i)
            if (dv.dof[i]) {
                const double v = P.admittance ? u6[i] + ft[i] : u6[i];
                g[r] = v - kvn * dxr[r];
                ++r;
            }
        flags |= (tracking ? IRLOSC_ST_VEL_BRANCH : 0) | (oob ? IRLOSC_ST_DX_RANGE : 0);
        if (io.ee_xyz)
            for (int i = 0; i < 3; ++i) io.ee_xyz[(inst * D + d) * 3 + i] = ee_p[d][i];
        if (io.ee_quat)
            for (int i = 0; i < 4; ++i) io.ee_quat[(inst * D + d) * 4 + i] = ee_q[d][i];
    }

    // ------------------------------------------------------------ A = blocks + j0 j0^T / d0, LDL^T solve
    double a[K * (K + 1) / 2];
    double fro2 = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const double ji = j0[i] * inv0;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const double v = fma(ji, j0[j], As[i][j]);
            a[i * (i + 1) / 2 + j] = v;
            fro2 = fma(v, (i == j) ? v : 2.0 * v, fro2);
        }
    }
    if (dbg && dbg->A) {
        for (int i = 0; i < K; ++i)
            for (int j = 0; j <= i; ++j) { dbg->A[i * K + j] = a[i * (i + 1) / 2 + j]; dbg->A[j * K + i] = a[i * (i + 1) / 2 + j]; }
        for (int i = 0; i < K; ++i) { dbg->g[i] = g[i]; dbg->dx[i] = dxr[i]; }
    }
    if (!m_ok) flags |= IRLOSC_ST_M_NOT_PD;
    const bool poison = (flags & (IRLOSC_ST_M_NOT_PD | IRLOSC_ST_DX_RANGE)) != 0;

    // in-place LDL^T: a[i][p] becomes l_ip, diagonal keeps d_p; dinv[p] = 1 / d_p
    double dinv[K], w[K];
    double detinv = 1.0;
    bool a_bad = false;
#pragma unroll
    for (int p = 0; p < K; ++p) {
        const double inv = rcp64(a[p * (p + 1) / 2 + p]);
        dinv[p] = inv;
        detinv *= inv;
        a_bad = a_bad || !(inv > 0.0);
#pragma unroll
        for (int i = p + 1; i < K; ++i) {
            const double aip = a[i * (i + 1) / 2 + p];
            const double l = aip * inv;
#pragma unroll
            for (int j = p + 1; j <= i; ++j) a[i * (i + 1) / 2 + j] = fma(-l, a[j * (j + 1) / 2 + p], a[i * (i + 1) / 2 + j]);
        }
#pragma unroll
        for (int i = p + 1; i < K; ++i) a[i * (i + 1) / 2 + p] *= inv;
    }
    // w = A^-1 g
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double z = g[i];
#pragma unroll
        for (int j = 0; j < i; ++j) z = fma(-a[i * (i + 1) / 2 + j], w[j], z);
        w[i] = z;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) w[i] *= dinv[i];
#pragma unroll
    for (int i = K - 1; i >= 0; --i) {
        double z = w[i];
#pragma unroll
        for (int j = i + 1; j < K; ++j) z = fma(-a[j * (j + 1) / 2 + i], w[j], z);
        w[i] = z;
    }
    // osc.py:52-55: |det| >= 1e-4 -> inverse.  Otherwise pinv(rcond = 1e-5), which equals the inverse
    // unless an eigenvalue is <= 1e-5 lambda_max.  lambda_max <= ||A||_F and 1 / lambda_min <= tr(A^-1),
    // so ||A||_F tr(A^-1) < 1e5 certifies that nothing is cut.  tr(A^-1) = sum_p dinv_p |row p of L^-1|^2.
    const bool small_det = !(fabs(detinv) <= 1.0 / kDetThreshold);
    bool certified = true;
    if (small_det && !a_bad) {
        // X = L^-1 in place (unit lower triangular), column by column
        double tr_inv = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            double x[K];
            x[j] = 1.0;
            double acc = dinv[j];
#pragma unroll
            for (int i = j + 1; i < K; ++i) {
                double z = -a[i * (i + 1) / 2 + j];
#pragma unroll
                for (int m = j + 1; m < i; ++m) z = fma(-a[i * (i + 1) / 2 + m], x[m], z);
                x[i] = z;
                acc = fma(z * z, dinv[i], acc);
            }
            tr_inv += acc;
        }
        certified = (fro2 * tr_inv * tr_inv < (1.0 / kPinvRcond) * (1.0 / kPinvRcond));
    }
    const bool hard = !poison && (a_bad || (small_det && !certified));
    if (small_det && !a_bad) flags |= IRLOSC_ST_PINV;

    // ------------------------------------------------------------ joint-space assembly + packing
    if (hard && hard_rec != nullptr) {
        // full A again (the factorisation overwrote it)
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const double ji = j0[i] * inv0;
#pragma unroll
            for (int j = 0; j < K; ++j) hard_rec[RC::A + i * K + j] = fma(ji, j0[j], As[i][j]);
            hard_rec[RC::G + i] = g[i];
            hard_rec[RC::JST + i] = jst[i];
        }
        hard_rec[RC::BASE] = base_st;
        for (int am = 0; am < 2; ++am)
            for (int i = 0; i < 6; ++i) {
                hard_rec[RC::BASE + 1 + 6 * am + i] = base_arm[am][i];
                for (int cr = 0; cr < KD; ++cr) hard_rec[RC::JARM + (am * 6 + i) * KD + cr] = jarm[am][i][cr];
            }
        hard_rec[RC::ABAD] = a_bad ? 1.0 : 0.0;
    }
    {
        double jt = 0.0;
#pragma unroll
        for (int r = 0; r < K; ++r) jt = fma(jst[r], w[r], jt);
        u[0] = base_st - jt;
    }
#pragma unroll
    for (int am = 0; am < 2; ++am)
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double jt = 0.0;
#pragma unroll
            for (int cr = 0; cr < KD; ++cr) jt = fma(jarm[am][i][cr], w[R.row_arm[am] + cr], jt);
            u[1 + 12 * am + i] = base_arm[am][i] - jt;
        }
    if (poison) {
        const double qnan = nan("");
        for (int j = 0; j < N; ++j) u[j] = qnan;
    }
    if (io.u_all)
        for (int j = 0; j < N; ++j) io.u_all[inst * N + j] = u[j];
    for (int d = 0; d < D; ++d) {
        const KDevice &dv = P.dev[d];
        for (int c = 0; c < dv.n_ctrl; ++c) io.ctrl[inst * P.n_ctrl + dv.ctrl0 + c] = u[dv.actuator[c]];
    }
    if (io.status) io.status[inst] = (uint8_t)flags;
    if (dbg && dbg->J) {
        for (int e = 0; e < K * N; ++e) dbg->J[e] = 0.0;
        for (int r = 0; r < K; ++r) dbg->J[r * N] = jst[r];
        for (int am = 0; am < 2; ++am)
            for (int i = 0; i < 6; ++i)
                for (int cr = 0; cr < KD; ++cr) dbg->J[(R.row_arm[am] + cr) * N + 1 + 12 * am + i] = jarm[am][i][cr];
    }
    return hard && hard_rec != nullptr;
}

// Finish one queued instance given w = pinv(A) g: the 13 joints that have Jacobian columns.
template <int KD, bool HAS_BASE>
IRLOSC_HD void fixup_finish(const KParams &P, const FRoles &R, const FIo &io, int64_t inst, const double *rec,
                            const double *w, int j_lo, int j_step) {
    using RC = Rec<KD, HAS_BASE>;
    constexpr int K = RC::K;
    // joint slots: 0 = stand, 1 + 6 a + i = arm a joint i
    for (int sl = j_lo; sl < 13; sl += j_step) {
        double jt = 0.0;
        int joint = 0;
        if (sl == 0) {
            for (int r = 0; r < K; ++r) jt = fma(rec[RC::JST + r], w[r], jt);
        } else {
            const int am = (sl - 1) / 6, i = (sl - 1) % 6;
            joint = 1 + 12 * am + i;
            for (int cr = 0; cr < KD; ++cr) jt = fma(rec[RC::JARM + (am * 6 + i) * KD + cr], w[R.row_arm[am] + cr], jt);
        }
        const double uj = rec[RC::BASE + sl] - jt;
        if (io.u_all) io.u_all[inst * kN + joint] = uj;
        for (int d = 0; d < P.D; ++d) {
            const KDevice &dv = P.dev[d];
            for (int c = 0; c < dv.n_ctrl; ++c)
                if (dv.actuator[c] == joint) io.ctrl[inst * P.n_ctrl + dv.ctrl0 + c] = uj;
        }
    }
}

#if defined(__CUDACC__) && !defined(IRLOSC_FUSED_NO_KERNELS)
// ---------------------------------------------------------------- kernels
template <int KD, bool HAS_BASE, int NT>
__global__ void __launch_bounds__(NT, 1)
osc_step_fused(const __grid_constant__ KParams P, const __grid_constant__ KModel Mdl, const FIo io, const int64_t B,
               const FRoles R, const HardQueue hq) {
    extern __shared__ __align__(16) double fused_smem[];
    const Scratch scr{fused_smem + threadIdx.x, NT};
    for (int64_t inst = (int64_t)blockIdx.x * NT + threadIdx.x; inst < B; inst += (int64_t)gridDim.x * NT) {
        // queue slot claimed only when needed: build the record in place
        // (record memory is per instance slot = inst while capacity >= B, so no atomics for the body)
        double *rec = hq.rec ? hq.rec + (size_t)inst * hq.rec_doubles : nullptr;
        const bool hard = fused_instance<KD, HAS_BASE>(P, Mdl, R, io, inst, scr, rec, nullptr);
        if (hard) {
            const int slot = atomicAdd(hq.count, 1);
            hq.inst[slot] = inst;
        }
    }
}

// One warp per queued instance: eigen-decomposition of A (tiled::eigen_solve), w = pinv(A) g, then the
// joints with Jacobian columns are rewritten.
template <int KD, bool HAS_BASE>
__global__ void __launch_bounds__(128, 1)
osc_fused_fixup(const __grid_constant__ KParams P, const FIo io, const FRoles R, const HardQueue hq) {
    using RC = Rec<KD, HAS_BASE>;
    constexpr int K = RC::K;
    struct WarpSmem {
        double As[K][K + 1], Vs[K][K + 1];
        double g[K], w[K], cbuf[32], sbuf[32];
        int flags;
    };
    __shared__ WarpSmem sm[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSmem &S = sm[warp];
    const int n_hard = *hq.count;
    for (int slot = blockIdx.x * 4 + warp; slot < n_hard; slot += gridDim.x * 4) {
        const int64_t inst = hq.inst[slot];
        const double *rec = hq.rec + (size_t)inst * hq.rec_doubles;
        for (int e = lane; e < K * K; e += 32) S.As[e / K][e % K] = rec[RC::A + e];
        if (lane < K) S.g[lane] = rec[RC::G + lane];
        if (lane == 0) S.flags = 0;
        __syncwarp();
        const bool a_bad = rec[RC::ABAD] != 0.0;
        tiled::eigen_solve<K>(S.As, S.Vs, S.g, S.w, S.cbuf, S.sbuf, !a_bad, lane, &S.flags);
        __syncwarp();
        fixup_finish<KD, HAS_BASE>(P, R, io, inst, rec, S.w, lane, 32);
        if (io.status && lane == 0) io.status[inst] = (uint8_t)(io.status[inst] | S.flags);
        __syncwarp();
    }
}
#endif  // __CUDACC__

}  // namespace fused
}  // namespace irlosc

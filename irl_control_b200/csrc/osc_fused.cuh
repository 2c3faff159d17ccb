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
// The few instances whose task-space solve the thread cannot decide (osc_tail.cuh) are finished by their
// warp with the cooperative eigen-solver right after the tile, inside this kernel.
//
// The per-instance function is __host__ __device__ so that tests/host_fused can run exactly this
// code on the CPU against the oracle (test infrastructure only; the product never does).
#pragma once
#include <cmath>
#include "irlosc_device.cuh"
#include "osc_fused_types.h"
#include "osc_tail.cuh"
#include "osc_sequence.cuh"

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
    // Iw = R Ic R^T; an isotropic tensor (every UR5 link of scenes/dual_ur5.xml) is rotation invariant
    if (jm.ic[3] == 0.0 && jm.ic[4] == 0.0 && jm.ic[5] == 0.0 && jm.ic[0] == jm.ic[1] && jm.ic[1] == jm.ic[2]) {
        Iw[0] = Iw[1] = Iw[2] = jm.ic[0];
        Iw[3] = Iw[4] = Iw[5] = 0.0;
        return;
    }
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
    const bool lin = comp < 3;                      // no dynamic indexing: e stays in registers
    const int c = lin ? comp : comp - 3;
    const double ux = (c == 0) ? 1.0 : 0.0, uy = (c == 1) ? 1.0 : 0.0, uz = (c == 2) ? 1.0 : 0.0;
    e[0] = lin ? p[1] * uz - p[2] * uy : ux;
    e[1] = lin ? p[2] * ux - p[0] * uz : uy;
    e[2] = lin ? p[0] * uy - p[1] * ux : uz;
    e[3] = lin ? ux : 0.0;
    e[4] = lin ? uy : 0.0;
    e[5] = lin ? uz : 0.0;
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

// ---------------------------------------------------------------- one instance
// osc.py:159-168,179-181 for one device whose EE pose is known: the task signal before the
// velocity-tracking term, plus the rotated F/T wrench when admittance is on -> gpre[row].
IRLOSC_HD void device_signal_early(const KParams &P, const FIo &io, int64_t inst, int d, const double *ee_p,
                                   const double *ee_q, const double *Rft, bool has_ft, double mv0_override,
                                   double *gpre) {
    const int D = P.D;
    const KDevice &dv = P.dev[d];
    double mv[2] = {dv.max_vel[0], dv.max_vel[1]};
    if (io.max_vel) { mv[0] = io.max_vel[(inst * D + d) * 2]; mv[1] = io.max_vel[(inst * D + d) * 2 + 1]; }
    if (mv0_override >= 0.0) mv[0] = mv0_override;            // action-sequence mode: active_arm.max_vel[0]
    double u6[6], txyz[3], tquat[4];
#pragma unroll
    for (int i = 0; i < 3; ++i) txyz[i] = io.target_xyz[(inst * D + d) * 3 + i];
#pragma unroll
    for (int i = 0; i < 4; ++i) tquat[i] = io.target_quat[(inst * D + d) * 4 + i];
    bool oob = false;
    device_task_signal(dv, ee_p, ee_q, txyz, tquat, nullptr, mv, nullptr, 0, u6, &oob);
    if (P.admittance) {
        double ft[6] = {0, 0, 0, 0, 0, 0};
        if (has_ft) {
            double raw[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) raw[i] = io.ft_raw[(inst * D + d) * 6 + i];
            rotate_wrench(Rft, raw, ft);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) u6[i] += ft[i];
    }
    int r = dv.row0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (dv.dof[i]) gpre[r++] = u6[i];
    if (io.ee_xyz)
        for (int i = 0; i < 3; ++i) io.ee_xyz[(inst * D + d) * 3 + i] = ee_p[i];
    if (io.ee_quat)
        for (int i = 0; i < 4; ++i) io.ee_quat[(inst * D + d) * 4 + i] = ee_q[i];
}

// One arm of one instance: downward pass, EE pose and task signal, gripper halves, upward articulated sweep with the
// arm's task rows.  Adds the arm's subtree momenta / bias wrenches to Hsum / FBsum and its articulated inertia at the
// stand to IAsum; leaves the arm's block of A (ak_out), the stand column of the reduced and original task rows (j0a,
// jsta), dx (dxa), the original Jacobian entries (jarm_a[joint][row]) and the joint terms (base_a) of ITS rows / joints,
// and the device's task signal in g (task-row order).  Returns false when a pivot is not positive.
template <int KD, bool SEQ>
IRLOSC_HD bool fused_arm(const KParams &P, const KModel &Mdl, const FRoles &R, const FIo &io, int64_t inst, const Scratch &scr,
                         int arm, unsigned vel_zero, double gb, const Body &S0, const double *s0, const double *q, const double *dq,
                         double *g, double *ak_out, double *j0a, double *jsta, double *dxa, double (*jarm_a)[KD], double *base_a,
                         double *Hsum, double *FBsum, double *IAsum, SeqResult &seq, const KSeq *Q, const Debug *dbg) {
    constexpr int N = kN;
    constexpr int KT = KD * (KD + 1) / 2;
    const int D = P.D;
    bool ok = true;
    const int jb = 1 + 12 * arm;
    const int dev = R.dev_arm[arm];
    const int row_a = R.row_arm[arm];
    // ---- downward pass over arm joints 1..6: FK, velocity, acceleration, momenta, prefix sums
    Body Bc = S0;
    double Hpre[6], FBpre[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { Hpre[i] = 0.0; FBpre[i] = 0.0; }
    const double cu_arm = coef_uv(P, R, vel_zero, jb);     // shared by the six arm joints when R.uniform_owner
    // joint angles / rates are fetched one joint ahead so that their latency hides behind a joint's work
    double q_nx = q[jb], dq_nx = dq[jb];
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
        const KJoint &jm = Mdl.arm[arm][i];
        Body Bn;
        double s[6], r[3], Iw[6], h[6], fb[6];
        const double q_i = q_nx, dq_i = dq_nx;
        q_nx = q[jb + i + 1];                  // i = 5: first gripper joint, still inside the row
        dq_nx = dq[jb + i + 1];
        joint_down(jm, Bc, q_i, dq_i, Bn, s, r, Iw);
        body_wrench(jm.mass, r, Iw, Bn.v, Bn.a, h, fb);
        const double cu = R.uniform_owner ? cu_arm : coef_uv(P, R, vel_zero, jb + i);
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
    // ---- EE pose, F/T frame, task signal and task forces (device.py:93-95,125-143; osc.py:156-181)
    double ee_p[3];
    {
        const KFrame &F = Mdl.ee[dev];
        double t[3], Re[9], Rf[9];
        mat3_vec(Bc.R, F.pos, t);
#pragma unroll
        for (int i = 0; i < 3; ++i) ee_p[i] = Bc.o[i] + t[i];
        mat3_mul(Bc.R, F.R, Re);
        double ee_q[4], mv0 = -1.0;
        mat_to_quat(Re, ee_q);
        if (SEQ && Q->mode == 1) {
            // gain_test.py:138-158: target = wps[idx]; after generate, |EE_XYZ - target| < threshold -> next
            // waypoint (wrapping).  The comparison uses the state this step is computed from.
            int idx = io.wp_idx[inst * D + dev];
            const double *wp = io.wps + (((size_t)inst * D + dev) * Q->W + idx) * 3;
            double *tx = io.seq_tgt_xyz + (inst * D + dev) * 3;
            double e2 = 0.0;
            for (int i = 0; i < 3; ++i) { tx[i] = wp[i]; const double dlt = ee_p[i] - wp[i]; e2 = fma(dlt, dlt, e2); }
            if (sqrt(e2) < Q->threshold) idx = (idx < Q->n_wp[dev] - 1) ? idx + 1 : 0;
            io.wp_idx[inst * D + dev] = idx;
        } else if (SEQ) {
            if (dev == Q->active_dev) {
                seq = seq_advance(*Q, P.dev[dev], io, inst, D, ee_p, ee_q);
                mv0 = io.seq_mv0[inst];
            } else if (seq.entered_wp) {
                for (int i = 0; i < 3; ++i) io.seq_tgt_xyz[(inst * D + dev) * 3 + i] = ee_p[i];
                for (int i = 0; i < 4; ++i) io.seq_tgt_quat[(inst * D + dev) * 4 + i] = Q->passive_quat[i];
            }
        }
        const KFrame &T = Mdl.ft[dev];
        if (P.admittance && T.has) mat3_mul(Bc.R, T.R, Rf);
        device_signal_early(P, io, inst, dev, ee_p, ee_q, Rf, T.has != 0, mv0, g);
#pragma unroll
        for (int cr = 0; cr < KD; ++cr) {
            double e0[6];
            task_force(P.row_comp[row_a + cr], ee_p, e0);
            dxa[cr] = dot6(e0, Bc.v);              // dx = J dq (osc.py:150)
            jsta[cr] = dot6(e0, s0);                // J[r][stand]
        }
    }
    // ---- gripper halves: leaves g1 -> g0 and g2, eliminated into link 6's articulated inertia
    double IA[21];
#pragma unroll
    for (int e = 0; e < 21; ++e) IA[e] = 0.0;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        const int gj = jb + 6 + 3 * half;
        const double cu_grip = coef_uv(P, R, vel_zero, gj);  // shared by the half's three joints when R.uniform_owner
        const double qg0 = q[gj], qg1 = q[gj + 1], qg2 = q[gj + 2];
        const double dqg0 = dq[gj], dqg1 = dq[gj + 1], dqg2 = dq[gj + 2];
        double IAg[21], f[6], inv;
        Body B0;
        double sa[6], ra[3], Iwa[6], ha[6], fba[6];
        joint_down(Mdl.grip[arm][half][0], Bc, qg0, dqg0, B0, sa, ra, Iwa);
        body_wrench(Mdl.grip[arm][half][0].mass, ra, Iwa, B0.v, B0.a, ha, fba);
        {
            Body B1;
            double sb[6], rb[3], Iwb[6], hb[6], fbb[6];
            joint_down(Mdl.grip[arm][half][1], B0, qg1, dqg1, B1, sb, rb, Iwb);
            body_wrench(Mdl.grip[arm][half][1].mass, rb, Iwb, B1.v, B1.a, hb, fbb);
            const double cu1 = R.uniform_owner ? cu_grip : coef_uv(P, R, vel_zero, gj + 1);
            const double uv1 = dot6(sb, hb), b1 = dot6(sb, fbb);
            put_joint(R, io.u_all ? io.u_all + inst * N : nullptr, io.ctrl + inst * P.n_ctrl, gj + 1, fma(cu1, uv1, gb * b1));
            if (dbg && dbg->uv) { dbg->uv[gj + 1] = uv1; dbg->bias[gj + 1] = b1; }
#pragma unroll
            for (int e = 0; e < 6; ++e) { ha[e] += hb[e]; fba[e] += fbb[e]; }     // subtree sums of g0
#pragma unroll
            for (int e = 0; e < 21; ++e) IAg[e] = 0.0;
            add_rigid(IAg, Mdl.grip[arm][half][1].mass, rb, Iwb);
            ok = joint_up(IAg, sb, f, &inv) && ok;
        }
        const double cu0 = cu_grip;
        const double uv0 = dot6(sa, ha), b0 = dot6(sa, fba);
        put_joint(R, io.u_all ? io.u_all + inst * N : nullptr, io.ctrl + inst * P.n_ctrl, gj, fma(cu0, uv0, gb * b0));
        if (dbg && dbg->uv) { dbg->uv[gj] = uv0; dbg->bias[gj] = b0; }
#pragma unroll
        for (int e = 0; e < 6; ++e) { Hpre[e] += ha[e]; FBpre[e] += fba[e]; }
        add_rigid(IAg, Mdl.grip[arm][half][0].mass, ra, Iwa);
        ok = joint_up(IAg, sa, f, &inv) && ok;
#pragma unroll
        for (int e = 0; e < 21; ++e) IA[e] += IAg[e];
        {   // g2
            Body B2;
            double sc[6], rc[3], Iwc[6], hc[6], fbc[6];
            joint_down(Mdl.grip[arm][half][2], Bc, qg2, dqg2, B2, sc, rc, Iwc);
            body_wrench(Mdl.grip[arm][half][2].mass, rc, Iwc, B2.v, B2.a, hc, fbc);
            const double cu2 = R.uniform_owner ? cu_grip : coef_uv(P, R, vel_zero, gj + 2);
            const double uv2 = dot6(sc, hc), b2 = dot6(sc, fbc);
            put_joint(R, io.u_all ? io.u_all + inst * N : nullptr, io.ctrl + inst * P.n_ctrl, gj + 2, fma(cu2, uv2, gb * b2));
            if (dbg && dbg->uv) { dbg->uv[gj + 2] = uv2; dbg->bias[gj + 2] = b2; }
#pragma unroll
            for (int e = 0; e < 6; ++e) { Hpre[e] += hc[e]; FBpre[e] += fbc[e]; }
#pragma unroll
            for (int e = 0; e < 21; ++e) IAg[e] = 0.0;
            add_rigid(IAg, Mdl.grip[arm][half][2].mass, rc, Iwc);
            ok = joint_up(IAg, sc, f, &inv) && ok;
#pragma unroll
            for (int e = 0; e < 21; ++e) IA[e] += IAg[e];
        }
    }
    // Hpre / FBpre now hold the sums over the whole arm subtree
#pragma unroll
    for (int e = 0; e < 6; ++e) { Hsum[e] += Hpre[e]; FBsum[e] += FBpre[e]; }
    // ---- upward articulated sweep over arm joints 6..1 with the arm's task rows
    double ak[KT], E[KD][6];
#pragma unroll
    for (int e = 0; e < KT; ++e) ak[e] = 0.0;
#pragma unroll
    for (int cr = 0; cr < KD; ++cr) task_force(P.row_comp[row_a + cr], ee_p, E[cr]);
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
        const double cu = R.uniform_owner ? cu_arm : coef_uv(P, R, vel_zero, jb + i);
        double x[6];
#pragma unroll
        for (int e = 0; e < 6; ++e) x[e] = fma(cu, Hpre[e], gb * FBpre[e]);
        base_a[i] = dot6(s, x) - scr(o + 15);
        if (dbg && dbg->uv) { dbg->uv[jb + i] += dot6(s, Hpre); dbg->bias[jb + i] += dot6(s, FBpre); }
        {   // original J[r][joint] for J^T w: [jacp; jacr] column = [a x (p - c); a] = [s_ang x p + s_lin; s_ang]
            double jp[3];
            cross3(s, ee_p, jp);
#pragma unroll
            for (int e = 0; e < 3; ++e) jp[e] += s[3 + e];
#pragma unroll
            for (int cr = 0; cr < KD; ++cr) {
                const int comp = P.row_comp[row_a + cr];
                jarm_a[i][cr] = comp == 0 ? jp[0] : comp == 1 ? jp[1] : comp == 2 ? jp[2] : comp == 3 ? s[0] : comp == 4 ? s[1] : s[2];
            }
        }
        add_rigid(IA, Mdl.arm[arm][i].mass, r, Iw);
        double f[6], inv;
        ok = joint_up(IA, s, f, &inv) && ok;
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
    for (int e = 0; e < 21; ++e) IAsum[e] += IA[e];
#pragma unroll
    for (int cr = 0; cr < KD; ++cr) j0a[cr] = dot6(s0, E[cr]);
#pragma unroll
    for (int e = 0; e < KT; ++e) ak_out[e] = ak[e];
    return ok;
}

// Returns true when the task-space solve must be finished by the warp (state_warp_finish on T rewrites the
// chain joints).
template <int KD, bool HAS_BASE, bool SEQ = false>
IRLOSC_HD bool fused_instance(const KParams &P, const KModel &Mdl, const FRoles &R, const FIo &io, int64_t inst,
                              const Scratch &scr, TailState<KD, HAS_BASE> &T, const Debug *dbg, const KSeq *Q = nullptr) {
    constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    constexpr int N = kN;
    constexpr int KT = KD * (KD + 1) / 2;
    const int D = P.D;
    const double *q = io.q + inst * N;
    const double *dq = io.dq + inst * N;
    const double gb = P.use_g ? 1.0 : 0.0;
    int flags = 0;

    // which devices track a target velocity (osc.py:172-177): known from the inputs alone
    unsigned vel_zero = 0;
#pragma unroll
    for (int d = 0; d < IRLOSC_MAX_DEVICES; ++d) {
        bool tracking = false;
        if (d < D && io.target_vel != nullptr) {
            tracking = true;
            for (int i = 0; i < 6; ++i) tracking = tracking && (io.target_vel[(inst * D + d) * 6 + i] != 0.0);
        }
        if (!tracking) vel_zero |= 1u << d;
    }

    // what survives the arm loop lives in T (dynamic indices -> local memory; kept as small as possible)
    double (&akA)[2][KT] = T.akA;
    double (&j0)[K] = T.j0, (&jst)[K] = T.jst, (&dxr)[K] = T.dxr, (&g)[K] = T.g;
    double (&jarm)[2][6][KD] = T.jarm, (&base_arm)[2][6] = T.base_arm;

    // ------------------------------------------------------------ stand joint (world -> joint 0)
    Body S0;
    double s0[6];
    double IA0[21], Htot[6], FBtot[6];
    {
        Body W;
#pragma unroll
        for (int e = 0; e < 9; ++e) W.R[e] = (e % 4 == 0) ? 1.0 : 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { W.o[i] = 0.0; W.v[i] = W.v[3 + i] = 0.0; W.a[i] = 0.0; W.a[3 + i] = -Mdl.gravity[i]; }
        double r0[3], Iw0[6];
        joint_down(Mdl.stand, W, q[0], dq[0], S0, s0, r0, Iw0);
        body_wrench(Mdl.stand.mass, r0, Iw0, S0.v, S0.a, Htot, FBtot);
#pragma unroll
        for (int e = 0; e < 21; ++e) IA0[e] = 0.0;
        add_rigid(IA0, Mdl.stand.mass, r0, Iw0);
    }
    bool m_ok = true;

    if (HAS_BASE) {
        const int d = R.dev_base;
        const KFrame &F = Mdl.ee[d];
        double t[3], Re[9], p[3], eq[4];
        mat3_vec(S0.R, F.pos, t);
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = S0.o[i] + t[i];
        mat3_mul(S0.R, F.R, Re);
        mat_to_quat(Re, eq);
        device_signal_early(P, io, inst, d, p, eq, Re, false, -1.0, g);
        double e6[6];
        task_force(P.row_comp[R.row_base], p, e6);
        const double jb0 = dot6(s0, e6);
        j0[R.row_base] = jb0;
        jst[R.row_base] = jb0;
        dxr[R.row_base] = dot6(e6, S0.v);
    }

    // ------------------------------------------------------------ the two arms
    // (action-sequence mode: the active arm first - its pose decides whether a waypoint starts, and the
    //  passive arm then latches its own xyz as target, insertion_task.py:211-213)
    SeqResult seq{false, 0.0};
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
        int arm = it;
        if (SEQ && Q->mode == 0) {
            const int act_arm = (R.dev_arm[0] == Q->active_dev) ? 0 : 1;
            arm = it == 0 ? act_arm : 1 - act_arm;
        }
        m_ok = fused_arm<KD, SEQ>(P, Mdl, R, io, inst, scr, arm, vel_zero, gb, S0, s0, q, dq, g, akA[arm], j0 + R.row_arm[arm],
                                  jst + R.row_arm[arm], dxr + R.row_arm[arm], jarm[arm], base_arm[arm], Htot, FBtot, IA0, seq, Q, dbg) && m_ok;
    }

    // ------------------------------------------------------------ stand joint couples the arms
    double f0[6], inv0;
    m_ok = joint_up(IA0, s0, f0, &inv0) && m_ok;
    const double cu_st = coef_uv(P, R, vel_zero, 0);
    const double uv_st = dot6(s0, Htot), b_st = dot6(s0, FBtot);
    const double base_st = fma(cu_st, uv_st, gb * b_st);
    if (dbg && dbg->uv) { dbg->uv[0] = uv_st; dbg->bias[0] = b_st; }

    T.base_st = base_st;
    T.inv0 = inv0;
    T.u_all_row = io.u_all ? io.u_all + inst * N : nullptr;
    T.ctrl_row = io.ctrl + inst * P.n_ctrl;
    T.status = io.status ? io.status + inst : nullptr;
    const bool hard = state_tail<KD, HAS_BASE>(P, R, io.target_vel ? io.target_vel + inst * D * 6 : nullptr, vel_zero, flags,
                                               m_ok, T, dbg);
    // send_forces: `if gripper_force: sim.data.ctrl[gripper_idx] = gripper_force` (insertion_task.py:160-161); the
    // gripper joint is not a chain joint, so a later warp finish does not touch this slot
    if (SEQ && seq.gripper_force != 0.0 && Q->gripper_slot >= 0) io.ctrl[inst * P.n_ctrl + Q->gripper_slot] = seq.gripper_force;
    return hard;
}

#if defined(__CUDACC__) && !defined(IRLOSC_FUSED_NO_KERNELS)
// ---------------------------------------------------------------- kernels
// SMEM = true: the per-thread chain scratch lives in shared memory (strided, conflict-free);
// SMEM = false: in local memory, leaving the whole 256 KB of the SM to L1, which then also catches
// the register spills and the small dynamically indexed arrays.
template <int KD, bool HAS_BASE, int NT, bool SMEM, bool SEQ = false>
__global__ void __launch_bounds__(NT, 1)
osc_step_fused(const __grid_constant__ KParams P, const __grid_constant__ KModel Mdl, const __grid_constant__ FIo io,
               const int64_t B, const __grid_constant__ FRoles R, const __grid_constant__ KSeq Q) {
    extern __shared__ __align__(16) double fused_smem[];
    double chain[SMEM ? 1 : kScratchDoubles];
    const Scratch scr = SMEM ? Scratch{fused_smem + threadIdx.x, NT} : Scratch{chain, 1};
    // scratch of the warp-cooperative finish (osc_tail.cuh), one per warp behind the chain scratch
    WarpFix<KD, HAS_BASE> &wfix = reinterpret_cast<WarpFix<KD, HAS_BASE> *>(fused_smem + (SMEM ? kScratchDoubles * NT : 0))[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count: the finish needs every lane of the warp
    for (int64_t base = (int64_t)blockIdx.x * NT + (threadIdx.x & ~31); base < B; base += (int64_t)gridDim.x * NT) {
        const int64_t inst = base + lane;
        TailState<KD, HAS_BASE> T;
        bool hard = false;
        if (inst < B) hard = fused_instance<KD, HAS_BASE, SEQ>(P, Mdl, R, io, inst, scr, T, nullptr, SEQ ? &Q : nullptr);
        __syncwarp();
        state_warp_finish<KD, HAS_BASE>(wfix, R, T, hard, lane);
    }
}

#endif  // __CUDACC__

}  // namespace fused
}  // namespace irlosc

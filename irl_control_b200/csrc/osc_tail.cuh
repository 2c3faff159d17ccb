// Second half of a DualUR5 OSC step, shared by the fused step (osc_fused.cuh: state computed from
// q, dq) and the streaming step (osc_stream.cuh: state read from HBM), one thread per instance.
//
// Input: what the leaves-first elimination of the joints leaves behind - the two arms' diagonal
// blocks of A = J M^-1 J^T, the stand column of the reduced and of the original J, dx = J dq, the
// task signal, the original Jacobian entries of the arm joints and c_j (M dq)_j + bias_j per joint.
// Here: velocity-tracking term (osc.py:175-177), A = blocks + j0 j0^T / d0, dense LDL^T solve in
// registers, the inverse-vs-pinv decision of osc.py:52-55, joint-space assembly (osc.py:184-200,
// collapsed form of DESIGN.md 4.1) and packing (osc.py:203-208).  Instances whose task-space inverse
// needs the eigen-decomposition are written to a record and finished by osc_tail_fixup.
#pragma once
#include <cmath>
#include "irlosc_device.cuh"
#include "osc_fused_types.h"

namespace irlosc {
namespace fused {

IRLOSC_HD double rcp64(double d) {
#ifdef __CUDA_ARCH__
    return fast_rcp(d);
#else
    return 1.0 / d;
#endif
}

// Coefficient of (M dq)_j in u_j: the velocity term of the device that owns joint j when it took
// the zero-target-velocity branch (osc.py:174, last device wins) plus the null-space term
// (osc.py:195-200, collapsed form).
IRLOSC_HD double coef_uv(const KParams &P, unsigned vel_zero, int j) {
    double c = 0.0;
#pragma unroll
    for (int d = 0; d < IRLOSC_MAX_DEVICES; ++d)
        if (d < P.D && ((vel_zero >> d) & 1u) && ((P.dev[d].joint_mask >> j) & 1u)) c = -1.0 * P.dev[d].kv;
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

// Outputs of one joint: u_all and, when the joint is an actuated one, its packed ctrl slot
// (osc.py:203-208 through the joint -> slot table built in irlosc_set_model).
IRLOSC_HD void put_joint(const FRoles &R, double *u_all_row, double *ctrl_row, int j, double uj) {
    if (u_all_row) u_all_row[j] = uj;
    const int slot = R.joint_slot[j];
    if (slot >= 0) ctrl_row[slot] = uj;
}


// In-thread resolution of the pinv branch (osc.py:52-55) for an instance the cheap certificate left open.
// numpy's pinv(A, rcond) keeps an eigenvalue iff it is > rcond * lambda_max.  With rigorous two-sided bounds
//   L <= lambda_max <= U    (repeated squaring of A / tr A: tr(B^(2^m))^(1/2^m) is within K^(1/2^m) of lambda_max)
// three outcomes are certain without an eigen-decomposition:
//   1  1 / tr(A^-1) > rcond U            : nothing is cut, pinv = inverse (w from the LDL^T stands);
//   2  the smallest Rayleigh quotient (3 inverse iterations on the LDL^T at hand, so lambda_min <= rho) is
//      <= rcond L, and the deflated matrix A' = A + U x x^T has 1 / tr(A'^-1) > rcond U.  Eigenvalues
//      interlace (lambda_2(A) >= lambda_1(A')), so exactly one eigenvalue is cut:
//      w = A'^-1 (g - x (x . g));
//   0  anything else (borderline within the bounds, two or more small eigenvalues, a failed pivot): left to
//      the eigen fix-up kernel.
// Runs on local arrays with rolled loops: it is the rare path of one lane, code size matters more than speed.
template <int K>
IRLOSC_HD int resolve_pinv(double (*Af)[K], double (*Lf)[K], const double *dinv, const double *g, double tr_inv, double *w) {
    constexpr int MS = 6;                        // squarings at most: bounds within K^(1/64) (4 % for K = 13)
    double tr = 0.0;
    for (int i = 0; i < K; ++i) tr += Af[i][i];
    if (!(tr > 0.0) || !(tr_inv > 0.0)) return 0;
    auto solve = [&](double (*Lm)[K], const double *dv, double *x) {      // x <- (L D L^T)^-1 x
        for (int i = 0; i < K; ++i) {
            double z = x[i];
            for (int j = 0; j < i; ++j) z = fma(-Lm[i][j], x[j], z);
            x[i] = z;
        }
        for (int i = 0; i < K; ++i) x[i] *= dv[i];
        for (int i = K - 1; i >= 0; --i) {
            double z = x[i];
            for (int j = i + 1; j < K; ++j) z = fma(-Lm[j][i], x[j], z);
            x[i] = z;
        }
    };
    // ---- smallest eigenpair: inverse iteration with the factors at hand.  It contracts by
    // lambda_min / lambda_2 per step, which the classification does not bound away from 1: iterate until the
    // extrapolated error of x is at rounding level (or do without outcome 2).
    double x[K], xp[K];
    for (int i = 0; i < K; ++i) x[i] = xp[i] = 1.0 + 0.1 * i;        // a component along every axis
    bool converged = false;
    double d_prev = HUGE_VAL;
    for (int it = 0; it < 64 && !converged; ++it) {
        solve(Lf, dinv, x);
        double n2 = 0.0, dot = 0.0;
        for (int i = 0; i < K; ++i) { n2 = fma(x[i], x[i], n2); dot = fma(x[i], xp[i], dot); }
        if (!(n2 > 0.0) || !(n2 < HUGE_VAL)) break;
        const double in = (dot < 0.0 ? -1.0 : 1.0) / sqrt(n2);
        double d2 = 0.0;
        for (int i = 0; i < K; ++i) {
            x[i] *= in;
            const double dlt = x[i] - xp[i];
            d2 = fma(dlt, dlt, d2);
            xp[i] = x[i];
        }
        const double d = sqrt(d2);
        if (it >= 2) {
            const double r = d / d_prev;                              // observed contraction
            converged = (d == 0.0) || (r < 0.97 && d * r / (1.0 - r) < 1e-13);
        }
        d_prev = d;
    }
    double rho = HUGE_VAL;                                            // Rayleigh quotient: lambda_min <= rho
    if (converged) {
        rho = 0.0;
        for (int i = 0; i < K; ++i) {
            double acc = 0.0;
            for (int j = 0; j < K; ++j) acc = fma(Af[i][j], x[j], acc);
            rho = fma(x[i], acc, rho);
        }
    }
    // ---- bounds on lambda_max, refined one squaring at a time until the instance is classified
    double Bm[K][K], Cm[K][K], dv2[K], sm[MS];
    {
        const double itr = 1.0 / tr;
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) Bm[i][j] = Af[i][j] * itr;
    }
    bool deflated = false;
    double tr2 = 0.0;
    for (int m = 0; m <= MS; ++m) {
        if (m > 0) {                             // B <- B^2 / tr(B^2): traces stay 1
            double t = 0.0;
            for (int i = 0; i < K; ++i)
                for (int j = 0; j <= i; ++j) {
                    double acc = 0.0;
                    for (int l = 0; l < K; ++l) acc = fma(Bm[i][l], Bm[l][j], acc);
                    Cm[i][j] = acc;
                    if (i == j) t += acc;
                }
            if (!(t > 0.0)) return 0;
            sm[m - 1] = t;
            const double it = 1.0 / t;
            for (int i = 0; i < K; ++i)
                for (int j = 0; j <= i; ++j) { Bm[i][j] = Cm[i][j] * it; Bm[j][i] = Bm[i][j]; }
        }
        double up = 1.0, lo = 1.0 / K;           // lambda_max(B_m) in [1 / K, 1] (PSD, trace 1)
        for (int j = m - 1; j >= 0; --j) { up = sqrt(sm[j] * up); lo = sqrt(sm[j] * lo); }
        const double c_hi = kPinvRcond * tr * up * (1.0 + 1e-12), c_lo = kPinvRcond * tr * lo * (1.0 - 1e-12);
        if (1.0 > c_hi * tr_inv) return 1;       // every eigenvalue is above the cutoff
        if (rho <= c_lo) {                       // the smallest one is below it
            if (!deflated) {                     // factor A' = A + tr(A) x x^T once
                deflated = true;
                for (int i = 0; i < K; ++i)
                    for (int j = 0; j <= i; ++j) Cm[i][j] = fma(tr * x[i], x[j], Af[i][j]);
                for (int p = 0; p < K; ++p) {
                    const double d = Cm[p][p];
                    if (!(d > 0.0)) return 0;
                    const double inv = 1.0 / d;
                    dv2[p] = inv;
                    for (int i = p + 1; i < K; ++i) {
                        const double l = Cm[i][p] * inv;
                        for (int j = p + 1; j <= i; ++j) Cm[i][j] = fma(-l, Cm[j][p], Cm[i][j]);
                    }
                    for (int i = p + 1; i < K; ++i) Cm[i][p] *= inv;
                }
                for (int j = 0; j < K; ++j) {    // tr(A'^-1) = sum_p dv2_p |row p of L^-1|^2
                    double col[K];
                    col[j] = 1.0;
                    double acc = dv2[j];
                    for (int i = j + 1; i < K; ++i) {
                        double z = -Cm[i][j];
                        for (int l = j + 1; l < i; ++l) z = fma(-Cm[i][l], col[l], z);
                        col[i] = z;
                        acc = fma(z * z, dv2[i], acc);
                    }
                    tr2 += acc;
                }
                // (Cm is needed again below: the squaring must not overwrite it)
                for (int i = 0; i < K; ++i)
                    for (int j = 0; j <= i; ++j) Af[j][i] = Cm[i][j];        // park the factors in Af's upper part
            }
            if (1.0 > c_hi * tr2) {              // and every other one is above it: cut exactly that one
                for (int i = 0; i < K; ++i)
                    for (int j = 0; j < i; ++j) Cm[i][j] = Af[j][i];
                double xg = 0.0;
                for (int i = 0; i < K; ++i) xg = fma(x[i], g[i], xg);
                for (int i = 0; i < K; ++i) w[i] = fma(-xg, x[i], g[i]);
                solve(Cm, dv2, w);
                return 2;
            }
        }
    }
    return 0;
}

template <int KD, bool HAS_BASE>
IRLOSC_HD bool osc_tail(const KParams &P, const FRoles &R, const double *target_vel, unsigned vel_zero, int flags,
                        bool m_ok, const double (*akA)[KD * (KD + 1) / 2], const double *j0, const double *jst,
                        const double *dxr, double *g, const double (*jarm)[6][KD], const double (*base_arm)[6],
                        double base_st, double inv0, double *u_all_row, double *ctrl_row, uint8_t *status,
                        double *hard_rec, const Debug *dbg) {
    constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    constexpr int N = kN;
    constexpr int KT = KD * (KD + 1) / 2;
    using RC = Rec<KD, HAS_BASE>;
    const int D = P.D;
    // ------------------------------------------------------------ velocity-tracking term (osc.py:175-177), g
    if (target_vel != nullptr) {
#pragma unroll 1
        for (int d = 0; d < D; ++d) {
            if ((vel_zero >> d) & 1u) continue;
            const KDevice &dv = P.dev[d];
            flags |= IRLOSC_ST_VEL_BRANCH;
            int r = 0;
            for (int i = 0; i < 6; ++i)
                if (dv.dof[i]) {
                    const int src = dv.dx_idx[r];
                    if (src >= K) flags |= IRLOSC_ST_DX_RANGE;     // IndexError in the reference (N3)
                    else g[dv.row0 + r] += dv.kv * (dxr[src] - target_vel[d * 6 + i]) * dv.damp[i];
                    ++r;
                }
        }
    }
    {
        const double kvn = P.has_nullspace ? P.nullspace_kv : 0.0;
#pragma unroll
        for (int r = 0; r < K; ++r) g[r] = fma(-kvn, dxr[r], g[r]);
    }

    // ------------------------------------------------------------ A = blocks + j0 j0^T / d0, LDL^T solve
    // Solved in CANONICAL row order (arm 0 rows, arm 1 rows, base row) so that every index below is
    // static; det, traces and the solution are invariant under the symmetric permutation.
    int perm[K];
#pragma unroll
    for (int i = 0; i < K; ++i) perm[i] = (i < KD) ? R.row_arm[0] + i : (i < 2 * KD) ? R.row_arm[1] + i - KD : R.row_base;
    double j0c[K], gc[K], b0[KT], b1[KT];
#pragma unroll
    for (int i = 0; i < K; ++i) { j0c[i] = j0[perm[i]]; gc[i] = g[perm[i]]; }
#pragma unroll
    for (int e = 0; e < KT; ++e) { b0[e] = akA[0][e]; b1[e] = akA[1][e]; }
    auto blk = [&](int i, int j) -> double {         // i >= j, canonical
        if (i < KD) return b0[i * (i + 1) / 2 + j];
        if (i < 2 * KD && j >= KD) return b1[(i - KD) * (i - KD + 1) / 2 + (j - KD)];
        return 0.0;
    };
    double a[K * (K + 1) / 2];
    double fro2 = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const double ji = j0c[i] * inv0;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const double v = fma(ji, j0c[j], blk(i, j));
            a[i * (i + 1) / 2 + j] = v;
            fro2 = fma(v, (i == j) ? v : 2.0 * v, fro2);
        }
    }
    if (dbg && dbg->A) {
        for (int i = 0; i < K; ++i)
            for (int j = 0; j <= i; ++j) {
                dbg->A[perm[i] * K + perm[j]] = a[i * (i + 1) / 2 + j];
                dbg->A[perm[j] * K + perm[i]] = a[i * (i + 1) / 2 + j];
            }
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
        double z = gc[i];
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
    double tr_inv = 0.0;
    if (small_det && !a_bad) {
        // X = L^-1 (unit lower triangular), column by column
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
    bool hard = !poison && (a_bad || (small_det && !certified));
    if (small_det && !a_bad) flags |= IRLOSC_ST_PINV;
    if (hard && !a_bad) {                           // rare: copy to local arrays, resolve without an eigen-decomposition
        double Af[K][K], Lf[K][K], wl[K], gl[K], dl[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const double ji = j0c[i] * inv0;
            gl[i] = gc[i];
            dl[i] = dinv[i];
            wl[i] = w[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double v = fma(ji, j0c[j], blk(i, j));
                Af[i][j] = v;
                Af[j][i] = v;
                Lf[i][j] = a[i * (i + 1) / 2 + j];
            }
        }
        const int how = resolve_pinv<K>(Af, Lf, dl, gl, tr_inv, wl);
        if (how != 0) hard = false;
        if (how == 2) {
            flags |= IRLOSC_ST_DEFLATED;
#pragma unroll
            for (int i = 0; i < K; ++i) w[i] = wl[i];
        }
    }

    // ------------------------------------------------------------ joint-space assembly + packing
    if (hard && hard_rec != nullptr) {              // record in canonical row order (w comes back canonical)
#pragma unroll 1
        for (int i = 0; i < K; ++i) {
            const double ji = j0[perm[i]] * inv0;
            for (int j = 0; j < K; ++j) {
                const int hi = i > j ? i : j, lo = i > j ? j : i;
                double bv = 0.0;
                if (hi < KD) bv = akA[0][hi * (hi + 1) / 2 + lo];
                else if (hi < 2 * KD && lo >= KD) bv = akA[1][(hi - KD) * (hi - KD + 1) / 2 + (lo - KD)];
                hard_rec[RC::A + i * K + j] = fma(ji, j0[perm[j]], bv);
            }
            hard_rec[RC::G + i] = g[perm[i]];
            hard_rec[RC::JST + i] = jst[perm[i]];
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
        for (int r = 0; r < K; ++r) jt = fma(jst[perm[r]], w[r], jt);
        put_joint(R, u_all_row, ctrl_row, 0, base_st - jt);
    }
#pragma unroll
    for (int am = 0; am < 2; ++am) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double jt = 0.0;
#pragma unroll
            for (int cr = 0; cr < KD; ++cr) jt = fma(jarm[am][i][cr], w[am * KD + cr], jt);
            put_joint(R, u_all_row, ctrl_row, 1 + 12 * am + i, base_arm[am][i] - jt);
        }
    }
    if (poison) {                        // M not positive definite / IndexError of the reference: NaN outputs
        const double qnan = nan("");
        if (u_all_row)
            for (int j = 0; j < N; ++j) u_all_row[j] = qnan;
        for (int c = 0; c < P.n_ctrl; ++c) ctrl_row[c] = qnan;
    }
    if (status) *status = (uint8_t)flags;
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
IRLOSC_HD void fixup_finish(const FRoles &R, double *u_all_row, double *ctrl_row, const double *rec, const double *w,
                            int j_lo, int j_step) {
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
            for (int cr = 0; cr < KD; ++cr) jt = fma(rec[RC::JARM + (am * 6 + i) * KD + cr], w[am * KD + cr], jt);
        }
        put_joint(R, u_all_row, ctrl_row, joint, rec[RC::BASE + sl] - jt);
    }
}

// Where the fix-up kernel rewrites the outputs of a queued instance.
struct TailOut {
    double *u_all, *ctrl;
    uint8_t *status;
    int32_t n_gather;
    int64_t gather_offset;
    double *ctrl_gather[IRLOSC_MAX_PEERS];
    double *ctrl_mc;
};

}  // namespace fused
}  // namespace irlosc

#if defined(__CUDACC__) && !defined(IRLOSC_FUSED_NO_KERNELS)
#include "osc_eigen.cuh"

namespace irlosc {
namespace fused {

// One warp per queued instance: eigen-decomposition of A (tiled::eigen_solve), w = pinv(A) g, then the
// joints with Jacobian columns are rewritten (and re-sent to the gather targets, if any).
template <int KD, bool HAS_BASE>
__global__ void __launch_bounds__(128, 1)
osc_tail_fixup(const __grid_constant__ KParams P, const __grid_constant__ TailOut out, const __grid_constant__ FRoles R,
               const __grid_constant__ HardQueue hq) {
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
        double *ctrl_row = out.ctrl + inst * P.n_ctrl;
        fixup_finish<KD, HAS_BASE>(R, out.u_all ? out.u_all + inst * kN : nullptr, ctrl_row, rec, S.w, lane, 32);
        if (out.status && lane == 0) out.status[inst] = (uint8_t)(out.status[inst] | S.flags);
        __syncwarp();
        if ((out.n_gather > 0 || out.ctrl_mc) && lane < P.n_ctrl) {
            const double v = ctrl_row[lane];
            const int64_t at = (out.gather_offset + inst) * P.n_ctrl + lane;
            if (out.ctrl_mc) multimem_st(out.ctrl_mc + at, v);
            else
                for (int gi = 0; gi < out.n_gather; ++gi) out.ctrl_gather[gi][at] = v;
        }
        __syncwarp();
    }
}

}  // namespace fused
}  // namespace irlosc
#endif

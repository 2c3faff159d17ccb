// Second half of a DualUR5 OSC step, shared by the thread-per-instance kernels: the lane kernel
// (osc_lane.cuh: state read from batch-interleaved tiles), the streaming kernel (osc_stream.cuh: state
// gathered from plain arrays) and the fused kernel (osc_fused.cuh: state computed from q, dq).
//
// Input: what the leaves-first elimination of the joints leaves behind - the two arms' diagonal
// blocks D0, D1 of J M^-1 J^T with the stand joint held back, the stand column v of the reduced J, the
// stand pivot d0, dx = J dq, the task signal, the original Jacobian entries of the chain joints and
// c_j (M dq)_j + bias_j per joint.  The task-space matrix of osc.py:47 is then
//
//     A = J M^-1 J^T = blockdiag(D0, D1, [0]) + v v^T / d0        (canonical row order: arm 0, arm 1, base)
//
// and everything osc.py:41-68 does with it is done on that structure, in the thread that owns the instance:
//   * two KD x KD LDL^T factorisations instead of one K x K (KD = 3 or 6, K = 2 KD [+ 1]); A^-1 r by
//     Sherman-Morrison (no base row) or by the bordered form (base row: its diagonal block is zero, so the
//     stand multiplier follows from the base equation alone); det A by the determinant lemma;
//   * the branch of osc.py:52-55: |det A| >= 1e-4 -> inverse.  Otherwise pinv(rcond = 1e-5), which keeps an
//     eigenvalue iff it is > 1e-5 lambda_max.  How many eigenvalues lie below a shift follows EXACTLY from the
//     inertia of the bordered matrix [[-d0, v^T], [v, D - sigma I]] (Sylvester): the signs of the pivots of
//     D0 - sigma I and D1 - sigma I plus the sign of -d0 - v^T (D - sigma I)^-1 v.  lambda_max is bracketed by
//     max diag(A) <= lambda_max <= ||A||_F and the bracket is bisected (again by inertia counts) until the
//     number m of eigenvalues under the cut-off is the same at both ends;
//   * m = 0: pinv = inverse.  m = 1, 2: the cut eigenvectors X come from inverse (subspace) iteration with the
//     block solver, and pinv(A) g = P A^-1 P g with P = I - X X^T (no eigen-decomposition of the rest);
//   * every solution is verified by its residual against the structured A.
// Instances this cannot decide (m > 2, slow convergence, failed pivots or residual; ~0.1 % of a k = 13 batch,
// none in the other configurations) are finished by their own warp right after the tile with the
// cooperative Jacobi eigen-solver (osc_eigen.cuh) - inside the same kernel, no second launch, no queue.
//
// Then: joint-space assembly (osc.py:184-200, collapsed form of DESIGN.md 4.1) and packing (osc.py:203-208).
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
IRLOSC_HD double sqrt64(double d) {
#ifdef __CUDA_ARCH__
    return fast_sqrt(d);
#else
    return sqrt(d);
#endif
}

// Coefficient of (M dq)_j in u_j: the velocity term of the device that owns joint j when it took
// the zero-target-velocity branch (osc.py:174, last device wins) plus the null-space term
// (osc.py:195-200, collapsed form).
IRLOSC_HD double coef_uv(const KParams &P, const FRoles &, unsigned vel_zero, int j) {
    // (a per-joint owner table in FRoles was measured SLOWER than these four mask tests: +10 % on the direct-load lane
    //  kernel, profiles/r02_summary.md)
    double c = 0.0;
#pragma unroll
    for (int d = 0; d < IRLOSC_MAX_DEVICES; ++d)
        if (d < P.D && ((vel_zero >> d) & 1u) && ((P.dev[d].joint_mask >> j) & 1u)) c = -1.0 * P.dev[d].kv;
    if (P.has_nullspace) c -= P.nullspace_kv;
    return c;
}

// The same for `count` consecutive joints of one group (an arm's six arm joints, a gripper half's three joints).
template <int COUNT>
IRLOSC_HD void coef_uv_group(const KParams &P, const FRoles &R, unsigned vel_zero, int j0, double *c) {
    if (R.uniform_owner) {
        const double c0 = coef_uv(P, R, vel_zero, j0);
#pragma unroll
        for (int i = 0; i < COUNT; ++i) c[i] = c0;
    } else {
#pragma unroll
        for (int i = 0; i < COUNT; ++i) c[i] = coef_uv(P, R, vel_zero, j0 + i);
    }
}

// Optional taps for the host test harness (all pointers may be null).
struct Debug {
    double *A;      // K x K
    double *g;      // K
    double *uv;     // n: M dq
    double *bias;   // n
    double *dx;     // K
    double *J;      // K x n
    int *how;       // per instance: TailHow bits (which path resolved the task-space solve)
};

// Outputs of one joint: u_all and, when the joint is an actuated one, its packed ctrl slot
// (osc.py:203-208 through the joint -> slot table built in irlosc_set_model).
IRLOSC_HD void put_joint(const FRoles &R, double *u_all_row, double *ctrl_row, int j, double uj) {
    if (u_all_row) u_all_row[j] = uj;
    const int slot = R.joint_slot[j];
    if (slot >= 0) ctrl_row[slot] = uj;
}

// ------------------------------------------------------------------------------------------------------
// Small symmetric blocks, packed lower triangle, every index static after unrolling.
IRLOSC_HD constexpr int ltri(int i, int j) { return i * (i + 1) / 2 + j; }

// (D - sigma I) = L Delta L^T without pivoting.  f: strict lower part = L, diagonal = 1 / Delta.
// Returns the number of negative pivots; *lost is set when a pivot lost all its digits to cancellation
// (its sign, and everything after it, means nothing then).
template <int KD>
IRLOSC_HD int blk_factor(const double *Dp, double sigma, double *f, bool *lost) {
    constexpr int KT = KD * (KD + 1) / 2;
#pragma unroll
    for (int e = 0; e < KT; ++e) f[e] = Dp[e];
    int nneg = 0;
#pragma unroll
    for (int p = 0; p < KD; ++p) {
        const double d = f[ltri(p, p)] - sigma;
        nneg += (d < 0.0) ? 1 : 0;
        *lost = *lost || !(fabs(d) > 1e-13 * (fabs(Dp[ltri(p, p)]) + fabs(sigma)));
        const double inv = rcp64(d);
#pragma unroll
        for (int i = p + 1; i < KD; ++i) {
            const double l = f[ltri(i, p)] * inv;
#pragma unroll
            for (int j = p + 1; j < i; ++j) f[ltri(i, j)] = fma(-l, f[ltri(j, p)], f[ltri(i, j)]);
            f[ltri(i, i)] = fma(-l, f[ltri(i, p)], f[ltri(i, i)]);
        }
#pragma unroll
        for (int i = p + 1; i < KD; ++i) f[ltri(i, p)] *= inv;
        f[ltri(p, p)] = inv;
    }
    return nneg;
}
// y <- L^-1 y;  returns sum y_i^2 / Delta_i  (= b^T (L Delta L^T)^-1 b for the original b)
template <int KD>
IRLOSC_HD double blk_forward(const double *f, double *y) {
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < KD; ++i) {
        double z = y[i];
#pragma unroll
        for (int j = 0; j < i; ++j) z = fma(-f[ltri(i, j)], y[j], z);
        y[i] = z;
        q = fma(z * z, f[ltri(i, i)], q);
    }
    return q;
}
// y <- (L Delta L^T)^-1 y
template <int KD>
IRLOSC_HD void blk_solve(const double *f, double *y) {
    blk_forward<KD>(f, y);
#pragma unroll
    for (int i = 0; i < KD; ++i) y[i] *= f[ltri(i, i)];
#pragma unroll
    for (int i = KD - 1; i >= 0; --i) {
        double z = y[i];
#pragma unroll
        for (int j = i + 1; j < KD; ++j) z = fma(-f[ltri(j, i)], y[j], z);
        y[i] = z;
    }
}

// A = blockdiag(D0, D1, [0]) + v v^T / d0 in canonical row order.
template <int KD, bool HAS_BASE>
struct TaskSys {
    static constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    static constexpr int KT = KD * (KD + 1) / 2;
    double D[2][KT];          // the arms' blocks
    double f[2][KT];          // their factors
    double v[K];
    double u[2 * KD];         // D^-1 v (arm rows)
    double d0, inv0;
    double piv;               // no base: 1 / (d0 + v^T D^-1 v);  base: 1 / v_b
};

template <int KD, bool HB>
IRLOSC_HD double sys_entry(const TaskSys<KD, HB> &S, int i, int j) {       // i >= j, static after unrolling
    double b = 0.0;
    if (i < KD) b = S.D[0][ltri(i, j)];
    else if (i < 2 * KD && j >= KD) b = S.D[1][ltri(i - KD, j - KD)];
    return fma(S.v[i] * S.inv0, S.v[j], b);
}

// y = A x
template <int KD, bool HB>
IRLOSC_HD void sys_matvec(const TaskSys<KD, HB> &S, const double *x, double *y) {
    constexpr int K = TaskSys<KD, HB>::K;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) s = fma(S.v[i], x[i], s);
    s *= S.inv0;
#pragma unroll
    for (int i = 0; i < K; ++i) y[i] = S.v[i] * s;
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int i = 0; i < KD; ++i)
#pragma unroll
            for (int j = 0; j < KD; ++j)
                y[b * KD + i] = fma(S.D[b][i >= j ? ltri(i, j) : ltri(j, i)], x[b * KD + j], y[b * KD + i]);
}

// w = A^-1 r (r is not modified)
template <int KD, bool HB>
IRLOSC_HD void sys_solve(const TaskSys<KD, HB> &S, const double *r, double *w) {
#pragma unroll
    for (int i = 0; i < 2 * KD; ++i) w[i] = r[i];
    blk_solve<KD>(S.f[0], w);
    blk_solve<KD>(S.f[1], w + KD);
    if (HB) {
        // base row of A w = r: v_b (v . w) / d0 = r_b fixes the stand multiplier s = (v . w) / d0 = r_b / v_b
        const double s = r[2 * KD] * S.piv;
        double acc = s * S.d0;
#pragma unroll
        for (int i = 0; i < 2 * KD; ++i) {
            w[i] = fma(-S.u[i], s, w[i]);
            acc = fma(-S.v[i], w[i], acc);
        }
        w[2 * KD] = acc * S.piv;
    } else {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 2 * KD; ++i) t = fma(S.v[i], w[i], t);
        t *= S.piv;
#pragma unroll
        for (int i = 0; i < 2 * KD; ++i) w[i] = fma(-S.u[i], t, w[i]);
    }
}

// Number of eigenvalues of A below sigma (sigma > 0), by the inertia of the bordered matrix.
template <int KD, bool HB>
IRLOSC_HD int sys_count_below(const TaskSys<KD, HB> &S, double sigma, bool *lost) {
    constexpr int KT = TaskSys<KD, HB>::KT;
    int n = 0;
    double q = 0.0;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        double f[KT], y[KD];
        n += blk_factor<KD>(S.D[b], sigma, f, lost);
#pragma unroll
        for (int i = 0; i < KD; ++i) y[i] = S.v[b * KD + i];
        q += blk_forward<KD>(f, y);
    }
    if (HB) {
        n += 1;                                              // the base's diagonal block is 0 - sigma < 0
        q -= S.v[2 * KD] * S.v[2 * KD] * rcp64(sigma);
    }
    const double last = -S.d0 - q;
    *lost = *lost || !(fabs(last) > 1e-13 * (S.d0 + fabs(q)));
    n += (last < 0.0) ? 1 : 0;
    return n - 1;                                            // the border -d0 itself is one negative
}

// How the task-space solve of an instance was resolved (Debug::how, tests)
enum TailHow : int {
    kHowInverse = 1,       // |det| >= 1e-4, or below it with no eigenvalue under the pinv cut-off
    kHowCut1 = 2,          // pinv removed one eigenvalue (deflation in the thread)
    kHowCut2 = 4,          // pinv removed two
    kHowWarp = 8,          // handed to the warp-cooperative eigen-solver
    // ... and why (statistics of the host harness)
    kWhyBlocks = 16,       // an arm block is not positive definite / lost a pivot
    kWhyBaseZero = 32,     // base row with a zero stand entry
    kWhyLost = 64,         // an inertia count lost a pivot to cancellation
    kWhyUndecided = 128,   // the bracket of lambda_max did not settle the count
    kWhyMany = 256,        // more than two eigenvalues under the cut-off
    kWhyNoConv = 512,      // inverse iteration did not converge
    kWhyResidual = 1024,   // residual test
};

constexpr int kBisectMax = 48;      // inertia evaluations while bracketing lambda_max
constexpr int kPowerSteps = 3;      // power steps for the Rayleigh-quotient lower bound on lambda_max
constexpr double kPowerMargin = 1.02;   // first upper-bound candidate: that far above the lower bound
constexpr int kIterMax = 40;        // inverse / subspace iterations for the cut eigenvectors
constexpr int kPlainIters = 8;      // ... of which plain ones before a single cut eigenvector takes a second vector along
constexpr int kRefineMax = 3;       // refinement steps of the task-space solution before the warp takes over

// The branch of osc.py:52-55 and its solution.  Returns false when the warp must finish the instance;
// *small_det then says whether the pinv branch is known to be the one taken.
template <int KD, bool HB>
IRLOSC_HD bool sys_resolve(TaskSys<KD, HB> &S, const double *gc, double *w, bool *small_det, int *how) {
    constexpr int K = TaskSys<KD, HB>::K;
    *small_det = false;
    *how = kHowWarp;
    // ---- factor the blocks; D must be positive definite for the block formulas to mean anything
    bool lost = false;
    int nneg = blk_factor<KD>(S.D[0], 0.0, S.f[0], &lost);
    nneg += blk_factor<KD>(S.D[1], 0.0, S.f[1], &lost);
    if (nneg != 0 || lost) { *how |= kWhyBlocks; return false; }
#pragma unroll
    for (int i = 0; i < 2 * KD; ++i) S.u[i] = S.v[i];
    blk_solve<KD>(S.f[0], S.u);
    blk_solve<KD>(S.f[1], S.u + KD);
    double detinv = S.d0;                               // 1 / det A
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int i = 0; i < KD; ++i) detinv *= S.f[b][ltri(i, i)];
    if (HB) {
        S.piv = rcp64(S.v[2 * KD]);
        detinv *= S.piv * S.piv;                        // det A = det D v_b^2 / d0
        if (!(fabs(S.v[2 * KD]) > 0.0)) { *how |= kWhyBaseZero; return false; }
    } else {
        double gam = S.d0;
#pragma unroll
        for (int i = 0; i < 2 * KD; ++i) gam = fma(S.v[i], S.u[i], gam);
        S.piv = rcp64(gam);
        detinv *= S.piv;                                // det A = det D (d0 + v^T D^-1 v) / d0
    }
    const bool small = !(fabs(detinv) <= 1.0 / kDetThreshold);      // |det A| < 1e-4 (osc.py:52)
    *small_det = small;
    // Frobenius norm and largest diagonal entry: max diag <= lambda_max <= ||A||_F
    double fro2 = 0.0, dmax = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const double a = sys_entry(S, i, j);
            fro2 = fma(a, (i == j) ? a : 2.0 * a, fro2);
            if (i == j) dmax = fmax(dmax, a);
        }
    const double fro = sqrt64(fro2);
    int m = 0;
    double hi_final = fro, lo_final = 0.0;
    if (small) {
        // bracket lambda_max until the count of eigenvalues under rcond * lambda_max is the same at both ends
        double hi = fro, lo = dmax, sigma = kPinvRcond * fro;
        int n_hi = -1, n_lo = -1, state = 0;
        bool moved_hi = false, decided = false;
        int n_counts = 0;
#pragma unroll 1
        for (int it = 0; it < kBisectMax && !decided; ++it) {
            const int n = sys_count_below(S, sigma, &lost);
            ++n_counts;
            if (lost) { *how |= kWhyLost; return false; }
            if (state == 0) {                           // cut-off from the upper bound ||A||_F
                n_hi = n;
                if (n == 0) { decided = true; break; }  // nothing can be cut (every k = 7 instance ends here)
                // Something lies under the loosest cut-off: sharpen the lower bound with the Rayleigh quotient of a few
                // power steps started from the column of the largest diagonal entry (within 1 % of lambda_max for 99 %
                // of the DualUR5 instances; ||A||_F is 60 % above it at k = 12, 13) ...
                double x[K], y[K];
#pragma unroll
                for (int i = 0; i < K; ++i) x[i] = (sys_entry(S, i, i) == dmax) ? 1.0 : 0.0;
                sys_matvec(S, x, y);
#pragma unroll 1
                for (int p = 0; p < kPowerSteps; ++p) {
                    double nn = 0.0, rho = 0.0;
#pragma unroll
                    for (int i = 0; i < K; ++i) nn = fma(y[i], y[i], nn);
                    nn = rcp64(sqrt64(nn));
#pragma unroll
                    for (int i = 0; i < K; ++i) x[i] = y[i] * nn;
                    sys_matvec(S, x, y);
#pragma unroll
                    for (int i = 0; i < K; ++i) rho = fma(x[i], y[i], rho);
                    lo = fmax(lo, rho * (1.0 - 1e-12));  // a Rayleigh quotient never exceeds lambda_max (rounding margin)
                }
                sigma = kPinvRcond * lo;
                state = 1;
            } else if (state == 1) {                    // cut-off from the lower bound
                n_lo = n;
                if (n_lo == n_hi) { decided = true; break; }
                sigma = lo * kPowerMargin;              // ... and try to confirm an upper bound just above it
                state = 2;
            } else if (state == 2) {                    // is lambda_max below sigma?
                moved_hi = (n == K);
                if (moved_hi) hi = sigma; else lo = sigma;
                sigma = kPinvRcond * (moved_hi ? hi : lo);
                state = 3;
            } else {                                    // recount at the end that moved
                if (moved_hi) n_hi = n; else n_lo = n;
                if (n_lo == n_hi) { decided = true; break; }
                sigma = sqrt64(lo * hi);
                state = 2;
            }
        }
        *how |= (n_counts & 0xff) << 12;
        if (!decided) { *how |= kWhyUndecided; return false; }
        hi_final = hi;
        lo_final = lo;
        m = n_hi;
        if (m > 2) { *how |= kWhyMany; return false; }
    }
    double geff[K];
#pragma unroll
    for (int i = 0; i < K; ++i) geff[i] = gc[i];
    double xa[K], xb[K];
    if (m >= 1) {
        // cut eigenvectors by inverse (subspace) iteration with the block solver
#pragma unroll
        for (int i = 0; i < K; ++i) { xa[i] = 1.0; xb[i] = (i & 1) ? -1.0 : 1.0; }
        bool conv = false, sub = false;
        int n_iter = 0;
#pragma unroll 1
        for (int it = 0; it < kIterMax && !conv; ++it) {
            double ya[K], yb[K];
            // One eigenvalue to cut and no convergence yet: lambda_1 / lambda_2 is close to one (the two straddle the
            // cut-off; 0.3 % of the cut instances) and plain inverse iteration crawls.  Iterate the span of the two
            // smallest eigenvectors instead (rate lambda_2 / lambda_3) and separate them by Rayleigh-Ritz at the end.
            if (m == 1 && it == kPlainIters) sub = true;
            sys_solve(S, xa, ya);
            double na = 0.0, dot = 0.0;
#pragma unroll
            for (int i = 0; i < K; ++i) na = fma(ya[i], ya[i], na);
            na = rcp64(sqrt64(na));
#pragma unroll
            for (int i = 0; i < K; ++i) { ya[i] *= na; dot = fma(ya[i], xa[i], dot); }
            double change = 0.0;
            if (m == 2 || sub) {
                sys_solve(S, xb, yb);
                double pab = 0.0, nb = 0.0;
#pragma unroll
                for (int i = 0; i < K; ++i) pab = fma(ya[i], yb[i], pab);
#pragma unroll
                for (int i = 0; i < K; ++i) { yb[i] = fma(-pab, ya[i], yb[i]); nb = fma(yb[i], yb[i], nb); }
                nb = rcp64(sqrt64(nb));
#pragma unroll
                for (int i = 0; i < K; ++i) yb[i] *= nb;
                // how far the new basis sticks out of the old span (the old basis is orthonormal after step one)
                double aa = 0.0, ab = 0.0, ba = 0.0, bb = 0.0;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    aa = fma(xa[i], ya[i], aa); ab = fma(xb[i], ya[i], ab);
                    ba = fma(xa[i], yb[i], ba); bb = fma(xb[i], yb[i], bb);
                }
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    change = fmax(change, fabs(ya[i] - aa * xa[i] - ab * xb[i]));
                    change = fmax(change, fabs(yb[i] - ba * xa[i] - bb * xb[i]));
                }
#pragma unroll
                for (int i = 0; i < K; ++i) xb[i] = yb[i];
            }
            if (m == 1) {                           // the single vector on its own (also while the second one rides along)
                const double sg = dot < 0.0 ? -1.0 : 1.0;
                double own = 0.0;
#pragma unroll
                for (int i = 0; i < K; ++i) own = fmax(own, fabs(fma(sg, ya[i], -xa[i])));
                if (!sub) change = own;
                else if (own < 1e-8) { change = own; sub = false; }      // it got there first: no Rayleigh-Ritz needed
            }
#pragma unroll
            for (int i = 0; i < K; ++i) xa[i] = ya[i];
            conv = (it >= 1) && (change < 1e-8);      // eigenvector error d ~ change: the deflation error is d^2 lambda_2 / lambda_1
            n_iter = it + 1;
        }
        *how |= (n_iter & 0xff) << 20;
        if (!conv) { *how |= kWhyNoConv; return false; }
        if (sub) {
            // Rayleigh-Ritz on span{xa, xb}: the eigenvector of the smaller Ritz value is the one to cut
            double av[K], bv[K];
            sys_matvec(S, xa, av);
            sys_matvec(S, xb, bv);
            double haa = 0.0, hab = 0.0, hbb = 0.0;
#pragma unroll
            for (int i = 0; i < K; ++i) { haa = fma(xa[i], av[i], haa); hab = fma(xa[i], bv[i], hab); hbb = fma(xb[i], bv[i], hbb); }
            const double mid = 0.5 * (haa + hbb), half = 0.5 * (hbb - haa), rad = sqrt64(fma(half, half, hab * hab));
            const double th1 = mid - rad, th2 = mid + rad;
            double c1 = hab, s1 = th1 - haa;
            const double c2 = th1 - hbb, s2 = hab;
            if (fma(c2, c2, s2 * s2) > fma(c1, c1, s1 * s1)) { c1 = c2; s1 = s2; }
            const double n2 = fma(c1, c1, s1 * s1);
            // exactly one eigenvalue lies under the cut-off: below rcond * lo, the next one above rcond * hi
            if (!(n2 > 0.0) || !(th1 < kPinvRcond * hi_final) || !(th2 > kPinvRcond * lo_final)) { *how |= kWhyNoConv; return false; }
            const double nr = rcp64(sqrt64(n2));
#pragma unroll
            for (int i = 0; i < K; ++i) xa[i] = (c1 * xa[i] + s1 * xb[i]) * nr;
        }
        // geff = P g
        double pa = 0.0, pb = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) { pa = fma(xa[i], geff[i], pa); if (m == 2) pb = fma(xb[i], geff[i], pb); }
#pragma unroll
        for (int i = 0; i < K; ++i) { geff[i] = fma(-pa, xa[i], geff[i]); if (m == 2) geff[i] = fma(-pb, xb[i], geff[i]); }
    }
    // onto the kept subspace (identity when nothing was cut)
    auto keep = [&](double *z) {
        if (m >= 1) {
            double pa = 0.0, pb = 0.0;
#pragma unroll
            for (int i = 0; i < K; ++i) { pa = fma(xa[i], z[i], pa); if (m == 2) pb = fma(xb[i], z[i], pb); }
#pragma unroll
            for (int i = 0; i < K; ++i) { z[i] = fma(-pa, xa[i], z[i]); if (m == 2) z[i] = fma(-pb, xb[i], z[i]); }
        }
    };
    sys_solve(S, geff, w);
    keep(w);
    // ---- residual of the structured system on the kept subspace.  The block formulas lose digits when an arm's
    // block is much worse conditioned than A itself (the stand joint supplies the direction the arm has lost):
    // the residual is exact enough to repair that by iterative refinement.
    double gmax = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) gmax = fmax(gmax, fabs(gc[i]));
    int passes = 0;
#pragma unroll 1
    for (;;) {
        double r[K];
        sys_matvec(S, w, r);
#pragma unroll
        for (int i = 0; i < K; ++i) r[i] -= geff[i];
        keep(r);
        double rmax = 0.0, wmax = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) { rmax = fmax(rmax, fabs(r[i])); wmax = fmax(wmax, fabs(w[i])); }
        if (rmax <= 1e-10 * fma(fro, wmax, gmax)) break;
        if (passes == kRefineMax || !(rmax == rmax)) { *how |= kWhyResidual; return false; }
        ++passes;
        double dw[K];
        sys_solve(S, r, dw);
        keep(dw);
#pragma unroll
        for (int i = 0; i < K; ++i) w[i] -= dw[i];
    }
    *how |= passes << 28;
    *how = (*how & ~0xfff) | (m == 0 ? kHowInverse : m == 1 ? kHowCut1 : kHowCut2);
    return true;
}

// ------------------------------------------------------------------------------------------------------
// Record a thread hands to its warp when sys_resolve gives up (layout: Rec<KD, HAS_BASE>, canonical rows).
template <int KD, bool HAS_BASE, class JA>
IRLOSC_HD void tail_record(const FRoles &R, const double (*akA)[KD * (KD + 1) / 2], const double *j0, const double *g,
                           const JA &ja, const double (*base_arm)[6], double base_st, double inv0, bool force_pinv,
                           double *rec) {
    using RC = Rec<KD, HAS_BASE>;
    constexpr int K = RC::K;
    int perm[K];
#pragma unroll
    for (int i = 0; i < K; ++i) perm[i] = (i < KD) ? R.row_arm[0] + i : (i < 2 * KD) ? R.row_arm[1] + i - KD : R.row_base;
#pragma unroll 1
    for (int i = 0; i < K; ++i) {
        const double ji = j0[perm[i]] * inv0;
        for (int j = 0; j < K; ++j) {
            const int hi = i > j ? i : j, lo = i > j ? j : i;
            double bv = 0.0;
            if (hi < KD) bv = akA[0][hi * (hi + 1) / 2 + lo];
            else if (hi < 2 * KD && lo >= KD) bv = akA[1][(hi - KD) * (hi - KD + 1) / 2 + (lo - KD)];
            rec[RC::A + i * K + j] = fma(ji, j0[perm[j]], bv);
        }
        rec[RC::G + i] = g[perm[i]];
        rec[RC::JST + i] = ja.stand(i);
    }
    rec[RC::BASE] = base_st;
    for (int am = 0; am < 2; ++am)
        for (int i = 0; i < 6; ++i) {
            rec[RC::BASE + 1 + 6 * am + i] = base_arm[am][i];
            for (int cr = 0; cr < KD; ++cr) rec[RC::JARM + (am * 6 + i) * KD + cr] = ja.arm(am, i, cr);
        }
    rec[RC::ABAD] = force_pinv ? 0.0 : 1.0;      // 1: the eigen-solver decides the branch from its own determinant
}

// Jacobian accessor over plain arrays (fused and streaming kernels keep the entries they have seen).
template <int KD>
struct JArrays {
    const double *jst;                 // [task row]
    const double (*jarm)[6][KD];       // [arm][joint][row of the arm]
    const int *perm;                   // canonical -> task row
    IRLOSC_HD double stand(int canon) const { return jst[perm[canon]]; }
    IRLOSC_HD double arm(int am, int i, int cr) const { return jarm[am][i][cr]; }
};

// What a thread-per-instance kernel carries from the elimination into the tail, and out of it for the warp finish.
template <int KD, bool HAS_BASE>
struct TailState {
    static constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    static constexpr int KT = KD * (KD + 1) / 2;
    double akA[2][KT];                   // the arms' diagonal blocks of A before the stand joint
    double j0[K], jst[K], dxr[K], g[K];  // stand column of the reduced / original J, dx = J dq, task signal (task-row order)
    double jarm[2][6][KD], base_arm[2][6];
    double base_st, inv0;
    double *u_all_row, *ctrl_row;
    uint8_t *status;
    bool force_pinv;
};

// Returns true when the instance must be finished by its warp (tail_record + tail_warp_finish).
//   ja: original Jacobian entries, ja.stand(canonical row), ja.arm(arm, joint 0..5, row of the arm)
template <int KD, bool HAS_BASE, class JA>
IRLOSC_HD bool osc_tail(const KParams &P, const FRoles &R, const double *target_vel, unsigned vel_zero, int flags,
                        bool m_ok, const double (*akA)[KD * (KD + 1) / 2], const double *j0, const double *dxr,
                        double *g, const JA &ja, const double (*base_arm)[6], double base_st, double inv0,
                        double *u_all_row, double *ctrl_row, uint8_t *status, bool *force_pinv, const Debug *dbg) {
    constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    constexpr int N = kN;
    constexpr int KT = KD * (KD + 1) / 2;
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

    // ------------------------------------------------------------ A w = g on the block structure
    // CANONICAL row order (arm 0 rows, arm 1 rows, base row) so that every index below is static; det,
    // the spectrum and the solution are invariant under the symmetric permutation.
    int perm[K];
#pragma unroll
    for (int i = 0; i < K; ++i) perm[i] = (i < KD) ? R.row_arm[0] + i : (i < 2 * KD) ? R.row_arm[1] + i - KD : R.row_base;
    TaskSys<KD, HAS_BASE> S;
    double gc[K], w[K];
#pragma unroll
    for (int i = 0; i < K; ++i) { S.v[i] = j0[perm[i]]; gc[i] = g[perm[i]]; }
#pragma unroll
    for (int e = 0; e < KT; ++e) { S.D[0][e] = akA[0][e]; S.D[1][e] = akA[1][e]; }
    S.inv0 = inv0;
    S.d0 = rcp64(inv0);
    if (dbg && dbg->A) {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double a = sys_entry(S, i, j);
                dbg->A[perm[i] * K + perm[j]] = a;
                dbg->A[perm[j] * K + perm[i]] = a;
            }
        for (int i = 0; i < K; ++i) { dbg->g[i] = g[i]; dbg->dx[i] = dxr[i]; }
    }
    if (!m_ok) flags |= IRLOSC_ST_M_NOT_PD;
    const bool poison = (flags & (IRLOSC_ST_M_NOT_PD | IRLOSC_ST_DX_RANGE)) != 0;
    bool small_det = false;
    int how = 0;
    bool solved = true;
    if (!poison) solved = sys_resolve(S, gc, w, &small_det, &how);
    if (solved && small_det) flags |= IRLOSC_ST_PINV;
    if (dbg && dbg->how) *dbg->how = how;
    *force_pinv = small_det;
    const bool hard = !poison && !solved;

    // ------------------------------------------------------------ joint-space assembly + packing
    if (!hard && !poison) {
        {
            double jt = 0.0;
#pragma unroll
            for (int r = 0; r < K; ++r) jt = fma(ja.stand(r), w[r], jt);
            put_joint(R, u_all_row, ctrl_row, 0, base_st - jt);
        }
#pragma unroll
        for (int am = 0; am < 2; ++am) {
            // all Jacobian entries of the arm first (they may be re-read from memory: one latency, not one per joint)
            double jv[6][KD];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int cr = 0; cr < KD; ++cr) jv[i][cr] = ja.arm(am, i, cr);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double jt = 0.0;
#pragma unroll
                for (int cr = 0; cr < KD; ++cr) jt = fma(jv[i][cr], w[am * KD + cr], jt);
                put_joint(R, u_all_row, ctrl_row, 1 + 12 * am + i, base_arm[am][i] - jt);
            }
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
        for (int r = 0; r < K; ++r) dbg->J[perm[r] * N] = ja.stand(r);
        for (int am = 0; am < 2; ++am)
            for (int i = 0; i < 6; ++i)
                for (int cr = 0; cr < KD; ++cr) dbg->J[(R.row_arm[am] + cr) * N + 1 + 12 * am + i] = ja.arm(am, i, cr);
    }
    return hard;
}

// osc_tail on a TailState (fused and streaming kernels)
template <int KD, bool HAS_BASE>
IRLOSC_HD bool state_tail(const KParams &P, const FRoles &R, const double *target_vel, unsigned vel_zero, int flags, bool m_ok,
                          TailState<KD, HAS_BASE> &T, const Debug *dbg) {
    constexpr int K = TailState<KD, HAS_BASE>::K;
    int perm[K];
#pragma unroll
    for (int i = 0; i < K; ++i) perm[i] = (i < KD) ? R.row_arm[0] + i : (i < 2 * KD) ? R.row_arm[1] + i - KD : R.row_base;
    const JArrays<KD> ja{T.jst, T.jarm, perm};
    return osc_tail<KD, HAS_BASE>(P, R, target_vel, vel_zero, flags, m_ok, T.akA, T.j0, T.dxr, T.g, ja, T.base_arm, T.base_st,
                                  T.inv0, T.u_all_row, T.ctrl_row, T.status, &T.force_pinv, dbg);
}
template <int KD, bool HAS_BASE>
IRLOSC_HD void state_record(const FRoles &R, const TailState<KD, HAS_BASE> &T, double *rec) {
    constexpr int K = TailState<KD, HAS_BASE>::K;
    int perm[K];
#pragma unroll
    for (int i = 0; i < K; ++i) perm[i] = (i < KD) ? R.row_arm[0] + i : (i < 2 * KD) ? R.row_arm[1] + i - KD : R.row_base;
    const JArrays<KD> ja{T.jst, T.jarm, perm};
    tail_record<KD, HAS_BASE>(R, T.akA, T.j0, T.g, ja, T.base_arm, T.base_st, T.inv0, T.force_pinv, rec);
}

// Finish one recorded instance given w = pinv(A) g (canonical order): the 13 joints that have Jacobian columns.
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

}  // namespace fused
}  // namespace irlosc

#if defined(__CUDACC__) && !defined(IRLOSC_FUSED_NO_KERNELS)
#include "osc_eigen.cuh"

namespace irlosc {
namespace fused {

// Per-warp shared memory of the cooperative finish: the record (its A block is decomposed in place),
// the eigenvectors and the rotation / coefficient buffers.
template <int KD, bool HAS_BASE>
struct WarpFix {
    static constexpr int K = Rec<KD, HAS_BASE>::K;
    double rec[Rec<KD, HAS_BASE>::SIZE];
    double Vs[K][K + 1];
    double w[K], cbuf[16], sbuf[16];
    int flags, pad_;
};

// Called by a whole warp after a tile: every lane whose instance sys_resolve could not decide (`hard`) writes its
// record with `write_rec(rec)` in turn, the warp runs the Jacobi eigen-solver on it (osc.py:52-55 on the
// eigenvalues), and the owner rewrites its chain joints with `finish(rec, w, flags)`.
template <int KD, bool HAS_BASE, class WR, class FIN>
__device__ __forceinline__ void tail_warp_finish(WarpFix<KD, HAS_BASE> &F, bool hard, int lane, WR write_rec, FIN finish) {
    using RC = Rec<KD, HAS_BASE>;
    constexpr int K = RC::K;
    static_assert(RC::A == 0 && K <= 16, "record layout");
    unsigned mask = __ballot_sync(0xffffffffu, hard);
    while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        if (lane == src) { write_rec(F.rec); F.flags = 0; }
        __syncwarp();
        tiled::eigen_solve<K, K>(reinterpret_cast<double (*)[K]>(F.rec + RC::A), F.Vs, F.rec + RC::G, F.w, F.cbuf, F.sbuf,
                                 F.rec[RC::ABAD] == 0.0, lane, &F.flags);
        __syncwarp();
        if (lane == src) finish(F.rec, F.w, F.flags);
        __syncwarp();
    }
}

// Warp finish of the instances whose TailState says `hard`.
template <int KD, bool HAS_BASE>
__device__ __forceinline__ void state_warp_finish(WarpFix<KD, HAS_BASE> &F, const FRoles &R, const TailState<KD, HAS_BASE> &T,
                                                  bool hard, int lane) {
    tail_warp_finish<KD, HAS_BASE>(
        F, hard, lane, [&](double *rec) { state_record<KD, HAS_BASE>(R, T, rec); },
        [&](const double *rec, const double *w, int fl) {
            fixup_finish<KD, HAS_BASE>(R, T.u_all_row, T.ctrl_row, rec, w, 0, 1);
            if (T.status) *T.status = (uint8_t)(*T.status | fl);
        });
}

}  // namespace fused
}  // namespace irlosc
#endif

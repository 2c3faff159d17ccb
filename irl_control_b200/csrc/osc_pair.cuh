// Pair kernel (sm_100a): TWO lanes per robot instance, one per arm, on the batch-interleaved tiles of osc_lane.cuh.
//
// Why.  osc_step_lane gives an instance to one thread.  ncu (profiles/r02_lane_*): issue slots 20 - 25 % used, FP64
// pipe under 20 %, HBM at half its rate - the kernel is bound by the LENGTH of the dependent chain a thread walks
// (5 900 instructions at k = 7, 10 000 at k = 12) times the latency per instruction, with 7 - 8 warps per SM because
// the thread needs all 255 registers (and still spills 1.5 KB at k = 12).  The DualUR5 tree makes the two arms
// independent until the stand joint, and the task-space matrix is blockdiag(D0, D1, [0]) + v v^T / d0: a lane per arm
// halves the chain, halves the state a thread holds (no spills, more warps per SM), and what couples the arms is a
// handful of scalar reductions, each one __shfl_xor between the two lanes of the pair:
//
//   * elimination: the lane runs the per-arm consumers of osc_stream.cuh (consume_cc / consume_grip / consume_rows)
//     on ITS arm's groups - the loop over the arms of lane_instance is gone;
//   * stand pivot d0 = c0[arm 0] + c0[arm 1], (M dq)_stand likewise: one reduction each;
//   * task-space solve (osc.py:41-56) on the block structure exactly as osc_tail.cuh does it, with every sum over
//     the 2 KD arm rows split into the lane's KD rows + the partner's partial: LDL^T of the own block, determinant
//     test, inertia counts, Rayleigh bracket, deflation, residual.  Decisions are taken on reduced values, which are
//     bit-identical in both lanes (a + b and b + a), so the two lanes of a pair never diverge from each other;
//   * joint-space assembly: the lane owns its arm's six joints and gripper; lane 0 of the pair also the stand joint.
//
// A warp therefore owns a HALF tile (16 instances): lane = 2 * (instance in the half) + arm.  Tiles keep their
// 32-instance layout; a load instruction of the warp touches two 128-byte runs (the arm-0 entry and the arm-1 entry
// of the 16 instances), every sector fully used.  The two halves of a tile go to neighbouring warps of one CTA.
//
// Instances the pair cannot decide (same criteria as osc_tail.cuh) are finished by the whole warp with the Jacobi
// eigen-solver (osc_eigen.cuh), as in the other kernels.
//
// Reference restated: ir-lab/irl_control osc.py:41-68, 120-210; robot.py:44-72; device.py:115-170.
#pragma once
#include "osc_lane.cuh"

#if defined(__CUDACC__) && !defined(IRLOSC_FUSED_NO_KERNELS)
namespace irlosc {
namespace pair {

using fused::FRoles;
using fused::blk_factor;
using fused::blk_forward;
using fused::blk_solve;
using fused::ltri;
using fused::rcp64;
using fused::sqrt64;
using fused::kN;
using lane::kTile;
using lane::LaneArgs;
using stream::kGroups;

constexpr int kHalf = 16;               // instances per warp

// ---------------------------------------------------------------- reductions over the two lanes of a pair
struct Duo {
    unsigned mask;                      // the pair's two lanes (shuffles are legal under divergence between pairs)
    int arm;                            // which of the two this lane is
    __device__ __forceinline__ double other(double x) const { return __shfl_xor_sync(mask, x, 1); }
    __device__ __forceinline__ double sum(double x) const { return x + other(x); }        // same bits in both lanes
    __device__ __forceinline__ double max(double x) const { return fmax(x, other(x)); }
    __device__ __forceinline__ int sum(int x) const { return x + __shfl_xor_sync(mask, x, 1); }
    // (the shuffle first: `x || shuffle` would skip it in the lane whose x decides, and strand the partner)
    __device__ __forceinline__ bool any(bool x) const { const int o = __shfl_xor_sync(mask, (int)x, 1); return x || o != 0; }
    __device__ __forceinline__ bool all(bool x) const { const int o = __shfl_xor_sync(mask, (int)x, 1); return x && o != 0; }
};

// A = blockdiag(D0, D1, [0]) + v v^T / d0 (osc_tail.cuh: TaskSys), the lane's share: block D of its arm and the
// arm's rows of v; the base row's entry vb and the scalars are held by both lanes.
template <int KD, bool HB>
struct PairSys {
    static constexpr int KT = KD * (KD + 1) / 2;
    double D[KT], f[KT], v[KD], u[KD];
    double vb, d0, inv0, piv;
    Duo duo;
};

// y = A x  (x, y: the lane's rows; xb, yb: base row, identical in both lanes)
template <int KD, bool HB>
__device__ __forceinline__ void pair_matvec(const PairSys<KD, HB> &S, const double *x, double xb, double *y, double *yb) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < KD; ++i) s = fma(S.v[i], x[i], s);
    s = S.duo.sum(s);
    if (HB) s = fma(S.vb, xb, s);
    s *= S.inv0;
#pragma unroll
    for (int i = 0; i < KD; ++i) {
        double z = S.v[i] * s;
#pragma unroll
        for (int j = 0; j < KD; ++j) z = fma(S.D[i >= j ? ltri(i, j) : ltri(j, i)], x[j], z);
        y[i] = z;
    }
    *yb = HB ? S.vb * s : 0.0;
}

// w = A^-1 r
template <int KD, bool HB>
__device__ __forceinline__ void pair_solve(const PairSys<KD, HB> &S, const double *r, double rb, double *w, double *wb) {
#pragma unroll
    for (int i = 0; i < KD; ++i) w[i] = r[i];
    blk_solve<KD>(S.f, w);
    if (HB) {
        const double s = rb * S.piv;                        // the stand multiplier (osc_tail.cuh: sys_solve)
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < KD; ++i) {
            w[i] = fma(-S.u[i], s, w[i]);
            acc = fma(S.v[i], w[i], acc);
        }
        *wb = (s * S.d0 - S.duo.sum(acc)) * S.piv;
    } else {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < KD; ++i) t = fma(S.v[i], w[i], t);
        t = S.duo.sum(t) * S.piv;
#pragma unroll
        for (int i = 0; i < KD; ++i) w[i] = fma(-S.u[i], t, w[i]);
        *wb = 0.0;
    }
}

// Number of eigenvalues of A below sigma (osc_tail.cuh: sys_count_below); *lost is pair-uniform.
template <int KD, bool HB>
__device__ __forceinline__ int pair_count_below(const PairSys<KD, HB> &S, double sigma, bool *lost) {
    constexpr int KT = PairSys<KD, HB>::KT;
    double f[KT], y[KD];
    bool l = false;
    int n = blk_factor<KD>(S.D, sigma, f, &l);
#pragma unroll
    for (int i = 0; i < KD; ++i) y[i] = S.v[i];
    double q = blk_forward<KD>(f, y);
    n = S.duo.sum(n);
    q = S.duo.sum(q);
    if (HB) {
        n += 1;
        q -= S.vb * S.vb * rcp64(sigma);
    }
    const double last = -S.d0 - q;
    l = l || !(fabs(last) > 1e-13 * (S.d0 + fabs(q)));
    *lost = *lost || S.duo.any(l);
    n += (last < 0.0) ? 1 : 0;
    return n - 1;
}

template <int KD>
__device__ __forceinline__ double dot_own(const double *a, const double *b) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < KD; ++i) s = fma(a[i], b[i], s);
    return s;
}

// osc.py:52-55 and its solution, pair form of osc_tail.cuh: sys_resolve (same constants, same decisions).
// Returns false when the warp must finish the instance.  Every branch is taken by both lanes of the pair.
template <int KD, bool HB>
__device__ bool pair_resolve(PairSys<KD, HB> &S, const double *gc, double gcb, double *w, double *wb, bool *small_det) {
    using namespace fused;
    const Duo &duo = S.duo;
    *small_det = false;
    bool lost = false;
    {
        const int nneg = blk_factor<KD>(S.D, 0.0, S.f, &lost);
        if (duo.sum(nneg) != 0 || duo.any(lost)) return false;
    }
    lost = false;
#pragma unroll
    for (int i = 0; i < KD; ++i) S.u[i] = S.v[i];
    blk_solve<KD>(S.f, S.u);
    double detinv;
    {
        double dl = 1.0;
#pragma unroll
        for (int i = 0; i < KD; ++i) dl *= S.f[ltri(i, i)];
        detinv = S.d0 * (dl * duo.other(dl));
    }
    if (HB) {
        S.piv = rcp64(S.vb);
        detinv *= S.piv * S.piv;
        if (!(fabs(S.vb) > 0.0)) return false;
    } else {
        S.piv = rcp64(S.d0 + duo.sum(dot_own<KD>(S.v, S.u)));
        detinv *= S.piv;
    }
    const bool small = !(fabs(detinv) <= 1.0 / kDetThreshold);
    *small_det = small;
    // Frobenius norm and largest diagonal entry
    double diag[KD], fro, dmax;
    {
        double own = 0.0, dm = 0.0;
#pragma unroll
        for (int i = 0; i < KD; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double a = fma(S.v[i] * S.inv0, S.v[j], S.D[ltri(i, j)]);
                own = fma(a, (i == j) ? a : 2.0 * a, own);
                if (i == j) { diag[i] = a; dm = fmax(dm, a); }
            }
        const double n2 = dot_own<KD>(S.v, S.v), n2o = duo.other(n2);
        double fro2 = fma(2.0 * (S.inv0 * S.inv0) * n2, n2o, duo.sum(own));      // the block coupling the two arms, twice
        dmax = duo.max(dm);
        if (HB) {
            const double ab = S.vb * S.inv0, bb = ab * S.vb;
            fro2 = fma(2.0 * ab * ab, n2 + n2o, fro2);
            fro2 = fma(bb, bb, fro2);
            dmax = fmax(dmax, bb);
        }
        fro = sqrt64(fro2);
    }
    int m = 0;
    double hi_final = fro, lo_final = 0.0;
    if (small) {
        double hi = fro, lo = dmax, sigma = kPinvRcond * fro;
        int n_hi = -1, n_lo = -1, state = 0;
        bool moved_hi = false, decided = false;
#pragma unroll 1
        for (int it = 0; it < kBisectMax && !decided; ++it) {
            const int n = pair_count_below(S, sigma, &lost);
            if (lost) return false;
            if (state == 0) {
                n_hi = n;
                if (n == 0) { decided = true; break; }
                double x[KD], y[KD], xb = 0.0, yb;
#pragma unroll
                for (int i = 0; i < KD; ++i) x[i] = (diag[i] == dmax) ? 1.0 : 0.0;
                if (HB) xb = ((S.vb * S.inv0) * S.vb == dmax) ? 1.0 : 0.0;
                pair_matvec(S, x, xb, y, &yb);
#pragma unroll 1
                for (int p = 0; p < kPowerSteps; ++p) {
                    double nn = duo.sum(dot_own<KD>(y, y));
                    if (HB) nn = fma(yb, yb, nn);
                    nn = rcp64(sqrt64(nn));
#pragma unroll
                    for (int i = 0; i < KD; ++i) x[i] = y[i] * nn;
                    xb = yb * nn;
                    pair_matvec(S, x, xb, y, &yb);
                    double rho = duo.sum(dot_own<KD>(x, y));
                    if (HB) rho = fma(xb, yb, rho);
                    lo = fmax(lo, rho * (1.0 - 1e-12));
                }
                sigma = kPinvRcond * lo;
                state = 1;
            } else if (state == 1) {
                n_lo = n;
                if (n_lo == n_hi) { decided = true; break; }
                sigma = lo * kPowerMargin;
                state = 2;
            } else if (state == 2) {
                moved_hi = (n == 2 * KD + (HB ? 1 : 0));
                if (moved_hi) hi = sigma; else lo = sigma;
                sigma = kPinvRcond * (moved_hi ? hi : lo);
                state = 3;
            } else {
                if (moved_hi) n_hi = n; else n_lo = n;
                if (n_lo == n_hi) { decided = true; break; }
                sigma = sqrt64(lo * hi);
                state = 2;
            }
        }
        if (!decided) return false;
        hi_final = hi;
        lo_final = lo;
        m = n_hi;
        if (m > 2) return false;
    }
    double geff[KD], geb = gcb;
#pragma unroll
    for (int i = 0; i < KD; ++i) geff[i] = gc[i];
    double xa[KD], xb[KD], xab = 0.0, xbb = 0.0;          // cut eigenvectors: own rows, base row
    if (m >= 1) {
        // start vectors in canonical row numbering: all ones, alternating signs (osc_tail.cuh)
        const int row0 = duo.arm * KD;
#pragma unroll
        for (int i = 0; i < KD; ++i) { xa[i] = 1.0; xb[i] = ((row0 + i) & 1) ? -1.0 : 1.0; }
        if (HB) { xab = 1.0; xbb = ((2 * KD) & 1) ? -1.0 : 1.0; }
        bool conv = false, sub = false;
#pragma unroll 1
        for (int it = 0; it < kIterMax && !conv; ++it) {
            double ya[KD], yb[KD], yab, ybb = 0.0;
            if (m == 1 && it == kPlainIters) sub = true;      // second vector rides along, Rayleigh-Ritz at the end (osc_tail.cuh)
            pair_solve(S, xa, xab, ya, &yab);
            double na = duo.sum(dot_own<KD>(ya, ya));
            if (HB) na = fma(yab, yab, na);
            na = rcp64(sqrt64(na));
#pragma unroll
            for (int i = 0; i < KD; ++i) ya[i] *= na;
            yab *= na;
            double dot = duo.sum(dot_own<KD>(ya, xa));
            if (HB) dot = fma(yab, xab, dot);
            double change = 0.0;
            if (m == 2 || sub) {
                pair_solve(S, xb, xbb, yb, &ybb);
                double pab = duo.sum(dot_own<KD>(ya, yb));
                if (HB) pab = fma(yab, ybb, pab);
#pragma unroll
                for (int i = 0; i < KD; ++i) yb[i] = fma(-pab, ya[i], yb[i]);
                ybb = fma(-pab, yab, ybb);
                double nb = duo.sum(dot_own<KD>(yb, yb));
                if (HB) nb = fma(ybb, ybb, nb);
                nb = rcp64(sqrt64(nb));
#pragma unroll
                for (int i = 0; i < KD; ++i) yb[i] *= nb;
                ybb *= nb;
                double aa = dot_own<KD>(xa, ya), ab = dot_own<KD>(xb, ya), ba = dot_own<KD>(xa, yb), bb = dot_own<KD>(xb, yb);
                aa = duo.sum(aa); ab = duo.sum(ab); ba = duo.sum(ba); bb = duo.sum(bb);
                if (HB) { aa = fma(xab, yab, aa); ab = fma(xbb, yab, ab); ba = fma(xab, ybb, ba); bb = fma(xbb, ybb, bb); }
#pragma unroll
                for (int i = 0; i < KD; ++i) {
                    change = fmax(change, fabs(ya[i] - aa * xa[i] - ab * xb[i]));
                    change = fmax(change, fabs(yb[i] - ba * xa[i] - bb * xb[i]));
                }
                if (HB) {
                    change = fmax(change, fabs(yab - aa * xab - ab * xbb));
                    change = fmax(change, fabs(ybb - ba * xab - bb * xbb));
                }
#pragma unroll
                for (int i = 0; i < KD; ++i) xb[i] = yb[i];
                xbb = ybb;
            }
            change = duo.max(change);
            if (m == 1) {
                const double sg = dot < 0.0 ? -1.0 : 1.0;
                double own = 0.0;
#pragma unroll
                for (int i = 0; i < KD; ++i) own = fmax(own, fabs(fma(sg, ya[i], -xa[i])));
                if (HB) own = fmax(own, fabs(fma(sg, yab, -xab)));
                own = duo.max(own);
                if (!sub) change = own;
                else if (own < 1e-8) { change = own; sub = false; }
            }
#pragma unroll
            for (int i = 0; i < KD; ++i) xa[i] = ya[i];
            xab = yab;
            conv = (it >= 1) && (change < 1e-8);
        }
        if (!conv) return false;
        if (sub) {                                       // Rayleigh-Ritz on span{xa, xb}
            double av[KD], bv[KD], avb, bvb;
            pair_matvec(S, xa, xab, av, &avb);
            pair_matvec(S, xb, xbb, bv, &bvb);
            double haa = duo.sum(dot_own<KD>(xa, av)), hab = duo.sum(dot_own<KD>(xa, bv)), hbb = duo.sum(dot_own<KD>(xb, bv));
            if (HB) { haa = fma(xab, avb, haa); hab = fma(xab, bvb, hab); hbb = fma(xbb, bvb, hbb); }
            const double mid = 0.5 * (haa + hbb), half = 0.5 * (hbb - haa), rad = sqrt64(fma(half, half, hab * hab));
            const double th1 = mid - rad, th2 = mid + rad;
            double c1 = hab, s1 = th1 - haa;
            const double c2 = th1 - hbb, s2 = hab;
            if (fma(c2, c2, s2 * s2) > fma(c1, c1, s1 * s1)) { c1 = c2; s1 = s2; }
            const double n2 = fma(c1, c1, s1 * s1);
            if (!(n2 > 0.0) || !(th1 < kPinvRcond * hi_final) || !(th2 > kPinvRcond * lo_final)) return false;
            const double nr = rcp64(sqrt64(n2));
#pragma unroll
            for (int i = 0; i < KD; ++i) xa[i] = (c1 * xa[i] + s1 * xb[i]) * nr;
            if (HB) xab = (c1 * xab + s1 * xbb) * nr;
        }
        double pa = duo.sum(dot_own<KD>(xa, geff)), pb = 0.0;
        if (HB) pa = fma(xab, geb, pa);
        if (m == 2) {
            pb = duo.sum(dot_own<KD>(xb, geff));
            if (HB) pb = fma(xbb, geb, pb);
        }
#pragma unroll
        for (int i = 0; i < KD; ++i) { geff[i] = fma(-pa, xa[i], geff[i]); if (m == 2) geff[i] = fma(-pb, xb[i], geff[i]); }
        if (HB) { geb = fma(-pa, xab, geb); if (m == 2) geb = fma(-pb, xbb, geb); }
    }
    auto keep = [&](double *z, double *zb) {              // onto the kept subspace
        if (m >= 1) {
            double pa = duo.sum(dot_own<KD>(xa, z)), pb = 0.0;
            if (HB) pa = fma(xab, *zb, pa);
            if (m == 2) {
                pb = duo.sum(dot_own<KD>(xb, z));
                if (HB) pb = fma(xbb, *zb, pb);
            }
#pragma unroll
            for (int i = 0; i < KD; ++i) { z[i] = fma(-pa, xa[i], z[i]); if (m == 2) z[i] = fma(-pb, xb[i], z[i]); }
            if (HB) { *zb = fma(-pa, xab, *zb); if (m == 2) *zb = fma(-pb, xbb, *zb); }
        }
    };
    pair_solve(S, geff, geb, w, wb);
    keep(w, wb);
    // residual on the kept subspace, iterative refinement (osc_tail.cuh)
    double gmax = 0.0;
#pragma unroll
    for (int i = 0; i < KD; ++i) gmax = fmax(gmax, fabs(gc[i]));
    if (HB) gmax = fmax(gmax, fabs(gcb));
    gmax = duo.max(gmax);
    int passes = 0;
#pragma unroll 1
    for (;;) {
        double r[KD], rb;
        pair_matvec(S, w, *wb, r, &rb);
#pragma unroll
        for (int i = 0; i < KD; ++i) r[i] -= geff[i];
        rb -= geb;
        keep(r, &rb);
        double rmax = 0.0, wmax = 0.0;
#pragma unroll
        for (int i = 0; i < KD; ++i) { rmax = fmax(rmax, fabs(r[i])); wmax = fmax(wmax, fabs(w[i])); }
        if (HB) { rmax = fmax(rmax, fabs(rb)); wmax = fmax(wmax, fabs(*wb)); }
        rmax = duo.max(rmax);
        wmax = duo.max(wmax);
        if (rmax <= 1e-10 * fma(fro, wmax, gmax)) break;
        if (passes == kRefineMax || !(rmax == rmax)) return false;
        ++passes;
        double dw[KD], dwb;
        pair_solve(S, r, rb, dw, &dwb);
        keep(dw, &dwb);
#pragma unroll
        for (int i = 0; i < KD; ++i) w[i] -= dw[i];
        *wb -= dwb;
    }
    return true;
}

// What follows the elimination, for the two lanes of an instance (all 32 lanes of the warp call this together: the
// warp-cooperative finish is inside).  The lane brings S (its block D, its rows of v, vb, d0, inv0), its rows of the
// task signal gc and of dx = J dq, the base row's gcb / dxb (both lanes), its arm's joint terms base_arm, the stand's
// base_st; ja(cr, i) returns J[row cr of the lane's arm][i = 0: stand, 1..6: the arm's joints].
template <int KD, bool HAS_BASE, class JA>
__device__ __forceinline__ void pair_tail(const KParams &P, const FRoles &R, PairSys<KD, HAS_BASE> &S, double *gc, double gcb,
                                          const double *dxa, double dxb, const double *base_arm, double base_st, bool m_ok,
                                          const JA &ja, const double *target_vel, unsigned vel_zero, int flags, double *u_all_row,
                                          double *ctrl_row, uint8_t *status, bool valid, fused::WarpFix<KD, HAS_BASE> &wfix, int lane) {
    using RC = fused::Rec<KD, HAS_BASE>;
    constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    const Duo duo = S.duo;
    const int arm = duo.arm, jb = 1 + 12 * arm, D = P.D;
    const double jb0 = S.vb;
    // ---- velocity-tracking term (osc.py:175-177) and the null-space part of g
    if (target_vel != nullptr && ((~vel_zero) & ((1u << D) - 1u)) != 0u) {
        double dxf[K];                                   // dx in task-row order (N3: dx_idx may point anywhere)
#pragma unroll
        for (int cr = 0; cr < KD; ++cr) {
            const double o = duo.other(dxa[cr]);
            dxf[R.row_arm[arm] + cr] = dxa[cr];
            dxf[R.row_arm[arm ^ 1] + cr] = o;
        }
        if (HAS_BASE) dxf[R.row_base] = dxb;
#pragma unroll 1
        for (int d = 0; d < D; ++d) {
            if ((vel_zero >> d) & 1u) continue;
            const KDevice &dv = P.dev[d];
            flags |= IRLOSC_ST_VEL_BRANCH;
            const bool mine = d == R.dev_arm[arm], base = HAS_BASE && d == R.dev_base;
            int r = 0;
            for (int i = 0; i < 6; ++i)
                if (dv.dof[i]) {
                    const int src = dv.dx_idx[r];
                    if (src >= K) flags |= IRLOSC_ST_DX_RANGE;
                    else {
                        const double add = dv.kv * (dxf[src] - target_vel[d * 6 + i]) * dv.damp[i];
                        if (mine) {
#pragma unroll
                            for (int s = 0; s < KD; ++s)
                                if (r == s) gc[s] += add;
                        }
                        if (base) gcb += add;
                    }
                    ++r;
                }
        }
    }
    {
        const double kvn = P.has_nullspace ? P.nullspace_kv : 0.0;
#pragma unroll
        for (int r = 0; r < KD; ++r) gc[r] = fma(-kvn, dxa[r], gc[r]);
        if (HAS_BASE) gcb = fma(-kvn, dxb, gcb);
    }
    if (!m_ok) flags |= IRLOSC_ST_M_NOT_PD;
    const bool poison = (flags & (IRLOSC_ST_M_NOT_PD | IRLOSC_ST_DX_RANGE)) != 0;
    double w[KD], wb = 0.0;
    bool small_det = false, solved = true;
    if (!poison) solved = pair_resolve(S, gc, gcb, w, &wb, &small_det);
    if (solved && small_det) flags |= IRLOSC_ST_PINV;
    const bool hard = !poison && !solved;
    // ---- joint-space assembly + packing
    if (!hard && !poison) {
        double jv[7][KD];
#pragma unroll
        for (int i = 0; i < 7; ++i)
#pragma unroll
            for (int cr = 0; cr < KD; ++cr) jv[i][cr] = ja(cr, i);
        {
            double jt = 0.0;
#pragma unroll
            for (int cr = 0; cr < KD; ++cr) jt = fma(jv[0][cr], w[cr], jt);
            jt = duo.sum(jt);
            if (HAS_BASE) jt = fma(jb0, wb, jt);
            if (arm == 0) fused::put_joint(R, u_all_row, ctrl_row, 0, base_st - jt);
        }
#pragma unroll
        for (int i = 1; i < 7; ++i) {
            double jt = 0.0;
#pragma unroll
            for (int cr = 0; cr < KD; ++cr) jt = fma(jv[i][cr], w[cr], jt);
            fused::put_joint(R, u_all_row, ctrl_row, jb + i - 1, base_arm[i - 1] - jt);
        }
    }
    if (poison) {
        __syncwarp(duo.mask);                            // the partner's gripper outputs are in place
        if (arm == 0) {
            const double qnan = nan("");
            if (u_all_row)
                for (int j = 0; j < kN; ++j) u_all_row[j] = qnan;
            for (int c = 0; c < P.n_ctrl; ++c) ctrl_row[c] = qnan;
        }
    }
        if (status && arm == 0) *status = (uint8_t)flags;
    // ---- instances left to the warp: record (canonical rows), Jacobi eigen-solver, owner finishes
    unsigned todo = __ballot_sync(0xffffffffu, hard && valid && arm == 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        if ((lane & ~1) == src) {
            double *rec = wfix.rec;
            double vo[KD];
#pragma unroll
            for (int i = 0; i < KD; ++i) vo[i] = duo.other(S.v[i]);
            const int r0 = arm * KD, o0 = (arm ^ 1) * KD;
#pragma unroll
            for (int i = 0; i < KD; ++i) {
                const double vi = S.v[i] * S.inv0;
#pragma unroll
                for (int j = 0; j < KD; ++j) {
                    rec[RC::A + (r0 + i) * K + r0 + j] = fma(vi, S.v[j], S.D[i >= j ? ltri(i, j) : ltri(j, i)]);
                    rec[RC::A + (r0 + i) * K + o0 + j] = vi * vo[j];
                }
                if (HAS_BASE) {
                    rec[RC::A + (r0 + i) * K + 2 * KD] = vi * jb0;
                    rec[RC::A + (2 * KD) * K + r0 + i] = vi * jb0;
                }
                rec[RC::G + r0 + i] = gc[i];
                rec[RC::JST + r0 + i] = ja(i, 0);
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                rec[RC::BASE + 1 + 6 * arm + i] = base_arm[i];
#pragma unroll
                for (int cr = 0; cr < KD; ++cr)
                    rec[RC::JARM + (arm * 6 + i) * KD + cr] = ja(cr, i + 1);
            }
            if (arm == 0) {
                if (HAS_BASE) {
                    rec[RC::A + (2 * KD) * K + 2 * KD] = (jb0 * S.inv0) * jb0;
                    rec[RC::G + 2 * KD] = gcb;
                    rec[RC::JST + 2 * KD] = jb0;
                }
                rec[RC::BASE] = base_st;
                rec[RC::ABAD] = small_det ? 0.0 : 1.0;
                wfix.flags = 0;
            }
        }
        __syncwarp();
        tiled::eigen_solve<K, K>(reinterpret_cast<double (*)[K]>(wfix.rec + RC::A), wfix.Vs, wfix.rec + RC::G, wfix.w, wfix.cbuf,
                                 wfix.sbuf, wfix.rec[RC::ABAD] == 0.0, lane, &wfix.flags);
        __syncwarp();
        if (lane == src) {
            fused::fixup_finish<KD, HAS_BASE>(R, u_all_row, ctrl_row, wfix.rec, wfix.w, 0, 1);
            if (status) *status = (uint8_t)(*status | wfix.flags);
        }
        __syncwarp();
    }
}

// Original Jacobian entries of the lane's arm, re-read from the tile (L2).
struct lane_jrow_t {
    const double *p;
    __device__ __forceinline__ double operator()(int cr, int i) const { return lane::LdCached{}(p + (size_t)(cr * 7 + i) * kTile); }
};

// ---------------------------------------------------------------- kernel
// Packed ctrl rows of `n_valid` consecutive instances starting at inst0 (a multiple of 16): local array, peer-mapped
// gathered arrays or the NVSwitch multicast mapping (stream::write_ctrl_tile for a half tile).
__device__ __forceinline__ void write_ctrl_rows(const double *ctile, double *ctrl, const stream::Gather &G, int n_ctrl, int64_t inst0,
                                                int n_valid, int lane) {
    const int64_t row0 = inst0 * (int64_t)n_ctrl;
    if (n_valid == kHalf && G.ctrl_vec) {
        const double2 *src = reinterpret_cast<const double2 *>(ctile);
        double2 *dst = reinterpret_cast<double2 *>(ctrl + row0);
        for (int e = lane; e < (kHalf / 2) * n_ctrl; e += 32) {
            const double2 v = src[e];
            dst[e] = v;
            if (G.ctrl_mc)
                multimem_st(reinterpret_cast<double2 *>(G.ctrl_mc + G.gather_offset * n_ctrl + row0) + e, v);
            else
                for (int gi = 0; gi < G.n_gather; ++gi)
                    reinterpret_cast<double2 *>(G.ctrl_gather[gi] + G.gather_offset * n_ctrl + row0)[e] = v;
        }
    } else {
        for (int e = lane; e < n_valid * n_ctrl; e += 32) {
            const double v = ctile[e];
            ctrl[row0 + e] = v;
            if (G.ctrl_mc) multimem_st(G.ctrl_mc + G.gather_offset * n_ctrl + row0 + e, v);
            else
                for (int gi = 0; gi < G.n_gather; ++gi) G.ctrl_gather[gi][G.gather_offset * n_ctrl + row0 + e] = v;
        }
    }
}

template <int KD, bool HAS_BASE, int NT>
__global__ void __launch_bounds__(NT, 1)
osc_step_pair(const __grid_constant__ KParams P, const __grid_constant__ LaneArgs A, const int64_t B,
              const __grid_constant__ FRoles R, const __grid_constant__ stream::Gather G, const int warp_bytes) {
    using namespace stream;
    const int W = blockDim.x >> 5;                       // <= NT / 32: small batches are launched as more, smaller CTAs
    extern __shared__ __align__(16) unsigned char pair_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int arm = lane & 1, li = lane >> 1;
    unsigned char *wbase = pair_smem + (size_t)warp * warp_bytes;
    double *ctile = reinterpret_cast<double *>(wbase);                                     // [16][n_ctrl]
    fused::WarpFix<KD, HAS_BASE> &wfix =
        *reinterpret_cast<fused::WarpFix<KD, HAS_BASE> *>(wbase + ((kHalf * P.n_ctrl * 8 + 15) & ~15));
    const int64_t n_half = (B + kHalf - 1) / kHalf;
    const int64_t tile_doubles = (int64_t)A.n_entries * kTile;
    const int64_t gw = (int64_t)blockIdx.x * W + warp, gstride = (int64_t)gridDim.x * W;
    const double gb = P.use_g ? 1.0 : 0.0;
    const int jb = 1 + 12 * arm;
    const int D = P.D;
    PairSys<KD, HAS_BASE> S;
    S.duo.mask = 3u << (lane & ~1);
    S.duo.arm = arm;
    const Duo duo = S.duo;
    // Half tiles: the first one by position (neighbouring warps share a tile), the rest from the launch's ticket
    // counter when there is one - path lengths differ (pinv branch), tickets keep the warps busy to the end.  Every
    // warp draws exactly one ticket past the end; the last of those rewinds the counter for the next launch.
    const int64_t extra = n_half > gstride ? n_half - gstride : 0;
    auto next_half = [&](int64_t cur) -> int64_t {
        if (A.sched == nullptr) return cur + gstride;
        int t = 0;
        if (lane == 0) {
            t = atomicAdd(A.sched, 1);
            if ((int64_t)t == extra + gstride - 1) *A.sched = 0;
        }
        t = __shfl_sync(0xffffffffu, t, 0);
        return gstride + t;
    };
    for (int64_t ht = gw < n_half ? gw : next_half(gw); ht < n_half; ht = next_half(ht)) {
        const int64_t tile = ht >> 1;
        const int half = (int)(ht & 1);
        const double *tb = A.tiles + tile * tile_doubles;
        const double *tl = tb + half * kHalf + li;
        const int64_t inst = ht * kHalf + li;
        const bool valid = inst < B;
        const int64_t inst_c = valid ? inst : B - 1;
        const double *target_vel = A.target_vel ? A.target_vel + inst_c * D * 6 : nullptr;
        double *u_all_row = (A.u_all && valid) ? A.u_all + inst * kN : nullptr;
        double *ctrl_row = ctile + li * P.n_ctrl;
        // the half-0 warp pulls the tile towards L2 ahead of both (cp.async.bulk.prefetch.L2)
        if (A.pf > 0 && lane == 0 && half == 0) {
            lane::bulk_prefetch_l2(tb, (uint32_t)(A.gbase[1 + (A.pf < 5 ? A.pf : 5)] * kTile * 8));
            lane::bulk_prefetch_l2(tb + (size_t)A.gbase[6] * kTile, (uint32_t)((A.gbase[6 + (A.pf < 5 ? A.pf : 5)] - A.gbase[6]) * kTile * 8));
        }
        auto reader = [tl](int first_entry) {
            const double *p = tl + (size_t)first_entry * kTile;
            return [p](int e) { return lane::LdStream{}(p + e * kTile); };
        };
        auto arm_group = [&](int j) {                  // group j = 0..4 of the lane's arm
            const int g = 1 + 5 * arm + j;
            if (A.pf > 0 && li == 0 && half == 0 && j + A.pf < 5)
                lane::bulk_prefetch_l2(tb + (size_t)A.gbase[g + A.pf] * kTile, (uint32_t)((A.gbase[g + A.pf + 1] - A.gbase[g + A.pf]) * kTile * 8));
            return reader(A.gbase[g]);
        };
        unsigned vel_zero = 0;
#pragma unroll
        for (int d = 0; d < IRLOSC_MAX_DEVICES; ++d) {
            bool tracking = false;
            if (d < D && target_vel != nullptr) {
                tracking = true;
                for (int i = 0; i < 6; ++i) tracking = tracking && (target_vel[d * 6 + i] != 0.0);
            }
            if (!tracking) vel_zero |= 1u << d;
        }
        int flags = 0;
        // ---- G0: stand / base device (both lanes)
        double bias0, jb0 = 0.0, gcb = 0.0, dxb = 0.0;
        {
            auto rd = reader(A.gbase[0]);
            bias0 = rd(kG0Bias0);
            if (HAS_BASE) {
                jb0 = rd(kG0Jbase);
                double u6[6];
                device_u6_staged(P, R.dev_base, rd, kG0Dev, true, u6);
                const KDevice &dv = P.dev[R.dev_base];
#pragma unroll
                for (int i = 0; i < 6; ++i)
                    if (dv.dof[i]) gcb = u6[i];
            }
        }
        // ---- the lane's arm
        ArmState<KD> Sa;
        {
            auto rd = arm_group(0);
            consume_cc<KD>(rd, arm, Sa);
        }
        if (HAS_BASE) dxb = jb0 * Sa.dqC[0];
        bool m_ok = true;
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
            auto rd = arm_group(1 + hf);
            m_ok = consume_grip<KD>(rd, P, R, vel_zero, gb, jb + 6 + 3 * hf, Sa, u_all_row, ctrl_row, nullptr) && m_ok;
        }
        double dxa[KD], gc[KD], base_arm[6], c0, uv0;
        {
            auto rd = arm_group(3);
            m_ok = consume_rows<KD>(rd, P, R, vel_zero, gb, jb, Sa, S.D, S.v, (double *)nullptr, dxa, (double (*)[KD]) nullptr,
                                    base_arm, &c0, &uv0, nullptr) && m_ok;
        }
        {
            auto rd = arm_group(4);
            double u6[6];
            device_u6_staged(P, R.dev_arm[arm], rd, 0, true, u6);
            const KDevice &dv = P.dev[R.dev_arm[arm]];
            int r = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i)
                if (dv.dof[i]) {
#pragma unroll
                    for (int s = 0; s < KD; ++s)
                        if (r == s) gc[s] = u6[i];
                    ++r;
                }
        }
        // ---- the stand joint couples the arms
        const double d0 = duo.sum(c0), uv_st = duo.sum(uv0);
        m_ok = duo.all(m_ok) && (d0 > 0.0);
        S.d0 = d0;
        S.inv0 = rcp64(d0);
        S.vb = jb0;
        const double base_st = fma(fused::coef_uv(P, R, vel_zero, 0), uv_st, gb * bias0);
        const lane_jrow_t jrow{tl + (size_t)A.gbase[4 + 5 * arm] * kTile};       // the arm's task rows: [cr][stand, joints 1..6]
        pair_tail<KD, HAS_BASE>(P, R, S, gc, gcb, dxa, dxb, base_arm, base_st, m_ok, jrow, target_vel, vel_zero, flags, u_all_row,
                                ctrl_row, (A.status && valid) ? A.status + inst : nullptr, valid, wfix, lane);
        __syncwarp();
        {
            const int64_t left = B - ht * kHalf;
            write_ctrl_rows(ctile, A.ctrl, G, P.n_ctrl, ht * kHalf, (int)(left < kHalf ? left : kHalf), lane);
        }
        __syncwarp();
    }
}

}  // namespace pair
}  // namespace irlosc
#endif

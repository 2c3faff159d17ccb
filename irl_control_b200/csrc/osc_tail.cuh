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
    if (small_det && !a_bad) {
        // X = L^-1 (unit lower triangular), column by column
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
#include "osc_fixup_coop.cuh"

namespace irlosc {
namespace fused {

// One warp per queued instance: eigen-decomposition of A (tiled::eigen_solve), w = pinv(A) g, then the
// joints with Jacobian columns are rewritten (and re-sent to the gather targets, if any).
template <int KD, bool HAS_BASE>
__global__ void __launch_bounds__(128, 1)
osc_tail_fixup(const __grid_constant__ KParams P, const __grid_constant__ TailOut out, const __grid_constant__ FRoles R,
               const __grid_constant__ HardQueue hq, const int coop) {
    using RC = Rec<KD, HAS_BASE>;
    constexpr int K = RC::K;
    struct WarpSmem {
        double As[K][K + 1], Vs[K][K + 1];
        double g[K], w[K], cbuf[32], sbuf[32];
        int flags;
    };
    __shared__ WarpSmem sm[4];
    __shared__ CoopSmem<K> csm[4];
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
        int how = 0;
        if (coop && !a_bad) {                    // experimental: bounds + deflation instead of the Jacobi sweeps
            CoopSmem<K> &CS = csm[warp];
            for (int e = lane; e < K * K; e += 32) CS.A[e / K][e % K] = S.As[e / K][e % K];
            if (lane < K) CS.g[lane] = S.g[lane];
            __syncwarp();
            CoopDevEx ex{lane};
            how = coop_resolve_pinv<K>(CS, ex);
            if (how != 0) {
                if (lane < K) S.w[lane] = CS.w[lane];
                if (lane == 0) S.flags = IRLOSC_ST_PINV | (how == 2 ? IRLOSC_ST_EIGEN : 0);
            }
            __syncwarp();
        }
        if (how == 0) tiled::eigen_solve<K>(S.As, S.Vs, S.g, S.w, S.cbuf, S.sbuf, !a_bad, lane, &S.flags);
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

// Generic OSC step kernel: any n <= 32, k <= 24, any input layout / stride.
// One warp per robot instance, working set in shared memory.  This is the
// correctness-first kernel and the fallback for shapes the register-tiled
// DualUR5 kernel (osc_tiled.cuh) is not instantiated for.
//
// Control law restated (ir-lab/irl_control):
//   osc.py:41-68   Mx from J M^-1 J^T      -> Cholesky(M), Y = L^-1 J^T, A = Y^T Y,
//                                             det(A) from chol(A); inverse or pinv(rcond 1e-5)
//   osc.py:150-152 dx = J dq, uv = M dq
//   osc.py:156-181 per-device task signal  -> device_task_signal (irlosc_device.cuh)
//   osc.py:184-200 u = -J^T Mx f + bias + (I - J^T Jbar^T) M (-kv_n dq)
//                  fused as  u = u_vel + bias - kv_n uv - J^T Mx (f - kv_n dx)
//                  (Mx symmetric, M^-1 M = I)
//   osc.py:203-208 packing u_all[actuator_trnids] per target
#pragma once
#include "irlosc_device.cuh"
#include "osc_eigen.cuh"

namespace irlosc {

constexpr int kGenericWarps = 4;

struct GenericSmem {
    // leading dimensions are odd multiples of 8 bytes to spread banks
    double Ms[IRLOSC_MAX_N][IRLOSC_MAX_N + 1];   // M, then its Cholesky factor in the lower triangle
    double Js[IRLOSC_MAX_K][IRLOSC_MAX_N + 1];   // stacked task Jacobian (target order)
    double Ys[IRLOSC_MAX_K][IRLOSC_MAX_N + 1];   // rows: L^-1 J_r^T
    double As[IRLOSC_MAX_K][IRLOSC_MAX_K + 1];   // A = J M^-1 J^T, later eigen work matrix
    double Ls[IRLOSC_MAX_K][IRLOSC_MAX_K + 1];   // chol(A) / eigenvectors
    double dq[IRLOSC_MAX_N], uv[IRLOSC_MAX_N], u[IRLOSC_MAX_N];
    double dx[IRLOSC_MAX_K], g[IRLOSC_MAX_K], w[IRLOSC_MAX_K], t[IRLOSC_MAX_K];
    int vel_zero[IRLOSC_MAX_DEVICES];
    int flags;
};


// Cyclic Jacobi eigen-decomposition of the symmetric k x k matrix in S.As (destroyed:
// eigenvalues end on its diagonal); eigenvectors in the columns of S.Ls.  Warp-cooperative.
__device__ void jacobi_eigen(GenericSmem &S, int k, int lane) {
    for (int i = lane; i < k * k; i += 32) S.Ls[i / k][i % k] = (i / k == i % k) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = lane; i < k * k; i += 32) {
            const int r = i / k, c = i % k;
            const double v = S.As[r][c];
            if (r == c) dia += v * v; else off += v * v;
        }
        off = warp_sum(off);
        dia = warp_sum(dia);
        if (off <= 1e-28 * dia || off == 0.0) break;
        for (int p = 0; p < k - 1; ++p) {
            for (int q = p + 1; q < k; ++q) {
                const double apq = S.As[p][q];
                const double app = S.As[p][p], aqq = S.As[q][q];
                __syncwarp();
                if (fabs(apq) <= 1e-300 || apq * apq <= 1e-31 * fabs(app * aqq)) continue;   // warp-uniform (same smem values)
                const double theta = (aqq - app) / (2.0 * apq);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0);
                const double s = tt * c;
                if (lane < k) {
                    const int i = lane;
                    if (i != p && i != q) {
                        const double aip = S.As[i][p], aiq = S.As[i][q];
                        const double nip = c * aip - s * aiq, niq = s * aip + c * aiq;
                        S.As[i][p] = nip; S.As[p][i] = nip;
                        S.As[i][q] = niq; S.As[q][i] = niq;
                    }
                    const double vip = S.Ls[i][p], viq = S.Ls[i][q];
                    S.Ls[i][p] = c * vip - s * viq;
                    S.Ls[i][q] = s * vip + c * viq;
                }
                if (lane == 0) {
                    S.As[p][p] = app - tt * apq;
                    S.As[q][q] = aqq + tt * apq;
                    S.As[p][q] = 0.0;
                    S.As[q][p] = 0.0;
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kGenericWarps * 32)
osc_step_generic(const KParams P, const KIo io, const int64_t B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GenericSmem &S = reinterpret_cast<GenericSmem *>(smem_raw)[warp];
    const int n = P.n, k = P.k, D = P.D;

    for (int64_t b = (int64_t)blockIdx.x * kGenericWarps + warp; b < B;
         b += (int64_t)gridDim.x * kGenericWarps) {
        // ------------------------------------------------ load M, J, dq
        const double *Mg = io.M + b * io.m_stride;
        if (io.m_layout == IRLOSC_M_DENSE) {
            for (int e = lane; e < n * n; e += 32) {
                const int i = e / n, j = e % n;
                S.Ms[i][j] = Mg[(int64_t)i * io.ldm + j];
            }
        } else {
            for (int e = lane; e < n * (n + 1) / 2; e += 32) {
                // row-major lower triangle: invert e = i(i+1)/2 + j
                int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                while (i * (i + 1) / 2 > e) --i;
                while ((i + 1) * (i + 2) / 2 <= e) ++i;
                const int j = e - i * (i + 1) / 2;
                const double v = Mg[e];
                S.Ms[i][j] = v;
                S.Ms[j][i] = v;
            }
        }
        const double *Jg = io.J + b * io.j_stride;
        for (int e = lane; e < k * n; e += 32) {
            const int r = e / n, j = e % n;
            int64_t src_row = r;
            if (io.j_layout == IRLOSC_J_FULL6) src_row = (int64_t)P.row_dev[r] * 6 + P.row_comp[r];
            S.Js[r][j] = Jg[src_row * io.ldj + j];
        }
        if (lane < n) S.dq[lane] = io.dq[b * n + lane];
        if (lane == 0) S.flags = 0;
        __syncwarp();

        // ------------------------------------------------ uv = M dq, dx = J dq (osc.py:150-151)
        if (lane < n) {
            double acc = 0.0;
            for (int j = 0; j < n; ++j) acc += S.Ms[lane][j] * S.dq[j];
            S.uv[lane] = acc;
        }
        if (lane < k) {
            double acc = 0.0;
            for (int j = 0; j < n; ++j) acc += S.Js[lane][j] * S.dq[j];
            S.dx[lane] = acc;
        }
        __syncwarp();

        // ------------------------------------------------ per-device task signal (osc.py:156-181)
        if (lane < D) {
            const int d = lane;
            const KDevice &dv = P.dev[d];
            const int64_t bd = b * D + d;
            double mv[2] = {dv.max_vel[0], dv.max_vel[1]};
            if (io.max_vel) { mv[0] = io.max_vel[bd * 2]; mv[1] = io.max_vel[bd * 2 + 1]; }
            double ee[3], eq[4], tx[3], tq[4], tv[6], u6[6];
            for (int i = 0; i < 3; ++i) { ee[i] = io.ee_xyz[bd * 3 + i]; tx[i] = io.target_xyz[bd * 3 + i]; }
            for (int i = 0; i < 4; ++i) { eq[i] = io.ee_quat[bd * 4 + i]; tq[i] = io.target_quat[bd * 4 + i]; }
            if (io.target_vel) for (int i = 0; i < 6; ++i) tv[i] = io.target_vel[bd * 6 + i];
            bool oob = false;
            const bool tracking = device_task_signal(dv, ee, eq, tx, tq, io.target_vel ? tv : nullptr,
                                                     mv, S.dx, k, u6, &oob);
            S.vel_zero[d] = tracking ? 0 : 1;
            double ft[6] = {0, 0, 0, 0, 0, 0};
            if (P.admittance) {
                double R[9], raw[6];
                for (int i = 0; i < 9; ++i) R[i] = io.ft_xmat[bd * 9 + i];
                for (int i = 0; i < 6; ++i) raw[i] = io.ft_raw[bd * 6 + i];
                rotate_wrench(R, raw, ft);
            }
            int r = dv.row0;
            for (int i = 0; i < 6; ++i)
                if (dv.dof[i]) {
                    double v = u6[i];
                    if (P.admittance) v += ft[i];          // osc.py:185
                    S.g[r++] = v;
                }
            int fl = 0;
            if (tracking) fl |= IRLOSC_ST_VEL_BRANCH;
            if (oob) fl |= IRLOSC_ST_DX_RANGE;
            if (fl) atomicOr(&S.flags, fl);
        }
        __syncwarp();
        if (P.has_nullspace && lane < k) S.g[lane] -= P.nullspace_kv * S.dx[lane];

        // ------------------------------------------------ Cholesky of M (lower, in place)
        bool m_bad = false;
        for (int p = 0; p < n; ++p) {
            const double dpp = S.Ms[p][p];
            if (!(dpp > 0.0)) m_bad = true;
            const double inv = 1.0 / sqrt(dpp);
            __syncwarp();
            if (lane == p) S.Ms[p][p] = dpp * inv;
            if (lane > p && lane < n) S.Ms[lane][p] *= inv;
            __syncwarp();
            if (lane > p && lane < n) {
                const double lip = S.Ms[lane][p];
                for (int j = p + 1; j <= lane; ++j) S.Ms[lane][j] -= lip * S.Ms[j][p];
            }
            __syncwarp();
        }
        // ------------------------------------------------ Y rows: L y_r = J_r^T (lane r)
        if (lane < k) {
            for (int i = 0; i < n; ++i) {
                double acc = S.Js[lane][i];
                for (int p = 0; p < i; ++p) acc -= S.Ms[i][p] * S.Ys[lane][p];
                S.Ys[lane][i] = acc / S.Ms[i][i];
            }
        }
        __syncwarp();
        // ------------------------------------------------ A = Y Y^T  (= J M^-1 J^T, osc.py:50)
        for (int e = lane; e < k * k; e += 32) {
            const int r = e / k, c = e % k;
            if (c <= r) {
                double acc = 0.0;
                for (int i = 0; i < n; ++i) acc += S.Ys[r][i] * S.Ys[c][i];
                S.As[r][c] = acc;
                S.As[c][r] = acc;
                S.Ls[r][c] = acc;
            }
        }
        __syncwarp();
        // ------------------------------------------------ chol(A) -> det, solve
        bool a_bad = false;
        double det = 1.0;
        for (int p = 0; p < k; ++p) {
            const double dpp = S.Ls[p][p];
            if (!(dpp > 0.0)) a_bad = true;
            det *= dpp;
            const double inv = 1.0 / sqrt(dpp);
            __syncwarp();
            if (lane == p) S.Ls[p][p] = dpp * inv;
            if (lane > p && lane < k) S.Ls[lane][p] *= inv;
            __syncwarp();
            if (lane > p && lane < k) {
                const double lip = S.Ls[lane][p];
                for (int j = p + 1; j <= lane; ++j) S.Ls[lane][j] -= lip * S.Ls[j][p];
            }
            __syncwarp();
        }
        int fl = 0;
        const bool use_pinv = !a_bad && !(fabs(det) >= kDetThreshold);
        if (!a_bad && !use_pinv) {
            // exact inverse branch (osc.py:53): w = A^-1 g through the Cholesky factor
            if (lane == 0) {
                for (int i = 0; i < k; ++i) {
                    double acc = S.g[i];
                    for (int p = 0; p < i; ++p) acc -= S.Ls[i][p] * S.t[p];
                    S.t[i] = acc / S.Ls[i][i];
                }
                for (int i = k - 1; i >= 0; --i) {
                    double acc = S.t[i];
                    for (int p = i + 1; p < k; ++p) acc -= S.Ls[p][i] * S.w[p];
                    S.w[i] = acc / S.Ls[i][i];
                }
            }
        } else {
            // pinv branch (osc.py:55) or numerically indefinite A: symmetric eigen-decomposition
            jacobi_eigen(S, k, lane);
            fl |= IRLOSC_ST_EIGEN;
            double lmax = 0.0, dete = 1.0;
            for (int i = 0; i < k; ++i) { lmax = fmax(lmax, fabs(S.As[i][i])); dete *= S.As[i][i]; }
            const bool pinv = a_bad ? !(fabs(dete) >= kDetThreshold) : true;
            if (pinv) fl |= IRLOSC_ST_PINV;
            if (lane < k) {
                const double lam = S.As[lane][lane];
                double proj = 0.0;
                for (int i = 0; i < k; ++i) proj += S.Ls[i][lane] * S.g[i];
                const bool keep = pinv ? (fabs(lam) > kPinvRcond * lmax) : true;
                S.t[lane] = keep ? proj / lam : 0.0;
            }
            __syncwarp();
            if (lane < k) {
                double acc = 0.0;
                for (int c = 0; c < k; ++c) acc += S.Ls[lane][c] * S.t[c];
                S.w[lane] = acc;
            }
        }
        __syncwarp();

        // ------------------------------------------------ joint-space assembly (osc.py:174,184-200)
        if (lane < n) {
            const int j = lane;
            double u = 0.0;
            for (int d = 0; d < D; ++d)
                if (S.vel_zero[d] && ((P.dev[d].joint_mask >> j) & 1u)) u = -1.0 * P.dev[d].kv * S.uv[j];
            double jt = 0.0;
            for (int r = 0; r < k; ++r) jt += S.Js[r][j] * S.w[r];
            u -= jt;
            if (P.use_g) u += io.bias[b * n + j];
            if (P.has_nullspace) u -= P.nullspace_kv * S.uv[j];
            if (m_bad || (S.flags & IRLOSC_ST_DX_RANGE)) u = nan("");
            S.u[j] = u;
            if (io.u_all) io.u_all[b * n + j] = u;
        }
        __syncwarp();
        // ------------------------------------------------ packing (osc.py:203-208)
        if (lane < P.n_ctrl) {
            int d = 0;
            while (d + 1 < D && lane >= P.dev[d + 1].ctrl0) ++d;
            store_ctrl(io, P.n_ctrl, b, lane, S.u[P.dev[d].actuator[lane - P.dev[d].ctrl0]]);
        }
        if (io.status && lane == 0)
            io.status[b] = (uint8_t)(S.flags | fl | (m_bad ? IRLOSC_ST_M_NOT_PD : 0));
        __syncwarp();
    }
}

}  // namespace irlosc

// Row-distributed, shuffle-broadcast DualUR5 OSC step kernel (sm_100a).
//
// Same augmented elimination as osc_tiled.cuh (see the header there), but the matrix is
// distributed by ROWS over the G lanes of an instance group and the pivot column is
// broadcast with warp shuffles straight out of the owners' registers:
//
//   lane l owns rows  i = l, l+G, l+2G, ...  of  S = [[M, J^T], [J, 0]]  (lower triangle,
//   row i keeps columns 0..i) plus one "extra" row: the right-hand side (-g) or one row of
//   the K x K identity block.  At pivot p every lane needs, for its own rows i > p,
//       a[i][j] -= a[i][p] * (1/d_p) * a[j][p]            for p < j <= i,
//   i.e. its OWN column-p entries (local) and the column-p entries a[j][p] of the other
//   rows.  a[j][p] lives in lane j % G at a compile-time register index, so it is fetched
//   with __shfl_sync(reg, j % G, G): no shared-memory scratch, no barrier, no lane-dependent
//   addressing.  (ncu on the column/scratch version showed the shared-memory pipe at ~70 %
//   of its wavefront rate and FP64 at 21 %: profiles/r01_*.)
//
//   The extra rows make the solve fall out of the elimination itself: with b = -g as row
//   NT, the multiplier of the rhs row at pivot N+c is z_c = (D^-1 L^-1 b)_c, and the
//   identity row r holds row r of L^-T, so  w_r = sum_c (L^-1)_{c r} z_c  accumulates
//   locally - no back substitution.  trace(A^-1) for the pinv certificate accumulates the
//   same way.
//
// Reference restated: ir-lab/irl_control osc.py:41-68 (Mx), 150-152, 156-181, 184-210.
#pragma once
#include "osc_tiled.cuh"

namespace irlosc {
namespace rows {

using tiled::sfor;
using tiled::rcp_nr;
using tiled::kWarpsPerCta;

template <int N, int K, int G>
struct RowTile {
    static constexpr int NT = N + K;
    static constexpr int WI = 32 / G;
    static constexpr int RB = (NT + G - 1) / G;      // row blocks (rows per lane)
    static constexpr int NX = K + 1;                 // extra rows: rhs + K identity rows
    static constexpr int XS = (NX + G - 1) / G;      // extra rows per lane
    __host__ __device__ static constexpr int rl(int m) { return G * (m + 1) < NT ? G * (m + 1) : NT; }
    __host__ __device__ static constexpr int roff(int m) {
        int o = 0;
        for (int i = 0; i < m; ++i) o += rl(i);
        return o;
    }
    static constexpr int TOT = roff(RB);
};

template <int N, int K, int D, int G, bool PACKED>
struct RowSmem {
    static constexpr int WI = 32 / G;
    static constexpr int MSZ = PACKED ? N * (N + 1) / 2 : N * N;
    alignas(16) double M[WI * MSZ];
    alignas(16) double J[WI * K * N];
    alignas(16) double dq[WI * N];
    alignas(16) double bias[WI * N];
    alignas(16) double ee_xyz[WI * 3 * D];
    alignas(16) double ee_quat[WI * 4 * D];
    alignas(16) double t_xyz[WI * 3 * D];
    alignas(16) double t_quat[WI * 4 * D];
    alignas(16) double t_vel[WI * 6 * D];
    alignas(16) double max_vel[WI * 2 * D];
    alignas(16) double ft_xmat[WI * 9 * D];
    alignas(16) double ft_raw[WI * 6 * D];
    alignas(16) double w[WI][(K + 1) & ~1];
    double As[WI][K][K + 1];
    double Vs[K][K + 1];
    double uv[WI][N], dx[WI][K], g[WI][K], u[WI][N];
    int vel_zero[WI][D];
    int flags[WI];
    alignas(8) unsigned long long bar_m;
    alignas(8) unsigned long long bar_v;
};

template <int N, int K, int D, int G, bool PACKED, int MINB>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MINB)
osc_step_rows(const KParams P, const KIo io, const int64_t B) {
    using T = RowTile<N, K, G>;
    using WS = RowSmem<N, K, D, G, PACKED>;
    constexpr int NT = T::NT, WI = T::WI, RB = T::RB, XS = T::XS, NX = T::NX, MSZ = WS::MSZ;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, l = lane % G;
    WS &S = reinterpret_cast<WS *>(smem_raw)[warp];

    if (lane == 0) {
        tiled::mbar_init(&S.bar_m, 1);
        tiled::mbar_init(&S.bar_v, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int64_t n_tiles = (B + WI - 1) / WI;
    const int64_t warp_global = (int64_t)blockIdx.x * kWarpsPerCta + warp;
    const int64_t warp_stride = (int64_t)gridDim.x * kWarpsPerCta;
    const bool has_tvel = io.target_vel != nullptr, has_mvel = io.max_vel != nullptr;
    const bool adm = P.admittance != 0;
    uint32_t par_m = 0, par_v = 0;

    auto tile_full = [&](int64_t t) { return (t + 1) * WI <= B; };
    auto issue_M = [&](int64_t t) {
        tiled::mbar_expect_tx(&S.bar_m, WI * MSZ * 8);
        tiled::bulk_g2s(S.M, io.M + t * WI * (int64_t)MSZ, WI * MSZ * 8, &S.bar_m);
    };
    auto issue_V = [&](int64_t t) {
        uint32_t bytes = WI * 8 * (K * N + N + 14 * D);
        if (P.use_g) bytes += WI * 8 * N;
        if (has_tvel) bytes += WI * 8 * 6 * D;
        if (has_mvel) bytes += WI * 8 * 2 * D;
        if (adm) bytes += WI * 8 * 15 * D;
        tiled::mbar_expect_tx(&S.bar_v, bytes);
        const int64_t i0 = t * WI;
        tiled::bulk_g2s(S.J, io.J + i0 * (K * N), WI * K * N * 8, &S.bar_v);
        tiled::bulk_g2s(S.dq, io.dq + i0 * N, WI * N * 8, &S.bar_v);
        if (P.use_g) tiled::bulk_g2s(S.bias, io.bias + i0 * N, WI * N * 8, &S.bar_v);
        tiled::bulk_g2s(S.ee_xyz, io.ee_xyz + i0 * 3 * D, WI * 3 * D * 8, &S.bar_v);
        tiled::bulk_g2s(S.ee_quat, io.ee_quat + i0 * 4 * D, WI * 4 * D * 8, &S.bar_v);
        tiled::bulk_g2s(S.t_xyz, io.target_xyz + i0 * 3 * D, WI * 3 * D * 8, &S.bar_v);
        tiled::bulk_g2s(S.t_quat, io.target_quat + i0 * 4 * D, WI * 4 * D * 8, &S.bar_v);
        if (has_tvel) tiled::bulk_g2s(S.t_vel, io.target_vel + i0 * 6 * D, WI * 6 * D * 8, &S.bar_v);
        if (has_mvel) tiled::bulk_g2s(S.max_vel, io.max_vel + i0 * 2 * D, WI * 2 * D * 8, &S.bar_v);
        if (adm) {
            tiled::bulk_g2s(S.ft_xmat, io.ft_xmat + i0 * 9 * D, WI * 9 * D * 8, &S.bar_v);
            tiled::bulk_g2s(S.ft_raw, io.ft_raw + i0 * 6 * D, WI * 6 * D * 8, &S.bar_v);
        }
    };
    auto copy_rows = [&](double *dst, const double *src, int per, int64_t i0, int valid) {
        for (int e = lane; e < WI * per; e += 32) dst[e] = (e / per < valid) ? src[i0 * per + e] : 0.0;
    };
    auto manual_M = [&](int64_t t) {
        const int valid = (int)(B - t * WI);
        copy_rows(S.M, io.M, MSZ, t * WI, valid);
        for (int s = valid; s < WI; ++s)
            for (int i = lane; i < N; i += 32) S.M[s * MSZ + (PACKED ? i * (i + 1) / 2 + i : i * N + i)] = 1.0;
    };
    auto manual_V = [&](int64_t t) {
        const int64_t i0 = t * WI;
        const int valid = (int)(B - i0);
        copy_rows(S.J, io.J, K * N, i0, valid);
        copy_rows(S.dq, io.dq, N, i0, valid);
        if (P.use_g) copy_rows(S.bias, io.bias, N, i0, valid);
        copy_rows(S.ee_xyz, io.ee_xyz, 3 * D, i0, valid);
        copy_rows(S.ee_quat, io.ee_quat, 4 * D, i0, valid);
        copy_rows(S.t_xyz, io.target_xyz, 3 * D, i0, valid);
        copy_rows(S.t_quat, io.target_quat, 4 * D, i0, valid);
        if (has_tvel) copy_rows(S.t_vel, io.target_vel, 6 * D, i0, valid);
        if (has_mvel) copy_rows(S.max_vel, io.max_vel, 2 * D, i0, valid);
        if (adm) { copy_rows(S.ft_xmat, io.ft_xmat, 9 * D, i0, valid); copy_rows(S.ft_raw, io.ft_raw, 6 * D, i0, valid); }
        for (int s = valid; s < WI; ++s)
            for (int dd = lane; dd < D; dd += 32) { S.ee_quat[(s * D + dd) * 4] = 1.0; S.t_quat[(s * D + dd) * 4] = 1.0; }
    };

    if (warp_global < n_tiles && tile_full(warp_global) && lane == 0) { issue_M(warp_global); issue_V(warp_global); }

    for (int64_t tile = warp_global; tile < n_tiles; tile += warp_stride) {
        const bool full = tile_full(tile);
        const int64_t inst = tile * WI + grp;
        const bool valid = inst < B;
        if (full) {
            tiled::mbar_wait(&S.bar_v, par_v); par_v ^= 1;
            tiled::mbar_wait(&S.bar_m, par_m); par_m ^= 1;
        } else {
            manual_M(tile);
            manual_V(tile);
            __syncwarp();
        }
        const double *Ms = S.M + grp * MSZ;
        const double *Js = S.J + grp * K * N;
        const double *dqs = S.dq + grp * N;
        if (l == 0) S.flags[grp] = 0;

        // -------------------------------------------------- rows of S into registers
        // a[roff(m) + j] = S[i_m][j], i_m = l + G m, j < rl(m); entries right of the diagonal are 0
        double a[T::TOT];
        sfor<0, RB>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            const int i = l + G * m;
            const double *row = (i < N) ? (Ms + (PACKED ? i * (i + 1) / 2 : i * N)) : (Js + (i - N) * N);
            sfor<0, T::rl(m)>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                double v = 0.0;
                if constexpr (j < N) {
                    if ((i < N) ? (j <= i) : (i < NT)) v = row[j];
                }
                a[T::roff(m) + j] = v;
            });
        });

        // -------------------------------------------------- uv = M dq, dx = J dq from the registers
        // row part: sum_{j<=i} a[i][j] dq[j]  (for rows >= N this is dx);  column part of the
        // symmetric product: sum_{i>j} a[i][j] dq[i], reduced over the group with xor-shuffles.
        {
            double own_dq[RB], rowacc[RB], diag[RB];
            sfor<0, RB>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                const int i = l + G * m;
                own_dq[m] = (i < N) ? dqs[i] : 0.0;
                diag[m] = (i < N) ? Ms[PACKED ? i * (i + 1) / 2 + i : i * N + i] : 0.0;
                rowacc[m] = 0.0;
            });
            double colmine[RB];
            sfor<0, RB>([&](auto mc) { colmine[decltype(mc)::value] = 0.0; });
            sfor<0, N>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                const double dqj = dqs[j];
                double part = 0.0;
                sfor<j / G, RB>([&](auto mc) {
                    constexpr int m = decltype(mc)::value;
                    if constexpr (j < T::rl(m)) {
                        const double v = a[T::roff(m) + j];
                        rowacc[m] = fma(v, dqj, rowacc[m]);
                        part = fma(v, own_dq[m], part);      // includes the diagonal once, removed below
                    }
                });
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
                if (l == j % G) colmine[j / G] = part;
            });
            sfor<0, RB>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                const int i = l + G * m;
                if (i < N) S.uv[grp][i] = rowacc[m] + colmine[m] - diag[m] * own_dq[m];
                else if (i < NT) S.dx[grp][i - N] = rowacc[m];
            });
        }
        __syncwarp();
        // the M stage is free: pull the next tile's M while this one is processed
        const int64_t next = tile + warp_stride;
        const bool next_full = next < n_tiles && tile_full(next);
        if (next_full && lane == 0) { tiled::fence_proxy_async(); issue_M(next); }

        // -------------------------------------------------- per-device task signal (osc.py:156-181)
        if (l < D) {
            const int d = l;
            const KDevice &dv = P.dev[d];
            const int sd = grp * D + d;
            double mv[2] = {dv.max_vel[0], dv.max_vel[1]};
            if (has_mvel) { mv[0] = S.max_vel[sd * 2]; mv[1] = S.max_vel[sd * 2 + 1]; }
            double tv[6], u6[6];
            if (has_tvel)
                for (int i = 0; i < 6; ++i) tv[i] = S.t_vel[sd * 6 + i];
            bool oob = false;
            const bool tracking = device_task_signal(dv, &S.ee_xyz[sd * 3], &S.ee_quat[sd * 4], &S.t_xyz[sd * 3],
                                                     &S.t_quat[sd * 4], has_tvel ? tv : nullptr, mv, S.dx[grp], K,
                                                     u6, &oob);
            S.vel_zero[grp][d] = tracking ? 0 : 1;
            double ft[6] = {0, 0, 0, 0, 0, 0};
            if (adm) rotate_wrench(&S.ft_xmat[sd * 9], &S.ft_raw[sd * 6], ft);
            int r = dv.row0;
            const double kvn = P.has_nullspace ? P.nullspace_kv : 0.0;
            for (int i = 0; i < 6; ++i)
                if (dv.dof[i]) {
                    const double v = adm ? u6[i] + ft[i] : u6[i];
                    S.g[grp][r] = v - kvn * S.dx[grp][r];
                    ++r;
                }
            const int fl = (tracking ? IRLOSC_ST_VEL_BRANCH : 0) | (oob ? IRLOSC_ST_DX_RANGE : 0);
            if (fl) atomicOr(&S.flags[grp], fl);
        }
        __syncwarp();

        // -------------------------------------------------- extra rows: rhs (-g) and identity
        double x[XS][K];
        sfor<0, XS>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            const int e = l + G * s;
            sfor<0, K>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                x[s][c] = (e == 0) ? -S.g[grp][c] : ((e == c + 1) ? 1.0 : 0.0);
            });
        });

        // -------------------------------------------------- elimination
        double detinv = 1.0;
        bool m_bad = false, a_bad = false;
        double wacc[XS], tin[XS];
        sfor<0, XS>([&](auto sc) { wacc[decltype(sc)::value] = 0.0; tin[decltype(sc)::value] = 0.0; });
        double trA = 0.0;
        double invc = rcp_nr(a[0]);
        sfor<0, NT>([&](auto pc) {
            constexpr int p = decltype(pc)::value;
            constexpr int mp = p / G;
            constexpr int mlo = (p + 1) / G;
            const double inv = __shfl_sync(FULL, invc, p % G, G);
            if constexpr (p < N) { m_bad = m_bad || !(inv > 0.0); }
            else { detinv *= -inv; a_bad = a_bad || !(-inv > 0.0); }
            double mult[RB];
            sfor<mlo, RB>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                double t = a[T::roff(m) + p] * inv;
                if constexpr (m == mp) t = (l > p % G) ? t : 0.0;    // rows <= p of the pivot block are finished
                mult[m] = t;
            });
            double multx[XS];
            if constexpr (p >= N) {
                constexpr int c = p - N;
                sfor<0, XS>([&](auto sc) {
                    constexpr int s = decltype(sc)::value;
                    multx[s] = x[s][c] * inv;
                });
                // rhs row is extra row 0 (lane 0, slot 0): its multiplier is z_c = (D^-1 L^-1 b)_c
                const double z = __shfl_sync(FULL, multx[0], 0, G);
                sfor<0, XS>([&](auto sc) {
                    constexpr int s = decltype(sc)::value;
                    wacc[s] = fma(x[s][c], z, wacc[s]);             // identity row r: w_r += (L^-1)_{c r} z_c
                    tin[s] = fma(x[s][c] * x[s][c], -inv, tin[s]);  // trace(A^-1) += (L^-1)_{c r}^2 / d'_c
                });
            }
            sfor<p + 1, NT>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                constexpr int mj = j / G;
                const double xj = __shfl_sync(FULL, a[T::roff(mj) + p], j % G, G);
                sfor<(mj > mlo ? mj : mlo), RB>([&](auto mc) {
                    constexpr int m = decltype(mc)::value;
                    a[T::roff(m) + j] = fma(-xj, mult[m], a[T::roff(m) + j]);
                });
                if constexpr (p >= N) {
                    sfor<0, XS>([&](auto sc) {
                        constexpr int s = decltype(sc)::value;
                        x[s][j - N] = fma(-xj, multx[s], x[s][j - N]);
                    });
                }
                if constexpr (j == p + 1) invc = rcp_nr(a[T::roff(mj) + j]);   // next pivot, off the critical path
            });
            // A = -(Schur block), saved right before its first pivot; trace(A) on the fly
            if constexpr (p == N - 1) {
                sfor<N / G, RB>([&](auto mc) {
                    constexpr int m = decltype(mc)::value;
                    const int i = l + G * m;
                    if (i >= N && i < NT) {
                        sfor<N, T::rl(m)>([&](auto jc) {
                            constexpr int j = decltype(jc)::value;
                            if (j <= i) {
                                const double v = -a[T::roff(m) + j];
                                S.As[grp][i - N][j - N] = v;
                                S.As[grp][j - N][i - N] = v;
                                if (j == i) trA += v;
                            }
                        });
                    }
                });
            }
        });

        // -------------------------------------------------- certificate, w
        double trAinv = 0.0;
        sfor<0, XS>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            const int e = l + G * s;
            if (e >= 1 && e < NX) { trAinv += tin[s]; S.w[grp][e - 1] = wacc[s]; }
        });
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            trA += __shfl_xor_sync(FULL, trA, o);
            trAinv += __shfl_xor_sync(FULL, trAinv, o);
        }
        const bool small_det = !(fabs(detinv) <= 1.0 / kDetThreshold);   // |det A| < 1e-4 (osc.py:52)
        const bool certified = (trA * trAinv < 1.0 / kPinvRcond);
        const bool hard = a_bad || (small_det && !certified);
        if (l == 0) {
            int fl = 0;
            if (m_bad) fl |= IRLOSC_ST_M_NOT_PD;
            if (small_det && !a_bad) fl |= IRLOSC_ST_PINV;
            if (fl) S.flags[grp] |= fl;
        }
        __syncwarp();
        unsigned hard_mask = __ballot_sync(FULL, hard && valid && (l == 0));
        while (hard_mask) {
            const int src = __ffs(hard_mask) - 1;
            hard_mask &= hard_mask - 1;
            const int gi = src / G;
            const bool gi_abad = __shfl_sync(FULL, a_bad ? 1 : 0, src) != 0;
            tiled::eigen_solve<K>(S.As[gi], S.Vs, S.g[gi], S.w[gi], S.u[gi], S.dx[gi], !gi_abad, lane, &S.flags[gi]);
        }
        __syncwarp();

        // -------------------------------------------------- joint-space assembly (osc.py:174,184-200)
        const int flg = S.flags[grp];
        const bool poison = (flg & (IRLOSC_ST_M_NOT_PD | IRLOSC_ST_DX_RANGE)) != 0;
#pragma unroll
        for (int t = 0; t < (N + G - 1) / G; ++t) {
            const int j = l + G * t;
            if (j < N) {
                const double uvj = S.uv[grp][j];
                double u = 0.0;
#pragma unroll
                for (int d = 0; d < D; ++d)
                    if (S.vel_zero[grp][d] && ((P.dev[d].joint_mask >> j) & 1u)) u = -1.0 * P.dev[d].kv * uvj;
                double jt = 0.0;
#pragma unroll
                for (int c = 0; c < K; ++c) jt = fma(Js[c * N + j], S.w[grp][c], jt);
                u -= jt;
                if (P.use_g) u += S.bias[grp * N + j];
                if (P.has_nullspace) u -= P.nullspace_kv * uvj;
                if (poison) u = nan("");
                S.u[grp][j] = u;
                if (io.u_all && valid) io.u_all[inst * N + j] = u;
            }
        }
        __syncwarp();
        // -------------------------------------------------- packing (osc.py:203-208)
#pragma unroll
        for (int t = 0; t < (32 + G - 1) / G; ++t) {
            const int c = l + G * t;
            if (c < P.n_ctrl && valid) {
                int d = 0;
                while (d + 1 < D && c >= P.dev[d + 1].ctrl0) ++d;
                store_ctrl(io, P.n_ctrl, inst, c, S.u[grp][P.dev[d].actuator[c - P.dev[d].ctrl0]]);
            }
        }
        if (io.status && valid && l == 0) io.status[inst] = (uint8_t)flg;
        __syncwarp();
        if (next_full && lane == 0) { tiled::fence_proxy_async(); issue_V(next); }
    }
}

}  // namespace rows
}  // namespace irlosc

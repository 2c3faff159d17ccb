// Register-tiled DualUR5 OSC step kernel (sm_100a).
//
// One robot instance is handled by a group of G lanes (G = 8 or 16), so a warp works on
// WI = 32/G instances at a time and a persistent grid of warps walks over the batch.
//
// Math (same control law as osc_generic.cuh / ir-lab/irl_control osc.py:41-68,150-210):
// the whole chain  M^-1 -> A = J M^-1 J^T -> Mx = A^-1 -> w = Mx g  is ONE symmetric
// elimination of the augmented quasi-definite matrix
//
//          [ M   J^T ]   N rows            after the first N pivots the trailing K x K block
//      S = [ J    0  ]   K rows            is the Schur complement  -J M^-1 J^T = -A;  the next
//          [ 0  -g^T ]   right-hand side   K pivots factor -A = L D L^T (D < 0) and carry the
//          [ 0    I  ]   K identity rows   rhs row to L^-1(-g) and the identity rows to L^-1.
//
// No square roots, one reciprocal per pivot.  det(A) = prod(-d_p) decides the reference's
// branch (osc.py:52): |det| >= 1e-4 -> exact inverse = the LDL^T solve.  Otherwise the
// reference takes pinv(rcond=1e-5), which equals the inverse whenever no eigenvalue falls
// below 1e-5 * lambda_max; that is certified cheaply by  trace(A) * trace(A^-1) < 1e5
// (trace(A^-1) from the identity rows).  Only uncertified instances go through the
// warp-cooperative Jacobi eigen-solver on a saved copy of A.
//
// Data layout: S is distributed column-cyclically over the G lanes (lane l owns columns
// l, l+G, ...), lower triangle only, entirely in registers; every index below is a
// compile-time constant after unrolling.  Per pivot the owner lane publishes its column
// through a double-buffered shared-memory scratch line (one __syncwarp per pivot) and all
// lanes read it back with broadcast 128-bit loads.
//
// HBM -> SM: each warp pulls its tile (WI consecutive instances of every input array) with
// 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) into its private shared
// memory stage; the M region is re-armed for the next tile as soon as the columns are in
// registers, so the next tile's largest transfer overlaps the elimination.
#pragma once
#include <type_traits>
#include "irlosc_device.cuh"
#include "osc_generic.cuh"
#include "osc_eigen.cuh"

namespace irlosc {
namespace tiled {

constexpr int kWarpsPerCta = 4;

template <int I, int E, typename F>
__device__ __forceinline__ void sfor(F &&f) {
    if constexpr (I < E) {
        f(std::integral_constant<int, I>{});
        sfor<I + 1, E>(f);
    }
}

// 1/d to ~1 ulp: hardware seed (MUFU.RCP64H) + two Newton steps.
__device__ __forceinline__ double rcp_nr(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// ---- TMA / mbarrier PTX ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- static description of the register tile ---------------------------------------------
template <int N, int K, int G>
struct Tile {
    static constexpr int NT = N + K;                 // pivots
    static constexpr int WI = 32 / G;                // instances per warp
    static constexpr int MC = (NT + G - 1) / G;      // column blocks (columns per lane)
    static constexpr int MK0 = N / G;                // first block that holds a K-column (j >= N)
    static constexpr int RHS = NT;                   // absolute row of the right-hand side
    static constexpr int XR = NT + 1 + K;            // rows incl. rhs and identity rows
    static constexpr int SCR = ((XR + 2) + 1) & ~1;  // scratch line: XR rows + 1/pivot, even length
    static constexpr int INV = XR;                   // slot of 1/pivot in the scratch line
    __host__ __device__ static constexpr int rows_end(int m) { return m >= MK0 ? XR : NT; }
    __host__ __device__ static constexpr int blk_len(int m) { return rows_end(m) - G * m; }
    __host__ __device__ static constexpr int blk_off(int m) {
        int o = 0;
        for (int i = 0; i < m; ++i) o += blk_len(i);
        return o;
    }
    static constexpr int TOT = blk_off(MC);
    // rows a consumer of pivot p touches: (p, hi(p))
    __host__ __device__ static constexpr int hi(int p) { return p < N ? NT : NT + 2 + (p - N); }
};

template <int N, int K, int D, int G, bool PACKED>
struct WarpSmem {
    using T = Tile<N, K, G>;
    static constexpr int WI = T::WI;
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
    alignas(16) double scr[WI][2][T::SCR];
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
osc_step_tiled(const KParams P, const KIo io, const int64_t B) {
    using T = Tile<N, K, G>;
    using WS = WarpSmem<N, K, D, G, PACKED>;
    constexpr int NT = T::NT, WI = T::WI, MC = T::MC, MSZ = WS::MSZ;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, l = lane % G;          // instance slot in the warp, lane in the group
    WS &S = reinterpret_cast<WS *>(smem_raw)[warp];
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));

    if (lane == 0) {
        mbar_init(&S.bar_m, 1);
        mbar_init(&S.bar_v, 1);
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
    // lane 0: arm + issue the bulk copies of one tile
    auto issue_M = [&](int64_t t) {
        mbar_expect_tx(&S.bar_m, WI * MSZ * 8);
        bulk_g2s(S.M, io.M + t * WI * (int64_t)MSZ, WI * MSZ * 8, &S.bar_m);
    };
    auto issue_V = [&](int64_t t) {
        uint32_t bytes = WI * 8 * (K * N + N + 14 * D);
        if (P.use_g) bytes += WI * 8 * N;
        if (has_tvel) bytes += WI * 8 * 6 * D;
        if (has_mvel) bytes += WI * 8 * 2 * D;
        if (adm) bytes += WI * 8 * 15 * D;
        mbar_expect_tx(&S.bar_v, bytes);
        const int64_t i0 = t * WI;
        bulk_g2s(S.J, io.J + i0 * (K * N), WI * K * N * 8, &S.bar_v);
        bulk_g2s(S.dq, io.dq + i0 * N, WI * N * 8, &S.bar_v);
        if (P.use_g) bulk_g2s(S.bias, io.bias + i0 * N, WI * N * 8, &S.bar_v);
        bulk_g2s(S.ee_xyz, io.ee_xyz + i0 * 3 * D, WI * 3 * D * 8, &S.bar_v);
        bulk_g2s(S.ee_quat, io.ee_quat + i0 * 4 * D, WI * 4 * D * 8, &S.bar_v);
        bulk_g2s(S.t_xyz, io.target_xyz + i0 * 3 * D, WI * 3 * D * 8, &S.bar_v);
        bulk_g2s(S.t_quat, io.target_quat + i0 * 4 * D, WI * 4 * D * 8, &S.bar_v);
        if (has_tvel) bulk_g2s(S.t_vel, io.target_vel + i0 * 6 * D, WI * 6 * D * 8, &S.bar_v);
        if (has_mvel) bulk_g2s(S.max_vel, io.max_vel + i0 * 2 * D, WI * 2 * D * 8, &S.bar_v);
        if (adm) {
            bulk_g2s(S.ft_xmat, io.ft_xmat + i0 * 9 * D, WI * 9 * D * 8, &S.bar_v);
            bulk_g2s(S.ft_raw, io.ft_raw + i0 * 6 * D, WI * 6 * D * 8, &S.bar_v);
        }
    };
    // ragged last tile: plain loads of the valid instances, zero-fill the rest
    auto copy_rows = [&](double *dst, const double *src, int per, int64_t i0, int valid) {
        for (int e = lane; e < WI * per; e += 32) dst[e] = (e / per < valid) ? src[i0 * per + e] : 0.0;
    };
    auto manual_M = [&](int64_t t) {
        const int valid = (int)(B - t * WI);
        copy_rows(S.M, io.M, MSZ, t * WI, valid);
        // keep the padded instances factorizable: identity diagonal
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

    // prologue: first tile of this warp
    if (warp_global < n_tiles) {
        if (tile_full(warp_global)) {
            if (lane == 0) { issue_M(warp_global); issue_V(warp_global); }
        }
    }

    for (int64_t tile = warp_global; tile < n_tiles; tile += warp_stride) {
        const bool full = tile_full(tile);
        const int64_t inst = tile * WI + grp;           // this group's instance
        const bool valid = inst < B;
        if (full) {
            mbar_wait(&S.bar_v, par_v); par_v ^= 1;
            mbar_wait(&S.bar_m, par_m); par_m ^= 1;
        } else {
            manual_M(tile);
            manual_V(tile);
            __syncwarp();
        }
        const double *Ms = S.M + grp * MSZ;
        const double *Js = S.J + grp * K * N;
        const double *dqs = S.dq + grp * N;
        if (l == 0) S.flags[grp] = 0;

        // -------------------------------------------------- uv = M dq, dx = J dq (osc.py:150-151)
#pragma unroll
        for (int t = 0; t < (N + G - 1) / G; ++t) {
            const int i = l + G * t;
            if (i < N) {
                const int ti = i * (i + 1) / 2;
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const int idx = PACKED ? ((j <= i) ? ti + j : j * (j + 1) / 2 + i) : i * N + j;
                    acc = fma(Ms[idx], dqs[j], acc);
                }
                S.uv[grp][i] = acc;
            }
        }
#pragma unroll
        for (int t = 0; t < (K + G - 1) / G; ++t) {
            const int c = l + G * t;
            if (c < K) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < N; ++j) acc = fma(Js[c * N + j], dqs[j], acc);
                S.dx[grp][c] = acc;
            }
        }
        __syncwarp();

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
            for (int i = 0; i < 6; ++i)
                if (dv.dof[i]) S.g[grp][r++] = adm ? u6[i] + ft[i] : u6[i];
            const int fl = (tracking ? IRLOSC_ST_VEL_BRANCH : 0) | (oob ? IRLOSC_ST_DX_RANGE : 0);
            if (fl) atomicOr(&S.flags[grp], fl);
        }
        __syncwarp();
        if (P.has_nullspace) {
#pragma unroll
            for (int t = 0; t < (K + G - 1) / G; ++t) {
                const int c = l + G * t;
                if (c < K) S.g[grp][c] -= P.nullspace_kv * S.dx[grp][c];
            }
        }
        __syncwarp();

        // -------------------------------------------------- columns of S into registers
        double a[T::TOT];
        sfor<0, MC>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            const int j = l + G * m;
            sfor<G * m, T::rows_end(m)>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                constexpr int slot = T::blk_off(m) + i - G * m;
                double v = 0.0;
                if constexpr (i < N) {                 // M block (lower triangle: i >= j)
                    if (j <= i && j < N) v = Ms[PACKED ? i * (i + 1) / 2 + j : i * N + j];
                } else if constexpr (i < NT) {         // J rows
                    if (j < N) v = Js[(i - N) * N + j];
                } else if constexpr (i == T::RHS) {    // rhs row: -g
                    if (j >= N && j < NT) v = -S.g[grp][j - N];
                } else {                               // identity rows
                    if (j == N + (i - NT - 1)) v = 1.0;
                }
                a[slot] = v;
            });
        });
        __syncwarp();
        // the M stage is free: pull the next tile's M while this one is eliminated
        const int64_t next = tile + warp_stride;
        const bool next_full = next < n_tiles && tile_full(next);
        if (next_full && lane == 0) { fence_proxy_async(); issue_M(next); }

        // -------------------------------------------------- elimination
        double *scr0 = &S.scr[grp][0][0];
        double *scr1 = &S.scr[grp][1][0];
        double det = 1.0;
        bool m_bad = false, a_bad = false;
        // publish column 0 (owner: lane 0 of the group)
        if (l == 0) {
            sfor<0, NT / 2 + 1>([&](auto ic) {
                constexpr int i = 2 * decltype(ic)::value;
                if constexpr (i < NT) {
                    double2 v;
                    v.x = a[T::blk_off(0) + i];
                    v.y = (i + 1 < NT) ? a[T::blk_off(0) + ((i + 1 < NT) ? i + 1 : i)] : 0.0;
                    *reinterpret_cast<double2 *>(scr0 + i) = v;
                }
            });
            scr0[T::INV] = rcp_nr(a[T::blk_off(0)]);
        }
        __syncwarp();

        sfor<0, NT>([&](auto pc) {
            constexpr int p = decltype(pc)::value;
            constexpr int hi = T::hi(p);
            constexpr int mlo = (p + 1) / G;              // first block with a live column
            double *cur = (p & 1) ? scr1 : scr0;
            double *nxt = (p & 1) ? scr0 : scr1;
            const double inv = cur[T::INV];
            const double dpv = cur[p];
            if constexpr (p < N) { m_bad = m_bad || !(dpv > 0.0); }
            else { det *= -dpv; a_bad = a_bad || !(-dpv > 0.0); }
            // multipliers of this lane's live columns
            double mult[MC];
            sfor<mlo, MC>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                const int j = l + G * m;
                const bool live = (j > p) && (j < NT);
                const double xj = cur[live ? j : p];
                mult[m] = live ? xj * inv : 0.0;
            });
            // rank-1 update, rows in pairs (broadcast 128-bit loads)
            constexpr int i_first = (p + 1) & ~1;
            sfor<0, (hi - i_first + 1) / 2>([&](auto qc) {
                constexpr int i0 = i_first + 2 * decltype(qc)::value;
                const double2 x = *reinterpret_cast<const double2 *>(cur + i0);
                sfor<0, 2>([&](auto ec) {
                    constexpr int i = i0 + decltype(ec)::value;
                    if constexpr (i > p && i < hi) {
                        const double xi = decltype(ec)::value ? x.y : x.x;
                        sfor<mlo, MC>([&](auto mc) {
                            constexpr int m = decltype(mc)::value;
                            if constexpr (G * m <= i && i < T::rows_end(m)) {
                                constexpr int slot = T::blk_off(m) + i - G * m;
                                a[slot] = fma(-xi, mult[m], a[slot]);
                            }
                        });
                    }
                });
            });
            // publish column p+1 for the next pivot
            if constexpr (p + 1 < NT) {
                constexpr int q = p + 1;
                constexpr int mq = q / G, lq = q % G;
                constexpr int hq = T::hi(q);
                if (l == lq) {
                    constexpr int j0 = q & ~1;
                    sfor<0, (hq - j0 + 1) / 2>([&](auto ic) {
                        constexpr int i = j0 + 2 * decltype(ic)::value;
                        constexpr int s0 = T::blk_off(mq) + i - G * mq;
                        double2 v;
                        v.x = (i >= G * mq && i < T::rows_end(mq)) ? a[(i >= G * mq && i < T::rows_end(mq)) ? s0 : T::blk_off(mq)] : 0.0;
                        v.y = (i + 1 < T::rows_end(mq)) ? a[(i + 1 < T::rows_end(mq)) ? s0 + 1 : T::blk_off(mq)] : 0.0;
                        *reinterpret_cast<double2 *>(nxt + i) = v;
                    });
                    nxt[T::INV] = rcp_nr(a[T::blk_off(mq) + q - G * mq]);
                }
                __syncwarp();
            }
            // save A = -(Schur block) right before its first pivot (p == N-1 just finished)
            if constexpr (p == N - 1) {
                sfor<T::MK0, MC>([&](auto mc) {
                    constexpr int m = decltype(mc)::value;
                    const int j = l + G * m;
                    if (j >= N && j < NT) {
                        sfor<(G * m > N ? G * m : N), NT>([&](auto ic) {
                            constexpr int i = decltype(ic)::value;
                            if (i >= j) {
                                const double v = -a[T::blk_off(m) + i - G * m];
                                S.As[grp][i - N][j - N] = v;
                                S.As[grp][j - N][i - N] = v;
                            }
                        });
                    }
                });
            }
        });

        // -------------------------------------------------- trace certificate, back substitution
        // per K-column of this lane: d (pivot), y (rhs), rows of L^-1; trace(A), trace(A^-1)
        double trA = 0.0, trAinv = 0.0;
        double invd[MC], acc[MC];
        sfor<T::MK0, MC>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            const int j = l + G * m;
            const bool kcol = (j >= N) && (j < NT);
            // the pivot of column j sits at row j of the block: pick it with a static scan
            double dj = 1.0;
            sfor<(G * m > N ? G * m : N), (G * (m + 1) < NT ? G * (m + 1) : NT)>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                if (i == j) dj = a[T::blk_off(m) + i - G * m];
            });
            invd[m] = kcol ? rcp_nr(dj) : 0.0;
            acc[m] = 0.0;
            double s2 = 0.0;
            sfor<0, K>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const double y = a[T::blk_off(m) + (NT + 1 + r) - G * m];
                s2 = fma(y, y, s2);
            });
            // rows r > c of the identity block were never touched and still hold 0 (or the unit)
            trAinv += kcol ? -s2 * invd[m] : 0.0;
        });
        for (int i = lane % G; i < K; i += G) trA += S.As[grp][i][i];
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            trA += __shfl_xor_sync(0xffffffffu, trA, o);
            trAinv += __shfl_xor_sync(0xffffffffu, trAinv, o);
        }
        // w = A^-1 g : L^T w = D^-1 L^-1 (-g), columns from last to first
        sfor<0, K>([&](auto cc) {
            constexpr int c = K - 1 - decltype(cc)::value;
            constexpr int j = N + c;
            constexpr int mj = j / G, lj = j % G;
            const double yj = a[T::blk_off(mj) + T::RHS - G * mj];
            double wc = (yj + acc[mj]) * invd[mj];
            wc = __shfl_sync(0xffffffffu, wc, grp * G + lj);
            if (l == 0) S.w[grp][c] = wc;
            // fold w_c into the pending sums of the earlier columns
            sfor<T::MK0, MC>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                if constexpr (G * m <= j && j < T::rows_end(m)) {
                    const int jj = l + G * m;
                    const double lij = a[T::blk_off(m) + j - G * m];
                    acc[m] = (jj >= N && jj < j) ? fma(-lij, wc, acc[m]) : acc[m];
                }
            });
        });
        const bool small_det = !(fabs(det) >= kDetThreshold);
        const bool certified = (trA * trAinv < 1.0 / kPinvRcond);
        const bool hard = a_bad || (small_det && !certified);
        if (l == 0) {
            int fl = 0;
            if (m_bad) fl |= IRLOSC_ST_M_NOT_PD;
            if (small_det && !a_bad) fl |= IRLOSC_ST_PINV;
            if (fl) S.flags[grp] |= fl;
        }
        __syncwarp();
        // uncertified instances: exact pinv / eigen inverse, one instance at a time, whole warp
        unsigned hard_mask = __ballot_sync(0xffffffffu, hard && valid && (l == 0));
        while (hard_mask) {
            const int src = __ffs(hard_mask) - 1;
            hard_mask &= hard_mask - 1;
            const int gi = src / G;
            const bool gi_abad = __shfl_sync(0xffffffffu, a_bad ? 1 : 0, src) != 0;
            eigen_solve<K>(S.As[gi], S.Vs, S.g[gi], S.w[gi], S.u[gi], S.dx[gi], !gi_abad, lane, &S.flags[gi]);
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
        // the vector stage is free again: next tile
        if (next_full && lane == 0) { fence_proxy_async(); issue_V(next); }
    }
}

}  // namespace tiled

}  // namespace irlosc

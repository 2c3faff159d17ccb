// Kinematic-tree-sparse DualUR5 OSC step kernel (sm_100a) - the default for the DualUR5.
//
// MuJoCo's joint-space inertia has branch-induced sparsity (Featherstone): M[i][j] != 0 only
// if one joint is an ancestor of the other, and the Jacobian of an end-effector body is
// non-zero only in the columns of its ancestor joints.  For the DualUR5
//
//     joint 0 (stand) -+- right arm 1..6 -+- 7 - 8        (left_outer_knuckle - left_inner_finger)
//                      |                  +- 9            (left_inner_knuckle)
//                      |                  +- 10 - 11      (right_outer_knuckle - right_inner_finger)
//                      |                  +- 12           (right_inner_knuckle)
//                      +- left arm 13..18 -+- 19 - 20, 21, 22 - 23, 24   (same shape)
//                                                                    (scenes/dual_ur5.xml:55-251)
//
// only 155 of the 325 lower-triangle entries of M are structurally non-zero, and eliminating
// the joints leaves-first (gripper joints, then arm joints 6..1, then the stand joint)
// produces NO fill-in.  The same augmented elimination as osc_tiled.cuh / osc_rows.cuh
//        S = [[M, J^T], [J, 0]]  ->  Schur block -A = -J M^-1 J^T  ->  w = A^-1 g
// then needs ~830 multiply-adds for k = 7 instead of ~5500 dense ones, and the two arms are
// independent until the stand joint couples them.  Same result as the dense elimination
// because the skipped entries are exact zeros (a contract of irlosc_params.has_topology;
// check_topology verifies it per instance).
//
// Mapping: 4 lanes per robot instance (8 instances per warp), lane = (arm a, half h):
//   phase A  each lane eliminates one gripper sub-branch {gb+1, gb, gb+2} of its arm into a
//            private copy of the 7x7 block over [stand, arm joints]; the two halves are then
//            summed with one xor-shuffle exchange.
//   phase B  lane h = 0 of each arm eliminates arm joints 6..1 together with the arm's task
//            rows of J, leaving the arm's contribution to the stand pivot, the stand column of
//            its task rows and its block of -A.
//   phase C  the stand joint couples everything (rank-1 update); the dense k x k system A w = g
//            is then solved by the 4 lanes with the row-distributed, shuffle-broadcast
//            LDL^T of osc_rows.cuh (rhs and identity rows give w and trace(A^-1) directly).
// uv = M dq, dx = J dq and J^T w use the same sparsity.  Inputs arrive by 1-D TMA bulk copies
// into a per-warp shared-memory stage exactly as in osc_rows.cuh.
//
// Reference restated: ir-lab/irl_control osc.py:41-68, 150-152, 156-181, 184-210.
#pragma once
#include "osc_rows.cuh"

namespace irlosc {
namespace tree {

using tiled::sfor;
using tiled::rcp_nr;
constexpr int kTreeWarps = 1; // warps are independent pipelines: 1-warp CTAs pack shared memory best

constexpr int kN = 25;        // DualUR5 robot DoF
constexpr int kG = 4;         // lanes per instance
constexpr int kWI = 8;        // instances per warp

// ancestors-or-self bit masks of the DualUR5 joints (bit j of kAnc[i]: joint j moves joint i's body)
__host__ __device__ constexpr uint32_t dual_ur5_anc(int i) {
    constexpr int parent[kN] = {-1, 0, 1, 2, 3, 4, 5, 6, 7, 6, 6, 10, 6, 0, 13, 14, 15, 16, 17, 18, 19, 18, 18, 22, 18};
    uint32_t m = 0;
    for (int j = i; j >= 0; j = parent[j]) m |= (1u << j);
    return m;
}
constexpr int kDualUr5Parent[kN] = {-1, 0, 1, 2, 3, 4, 5, 6, 7, 6, 6, 10, 6, 0, 13, 14, 15, 16, 17, 18, 19, 18, 18, 22, 18};

struct Roles {            // which target device plays which part (indices into KParams.dev)
    int dev_arm[2];       // device whose EE hangs off arm joint 6 / 18
    int dev_base;         // device whose EE hangs off the stand joint, -1 if not targeted
    int row_arm[2];       // first stacked task row of each arm
    int row_base;
};

template <int KD, bool HAS_BASE, bool PACKED>
struct TreeSmem {
    static constexpr int N = kN, WI = kWI;
    static constexpr int D = 2 + (HAS_BASE ? 1 : 0);
    static constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
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
    double uv[WI][N], dx[WI][K], g[WI][K], u[WI][N], j0[WI][K];
    double inv0[WI];
    int vel_zero[WI][D];
    int flags[WI];
    alignas(8) unsigned long long bar_m;
    alignas(8) unsigned long long bar_v;
};

// Dense SPD k x k solve A w = g on 4 lanes (rows i -> lane i % 4), shuffle-broadcast LDL^T with
// the rhs and identity rows carried along (see osc_rows.cuh).  Returns det, trace(A^-1), a_bad.
template <int K>
struct KStage {
    static constexpr int G = kG;
    static constexpr int RB = (K + G - 1) / G;
    static constexpr int NX = K + 1;
    static constexpr int XS = (NX + G - 1) / G;
    __host__ __device__ static constexpr int rl(int m) { return G * (m + 1) < K ? G * (m + 1) : K; }
    __host__ __device__ static constexpr int roff(int m) {
        int o = 0;
        for (int i = 0; i < m; ++i) o += rl(i);
        return o;
    }
    static constexpr int TOT = roff(RB);
};

template <int KD, bool HAS_BASE, bool PACKED, int MINB>
__global__ void __launch_bounds__(kTreeWarps * 32, MINB)
osc_step_tree(const KParams P, const KIo io, const int64_t B, const Roles R) {
    using WS = TreeSmem<KD, HAS_BASE, PACKED>;
    constexpr int N = kN, WI = kWI, G = kG, D = WS::D, K = WS::K, MSZ = WS::MSZ;
    constexpr unsigned FULL = 0xffffffffu;
    using KS = KStage<K>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, l = lane % G;
    const int arm = l >> 1, h = l & 1;
    WS &S = reinterpret_cast<WS *>(smem_raw)[warp];

    if (lane == 0) {
        tiled::mbar_init(&S.bar_m, 1);
        tiled::mbar_init(&S.bar_v, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int64_t n_tiles = (B + WI - 1) / WI;
    const int64_t warp_global = (int64_t)blockIdx.x * kTreeWarps + warp;
    const int64_t warp_stride = (int64_t)gridDim.x * kTreeWarps;
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
        for (int s = valid; s < WI; ++s) {
            for (int dd = lane; dd < D; dd += 32) { S.ee_quat[(s * D + dd) * 4] = 1.0; S.t_quat[(s * D + dd) * 4] = 1.0; }
            // padded instances need a non-singular A: unit Jacobian entries on distinct joints
            for (int r = lane; r < K; r += 32) S.J[(s * K + r) * N + r] = 1.0;
        }
    };

    if (warp_global < n_tiles && tile_full(warp_global) && lane == 0) { issue_M(warp_global); issue_V(warp_global); }

    // per-lane constants of the tree walk
    const int jb = 1 + 12 * arm;            // first arm joint of this lane's arm
    const int gb = jb + 6 + 3 * h;          // gripper sub-branch: gb (child of arm joint 6), gb+1 (child of gb), gb+2
    const int row_a = R.row_arm[arm];
    auto rowoff = [&](int i) { return PACKED ? i * (i + 1) / 2 : i * N; };

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
        // zero the K x K staging of A (block-diagonal part is filled in phase B)
        for (int e = l; e < K * (K + 1); e += G) (&S.As[grp][0][0])[e] = 0.0;

        // optional verification of the declared sparsity (irlosc_params.check_topology)
        bool sparse_bad = false;
        if (P.check_topology) {
            for (int i = 0; i < N; ++i) {
                const uint32_t anc = dual_ur5_anc(i);
                for (int j = l; j < (PACKED ? i + 1 : N); j += G) {
                    const bool related = (j <= i) ? ((anc >> j) & 1u) : ((dual_ur5_anc(j) >> i) & 1u);
                    const double v = Ms[PACKED ? i * (i + 1) / 2 + j : i * N + j];
                    if (!related && v != 0.0) sparse_bad = true;
                }
            }
            for (int r = 0; r < K; ++r) {
                int ee = 0;
                if (r >= R.row_arm[0] && r < R.row_arm[0] + KD) ee = 6;
                if (r >= R.row_arm[1] && r < R.row_arm[1] + KD) ee = 18;
                const uint32_t anc = dual_ur5_anc(ee);
                for (int j = l; j < N; j += G)
                    if (!((anc >> j) & 1u) && Js[r * N + j] != 0.0) sparse_bad = true;
            }
        }

        // ================================================== phase A: 7x7 block + gripper sub-branch
        // C index 0 = stand joint, 1..6 = arm joints jb..jb+5.  c[i(i+1)/2 + j], i >= j.
        double c[28], dqc[7], uvC[7];
        dqc[0] = dqs[0];
        sfor<1, 7>([&](auto ic) { dqc[decltype(ic)::value] = dqs[jb + decltype(ic)::value - 1]; });
        sfor<0, 7>([&](auto ic) { uvC[decltype(ic)::value] = 0.0; });
        {
            const double m00 = Ms[0];
            c[0] = (l == 0) ? m00 : 0.0;                      // M[0][0] enters once (right arm, h = 0)
            uvC[0] = c[0] * dqc[0];
        }
        sfor<1, 7>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            const int ro = rowoff(jb + i - 1);
            sfor<0, i + 1>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                double v = Ms[ro + (j == 0 ? 0 : jb + j - 1)];
                v = (h == 0) ? v : 0.0;                       // the h = 1 copy accumulates updates only
                c[i * (i + 1) / 2 + j] = v;
                uvC[i] = fma(v, dqc[j], uvC[i]);
                if constexpr (j != i) uvC[j] = fma(v, dqc[i], uvC[j]);
            });
        });
        // gripper rows of this lane: r = 0 -> gb, 1 -> gb+1 (child of gb), 2 -> gb+2
        double rg[3][7], dg[3], dqg[3], uvg[3];
        sfor<0, 3>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            const int gj = gb + r;
            const int ro = rowoff(gj);
            dqg[r] = dqs[gj];
            dg[r] = Ms[ro + gj];
            uvg[r] = dg[r] * dqg[r];
            sfor<0, 7>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                const double v = Ms[ro + (i == 0 ? 0 : jb + i - 1)];
                rg[r][i] = v;
                uvg[r] = fma(v, dqc[i], uvg[r]);
                uvC[i] = fma(v, dqg[r], uvC[i]);
            });
        });
        const double e10 = Ms[rowoff(gb + 1) + gb];           // M[gb+1][gb]
        uvg[1] = fma(e10, dqg[0], uvg[1]);
        uvg[0] = fma(e10, dqg[1], uvg[0]);
        bool m_bad = false;
        // eliminate gb+1 (leaf): touches the 7x7 block, row gb and pivot gb
        {
            m_bad = m_bad || !(dg[1] > 0.0);
            const double inv = rcp_nr(dg[1]);
            double t[7];
            sfor<0, 7>([&](auto ic) { t[decltype(ic)::value] = rg[1][decltype(ic)::value] * inv; });
            const double t7 = e10 * inv;
            sfor<0, 7>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                sfor<0, i + 1>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    c[i * (i + 1) / 2 + j] = fma(-rg[1][i], t[j], c[i * (i + 1) / 2 + j]);
                });
                rg[0][i] = fma(-rg[1][i], t7, rg[0][i]);
            });
            dg[0] = fma(-e10, t7, dg[0]);
        }
        // eliminate gb, then gb+2
        sfor<0, 2>([&](auto qc) {
            constexpr int r = decltype(qc)::value == 0 ? 0 : 2;
            m_bad = m_bad || !(dg[r] > 0.0);
            const double inv = rcp_nr(dg[r]);
            double t[7];
            sfor<0, 7>([&](auto ic) { t[decltype(ic)::value] = rg[r][decltype(ic)::value] * inv; });
            sfor<0, 7>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                sfor<0, i + 1>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    c[i * (i + 1) / 2 + j] = fma(-rg[r][i], t[j], c[i * (i + 1) / 2 + j]);
                });
            });
        });
        sfor<0, 3>([&](auto rc) { S.uv[grp][gb + decltype(rc)::value] = uvg[decltype(rc)::value]; });
        // merge the two halves of the arm: h = 0 receives the other half's updates
        sfor<0, 28>([&](auto ec) { c[decltype(ec)::value] += __shfl_xor_sync(FULL, c[decltype(ec)::value], 1); });
        sfor<0, 7>([&](auto ec) { uvC[decltype(ec)::value] += __shfl_xor_sync(FULL, uvC[decltype(ec)::value], 1); });
        m_bad = __shfl_xor_sync(FULL, m_bad ? 1 : 0, 1) != 0 || m_bad;

        // the M stage is free: pull the next tile's M
        __syncwarp();
        const int64_t next = tile + warp_stride;
        const bool next_full = next < n_tiles && tile_full(next);
        if (next_full && lane == 0) { tiled::fence_proxy_async(); issue_M(next); }

        // ================================================== phase B: arm joints 6..1 with the arm's task rows
        double jr[KD][7], ak[KD * (KD + 1) / 2];
        sfor<0, KD>([&](auto cc) {
            constexpr int cr = decltype(cc)::value;
            double dxc = 0.0;
            sfor<0, 7>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                const double v = Js[(row_a + cr) * N + (i == 0 ? 0 : jb + i - 1)];
                jr[cr][i] = v;
                dxc = fma(v, dqc[i], dxc);
            });
            if (h == 0) S.dx[grp][row_a + cr] = dxc;
        });
        sfor<0, KD *(KD + 1) / 2>([&](auto ec) { ak[decltype(ec)::value] = 0.0; });
        sfor<0, 6>([&](auto kc) {
            constexpr int k = 6 - decltype(kc)::value;         // C index of the pivot: 6, 5, ..., 1
            const double d = c[k * (k + 1) / 2 + k];
            m_bad = m_bad || !(d > 0.0);
            const double inv = rcp_nr(d);
            double t[k];
            sfor<0, k>([&](auto jc) { t[decltype(jc)::value] = c[k * (k + 1) / 2 + decltype(jc)::value] * inv; });
            sfor<0, k>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                sfor<0, i + 1>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    c[i * (i + 1) / 2 + j] = fma(-c[k * (k + 1) / 2 + i], t[j], c[i * (i + 1) / 2 + j]);
                });
            });
            sfor<0, KD>([&](auto cc) {
                constexpr int cr = decltype(cc)::value;
                const double jk = jr[cr][k];
                sfor<0, k>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    jr[cr][j] = fma(-jk, t[j], jr[cr][j]);
                });
                const double tc = jk * inv;
                sfor<cr, KD>([&](auto c2) {
                    constexpr int c2r = decltype(c2)::value;                 // rows c2r >= cr
                    ak[c2r * (c2r + 1) / 2 + cr] = fma(-jr[c2r][k], tc, ak[c2r * (c2r + 1) / 2 + cr]);
                });
            });
        });
        // ================================================== phase C: stand joint couples the arms
        const double d0 = c[0] + __shfl_xor_sync(FULL, c[0], 2);
        m_bad = m_bad || !(d0 > 0.0);
        m_bad = __shfl_xor_sync(FULL, m_bad ? 1 : 0, 2) != 0 || m_bad;
        const double uv0 = uvC[0] + __shfl_xor_sync(FULL, uvC[0], 2);
        if (h == 0) {
            sfor<0, KD>([&](auto cc) {
                constexpr int cr = decltype(cc)::value;
                S.j0[grp][row_a + cr] = jr[cr][0];
                sfor<0, cr + 1>([&](auto c2) {
                    constexpr int c2r = decltype(c2)::value;
                    const double v = -ak[cr * (cr + 1) / 2 + c2r];          // A block = -(Schur block)
                    S.As[grp][row_a + cr][row_a + c2r] = v;
                    S.As[grp][row_a + c2r][row_a + cr] = v;
                });
            });
            sfor<1, 7>([&](auto ic) { S.uv[grp][jb + decltype(ic)::value - 1] = uvC[decltype(ic)::value]; });
        }
        if (l == 0) {
            S.uv[grp][0] = uv0;
            S.inv0[grp] = rcp_nr(d0);
            if constexpr (HAS_BASE) {
                const double jb0 = Js[R.row_base * N + 0];
                S.j0[grp][R.row_base] = jb0;
                S.dx[grp][R.row_base] = jb0 * dqc[0];
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

        // -------------------------------------------------- dense A = blocks + j0 j0^T / d0, rows -> lanes
        double a[KS::TOT];
        double trA = 0.0;
        {
            const double inv0 = S.inv0[grp];
            sfor<0, KS::RB>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                const int i = l + G * m;
                const double j0i = (i < K) ? S.j0[grp][i] * inv0 : 0.0;
                sfor<0, KS::rl(m)>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    double v = 0.0;
                    if (i < K && j <= i) {
                        v = fma(j0i, S.j0[grp][j], S.As[grp][i][j]);
                        if (j == i) trA += v;
                    }
                    a[KS::roff(m) + j] = v;
                });
            });
            __syncwarp();
            // full A back to shared memory for the (rare) eigen path
            sfor<0, KS::RB>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                const int i = l + G * m;
                sfor<0, KS::rl(m)>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    if (i < K && j <= i) { S.As[grp][i][j] = a[KS::roff(m) + j]; S.As[grp][j][i] = a[KS::roff(m) + j]; }
                });
            });
        }
        // extra rows: rhs g and identity
        double x[KS::XS][K];
        sfor<0, KS::XS>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            const int e = l + G * s;
            sfor<0, K>([&](auto cc) {
                constexpr int cr = decltype(cc)::value;
                x[s][cr] = (e == 0) ? S.g[grp][cr] : ((e == cr + 1) ? 1.0 : 0.0);
            });
        });
        double detinv = 1.0;
        bool a_bad = false;
        double wacc[KS::XS], tin[KS::XS];
        sfor<0, KS::XS>([&](auto sc) { wacc[decltype(sc)::value] = 0.0; tin[decltype(sc)::value] = 0.0; });
        double invc = rcp_nr(a[0]);
        sfor<0, K>([&](auto pc) {
            constexpr int p = decltype(pc)::value;
            constexpr int mp = p / G;
            constexpr int mlo = (p + 1) / G;
            const double inv = __shfl_sync(FULL, invc, p % G, G);
            detinv *= inv;
            a_bad = a_bad || !(inv > 0.0);
            double mult[KS::RB];
            sfor<mlo, KS::RB>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                double t = a[KS::roff(m) + p] * inv;
                if constexpr (m == mp) t = (l > p % G) ? t : 0.0;
                mult[m] = t;
            });
            double multx[KS::XS];
            sfor<0, KS::XS>([&](auto sc) { multx[decltype(sc)::value] = x[decltype(sc)::value][p] * inv; });
            const double z = __shfl_sync(FULL, multx[0], 0, G);      // rhs row: z_p = (D^-1 L^-1 g)_p
            sfor<0, KS::XS>([&](auto sc) {
                constexpr int s = decltype(sc)::value;
                wacc[s] = fma(x[s][p], z, wacc[s]);
                tin[s] = fma(x[s][p] * x[s][p], inv, tin[s]);
            });
            sfor<p + 1, K>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                constexpr int mj = j / G;
                const double xj = __shfl_sync(FULL, a[KS::roff(mj) + p], j % G, G);
                sfor<(mj > mlo ? mj : mlo), KS::RB>([&](auto mc) {
                    constexpr int m = decltype(mc)::value;
                    a[KS::roff(m) + j] = fma(-xj, mult[m], a[KS::roff(m) + j]);
                });
                sfor<0, KS::XS>([&](auto sc) {
                    constexpr int s = decltype(sc)::value;
                    x[s][j] = fma(-xj, multx[s], x[s][j]);
                });
                if constexpr (j == p + 1) invc = rcp_nr(a[KS::roff(mj) + j]);
            });
        });
        double trAinv = 0.0;
        sfor<0, KS::XS>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            const int e = l + G * s;
            if (e >= 1 && e < KS::NX) { trAinv += tin[s]; S.w[grp][e - 1] = wacc[s]; }
        });
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            trA += __shfl_xor_sync(FULL, trA, o);
            trAinv += __shfl_xor_sync(FULL, trAinv, o);
            sparse_bad = __shfl_xor_sync(FULL, sparse_bad ? 1 : 0, o) != 0 || sparse_bad;
        }
        m_bad = __shfl_xor_sync(FULL, m_bad ? 1 : 0, 1) != 0 || m_bad;
        const bool small_det = !(fabs(detinv) <= 1.0 / kDetThreshold);   // |det A| < 1e-4 (osc.py:52)
        const bool certified = (trA * trAinv < 1.0 / kPinvRcond);
        const bool hard = a_bad || (small_det && !certified);
        if (l == 0) {
            int fl = 0;
            if (m_bad) fl |= IRLOSC_ST_M_NOT_PD;
            if (small_det && !a_bad) fl |= IRLOSC_ST_PINV;
            if (sparse_bad) fl |= IRLOSC_ST_SPARSITY;
            if (fl) S.flags[grp] |= fl;
        }
        __syncwarp();
        unsigned hard_mask = __ballot_sync(FULL, hard && valid && (l == 0));
        while (hard_mask) {
            const int src = __ffs(hard_mask) - 1;
            hard_mask &= hard_mask - 1;
            const int gi = src / G;
            const bool gi_abad = __shfl_sync(FULL, a_bad ? 1 : 0, src) != 0;
            tiled::eigen_solve<K>(S.As[gi], S.Vs, S.g[gi], S.w[gi], S.u[gi], !gi_abad, lane, &S.flags[gi]);
        }
        __syncwarp();

        // -------------------------------------------------- joint-space assembly (osc.py:174,184-200)
        const int flg = S.flags[grp];
        const bool poison = (flg & (IRLOSC_ST_M_NOT_PD | IRLOSC_ST_DX_RANGE | IRLOSC_ST_SPARSITY)) != 0;
        auto finish = [&](int j, double jt) {
            const double uvj = S.uv[grp][j];
            double u = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (S.vel_zero[grp][d] && ((P.dev[d].joint_mask >> j) & 1u)) u = -1.0 * P.dev[d].kv * uvj;
            u -= jt;
            if (P.use_g) u += S.bias[grp * N + j];
            if (P.has_nullspace) u -= P.nullspace_kv * uvj;
            if (poison) u = nan("");
            S.u[grp][j] = u;
            if (io.u_all && valid) io.u_all[inst * N + j] = u;
        };
        // gripper joints of this lane: J is zero there
        sfor<0, 3>([&](auto rc) { finish(gb + decltype(rc)::value, 0.0); });
        // arm joints: h = 0 takes 1..3, h = 1 takes 4..6 (C indices); only the arm's own task rows reach them
        sfor<0, 3>([&](auto tc) {
            const int j = jb + 3 * h + decltype(tc)::value;
            double jt = 0.0;
            sfor<0, KD>([&](auto cc) {
                constexpr int cr = decltype(cc)::value;
                jt = fma(Js[(row_a + cr) * N + j], S.w[grp][row_a + cr], jt);
            });
            finish(j, jt);
        });
        if (l == 0) {                                         // stand joint: every task row
            double jt = 0.0;
            sfor<0, K>([&](auto rc) { jt = fma(Js[decltype(rc)::value * N], S.w[grp][decltype(rc)::value], jt); });
            finish(0, jt);
        }
        __syncwarp();
        // -------------------------------------------------- packing (osc.py:203-208)
#pragma unroll
        for (int t = 0; t < (32 + G - 1) / G; ++t) {
            const int cidx = l + G * t;
            if (cidx < P.n_ctrl && valid) {
                int d = 0;
                while (d + 1 < D && cidx >= P.dev[d + 1].ctrl0) ++d;
                io.ctrl[inst * P.n_ctrl + cidx] = S.u[grp][P.dev[d].actuator[cidx - P.dev[d].ctrl0]];
            }
        }
        if (io.status && valid && l == 0) io.status[inst] = (uint8_t)flg;
        __syncwarp();
        if (next_full && lane == 0) { tiled::fence_proxy_async(); issue_V(next); }
    }
}

}  // namespace tree
}  // namespace irlosc

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
//   phase B  arm joints 6..1 are eliminated together with the arm's task rows of J, leaving the
//            arm's contribution to the stand pivot, the stand column of its task rows and its
//            block of -A.
//   phase C  the stand joint couples everything (rank-1 update); the dense k x k system A w = g
//            is then solved by the 4 lanes with the row-distributed, shuffle-broadcast
//            LDL^T of osc_rows.cuh (rhs and identity rows give w and trace(A^-1) directly).
// uv = M dq, dx = J dq and J^T w use the same sparsity.
//
// HBM -> SM: ncu on the first version (one private 51 KB TMA stage per warp) showed 1 warp per
// SM sub-partition and 18 % issue utilisation, i.e. occupancy-bound by shared memory.  Here a
// CTA of W warps shares a ring of NS input slots filled by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx).  A warp claims the next tile, waits for its slot,
// pulls everything it needs into registers / a small private scratch, releases the slot and
// re-arms it for the tile NS ahead, and only then runs the arithmetic - so a slot is busy for
// a fraction of a tile's life and W can be twice what private stages would allow.
//
// Reference restated: ir-lab/irl_control osc.py:41-68, 150-152, 156-181, 184-210.
#pragma once
#include "osc_tma.cuh"

namespace irlosc {
namespace tree {

using tiled::sfor;
using tiled::rcp_nr;

constexpr int kN = 25;        // DualUR5 robot DoF
constexpr int kG = 4;         // lanes per instance
constexpr int kWI = 8;        // instances per warp

constexpr int kDualUr5Parent[kN] = {-1, 0, 1, 2, 3, 4, 5, 6, 7, 6, 6, 10, 6, 0, 13, 14, 15, 16, 17, 18, 19, 18, 18, 22, 18};
// ancestors-or-self bit mask (bit j set: joint j moves joint i's body)
__host__ __device__ constexpr uint32_t dual_ur5_anc(int i) {
    constexpr int parent[kN] = {-1, 0, 1, 2, 3, 4, 5, 6, 7, 6, 6, 10, 6, 0, 13, 14, 15, 16, 17, 18, 19, 18, 18, 22, 18};
    uint32_t m = 0;
    for (int j = i; j >= 0; j = parent[j]) m |= (1u << j);
    return m;
}

struct Roles {            // which target device plays which part (indices into KParams.dev)
    int dev_arm[2];       // device whose EE hangs off arm joint 6 / 18
    int dev_base;         // device whose EE hangs off the stand joint, -1 if not targeted
    int row_arm[2];       // first stacked task row of each arm
    int row_base;
    // shared-memory plan (bytes), computed by the host
    int opt_tvel, opt_mvel, opt_ftx, opt_ftr;   // double offsets of the optional arrays in a slot's tail, -1 = absent
    int opt_doubles;                            // doubles in the optional tail (per slot and per warp scratch)
    int slot_bytes, priv_bytes;
    int ctrl_vec;         // ctrl (and every gather target) is 16-byte aligned: packed rows go out as double2
};

// Fixed part of an input slot; the optional per-instance arrays follow in a tail.
// MuJoCo's sparse inertia qM (IRLOSC_M_QM) of the DualUR5 tree: dof i owns M[i][i], M[i][parent], ... down to the
// stand joint, stored from kQmAdr(i); 155 doubles in all.  Arm a's chain rows start at kQmArm(a), its gripper rows
// 27 doubles later (depths 2..7 of the six arm joints).
constexpr int kQmSize = 155;
__host__ __device__ constexpr int qm_adr(int i) {
    constexpr int parent[kN] = {-1, 0, 1, 2, 3, 4, 5, 6, 7, 6, 6, 10, 6, 0, 13, 14, 15, 16, 17, 18, 19, 18, 18, 22, 18};
    int adr = 0;
    for (int r = 0; r < i; ++r)
        for (int j = r; j >= 0; j = parent[j]) ++adr;
    return adr;
}
static_assert(qm_adr(1) == 1 && qm_adr(7) == 28 && qm_adr(13) == 78 && qm_adr(24) + 8 == kQmSize, "qM addressing");

template <int KD, bool HAS_BASE, bool PACKED, bool QM = false>
struct TreeSlot {
    static constexpr int N = kN, WI = kWI;
    static constexpr int D = 2 + (HAS_BASE ? 1 : 0);
    static constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    static constexpr int MSZ = QM ? kQmSize : (PACKED ? N * (N + 1) / 2 : N * N);
    alignas(16) double M[WI * MSZ];
    alignas(16) double J[WI * K * N];
    alignas(16) double dq[WI * N];
    alignas(16) double bias[WI * N];
    alignas(16) double ee_xyz[WI * 3 * D];
    alignas(16) double ee_quat[WI * 4 * D];
    alignas(16) double t_xyz[WI * 3 * D];
    alignas(16) double t_quat[WI * 4 * D];
    alignas(16) unsigned long long full;
    alignas(16) double tail[2];              // optional arrays start here (16-byte aligned)
};

// Per-warp scratch that outlives the slot.
template <int KD, bool HAS_BASE>
struct TreePriv {
    static constexpr int N = kN, WI = kWI;
    static constexpr int D = 2 + (HAS_BASE ? 1 : 0);
    static constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    alignas(16) double w[WI][(K + 1) & ~1];
    double As[WI][K][K + 1];
    double Vs[K][K + 1];
    double uv[WI][N];            // M dq, overwritten in place by the joint-space signal u
    alignas(16) double bias[WI][N];   // reused as the packed-output staging tile after assembly
    double dx[WI][K], g[WI][K], j0[WI][K];
    double jc[WI][2][KD][7];     // original J entries of the arm rows on [stand, arm joints] (for J^T w)
    double jst[WI];              // J[base row][stand]
    double inv0[WI];
    double ee_xyz[WI * 3 * D], ee_quat[WI * 4 * D], t_xyz[WI * 3 * D], t_quat[WI * 4 * D];   // copies for the task law
    int vel_zero[WI][D];
    int flags[WI];
    alignas(16) double tail[2];  // copy of the slot's optional arrays (same offsets)
};

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

__device__ __forceinline__ void mbar_arrive(void *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tiled::smem_u32(bar)) : "memory");
}

template <int KD, bool HAS_BASE, bool PACKED, int W, int NS, int GW, bool QM = false>
__global__ void __launch_bounds__(W * 32, 1)
osc_step_tree(const KParams P, const KIo io, const int64_t B, const Roles R) {
    using SLT = TreeSlot<KD, HAS_BASE, PACKED, QM>;
    using PVT = TreePriv<KD, HAS_BASE>;
    constexpr int N = kN, WI = kWI, G = kG, D = SLT::D, K = SLT::K, MSZ = SLT::MSZ;
    constexpr unsigned FULL = 0xffffffffu;
    using KS = KStage<K>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, l = lane % G;
    const int arm = l >> 1, h = l & 1;
    // CTA header: [0] next tile ticket, [1 + s] number of the last use slot s has been armed for
    int *next_tile = reinterpret_cast<int *>(smem_raw);
    volatile int *armed = reinterpret_cast<volatile int *>(smem_raw) + 1;
    constexpr int kHdr = 64;
    auto slot_at = [&](int s) -> SLT & { return *reinterpret_cast<SLT *>(smem_raw + kHdr + (size_t)s * R.slot_bytes); };
    PVT &PV = *reinterpret_cast<PVT *>(smem_raw + kHdr + (size_t)NS * R.slot_bytes + (size_t)warp * R.priv_bytes);

    const int64_t n_tiles = (B + WI - 1) / WI;
    const bool has_tvel = R.opt_tvel >= 0, has_mvel = R.opt_mvel >= 0, adm = P.admittance != 0;
    auto tile_of = [&](int k) { return (int64_t)blockIdx.x + (int64_t)k * gridDim.x; };
    auto tile_full = [&](int64_t t) { return (t + 1) * WI <= B; };
    // one lane: arm a slot and start the bulk copies of a full tile
    auto issue = [&](SLT &SL, int64_t t) {
        uint32_t bytes = WI * 8 * (MSZ + K * N + N + 14 * D);
        if (P.use_g) bytes += WI * 8 * N;
        if (has_tvel) bytes += WI * 8 * 6 * D;
        if (has_mvel) bytes += WI * 8 * 2 * D;
        if (adm) bytes += WI * 8 * 15 * D;
        tiled::mbar_expect_tx(&SL.full, bytes);
        const int64_t i0 = t * WI;
        tiled::bulk_g2s(SL.M, io.M + i0 * (int64_t)MSZ, WI * MSZ * 8, &SL.full);
        tiled::bulk_g2s(SL.J, io.J + i0 * (K * N), WI * K * N * 8, &SL.full);
        tiled::bulk_g2s(SL.dq, io.dq + i0 * N, WI * N * 8, &SL.full);
        if (P.use_g) tiled::bulk_g2s(SL.bias, io.bias + i0 * N, WI * N * 8, &SL.full);
        tiled::bulk_g2s(SL.ee_xyz, io.ee_xyz + i0 * 3 * D, WI * 3 * D * 8, &SL.full);
        tiled::bulk_g2s(SL.ee_quat, io.ee_quat + i0 * 4 * D, WI * 4 * D * 8, &SL.full);
        tiled::bulk_g2s(SL.t_xyz, io.target_xyz + i0 * 3 * D, WI * 3 * D * 8, &SL.full);
        tiled::bulk_g2s(SL.t_quat, io.target_quat + i0 * 4 * D, WI * 4 * D * 8, &SL.full);
        if (has_tvel) tiled::bulk_g2s(SL.tail + R.opt_tvel, io.target_vel + i0 * 6 * D, WI * 6 * D * 8, &SL.full);
        if (has_mvel) tiled::bulk_g2s(SL.tail + R.opt_mvel, io.max_vel + i0 * 2 * D, WI * 2 * D * 8, &SL.full);
        if (adm) {
            tiled::bulk_g2s(SL.tail + R.opt_ftx, io.ft_xmat + i0 * 9 * D, WI * 9 * D * 8, &SL.full);
            tiled::bulk_g2s(SL.tail + R.opt_ftr, io.ft_raw + i0 * 6 * D, WI * 6 * D * 8, &SL.full);
        }
    };
    auto arm_slot = [&](SLT &SL, int64_t t) {     // full tile: TMA; ragged tile: the consumer fills it by hand
        if (tile_full(t)) issue(SL, t); else mbar_arrive(&SL.full);
    };
    auto copy_rows = [&](double *dst, const double *src, int per, int64_t i0, int valid) {
        for (int e = lane; e < WI * per; e += 32) dst[e] = (e / per < valid) ? src[i0 * per + e] : 0.0;
    };
    auto manual_fill = [&](SLT &SL, int64_t t) {
        const int64_t i0 = t * WI;
        const int valid = (int)(B - i0);
        copy_rows(SL.M, io.M, MSZ, i0, valid);
        copy_rows(SL.J, io.J, K * N, i0, valid);
        copy_rows(SL.dq, io.dq, N, i0, valid);
        if (P.use_g) copy_rows(SL.bias, io.bias, N, i0, valid);
        copy_rows(SL.ee_xyz, io.ee_xyz, 3 * D, i0, valid);
        copy_rows(SL.ee_quat, io.ee_quat, 4 * D, i0, valid);
        copy_rows(SL.t_xyz, io.target_xyz, 3 * D, i0, valid);
        copy_rows(SL.t_quat, io.target_quat, 4 * D, i0, valid);
        if (has_tvel) copy_rows(SL.tail + R.opt_tvel, io.target_vel, 6 * D, i0, valid);
        if (has_mvel) copy_rows(SL.tail + R.opt_mvel, io.max_vel, 2 * D, i0, valid);
        if (adm) { copy_rows(SL.tail + R.opt_ftx, io.ft_xmat, 9 * D, i0, valid); copy_rows(SL.tail + R.opt_ftr, io.ft_raw, 6 * D, i0, valid); }
        for (int s = valid; s < WI; ++s) {       // padded instances: keep the arithmetic finite
            for (int i = lane; i < N; i += 32) SL.M[s * MSZ + (QM ? qm_adr(i) : PACKED ? i * (i + 1) / 2 + i : i * N + i)] = 1.0;
            for (int dd = lane; dd < D; dd += 32) { SL.ee_quat[(s * D + dd) * 4] = 1.0; SL.t_quat[(s * D + dd) * 4] = 1.0; }
        }
    };

    if (threadIdx.x == 0) {
        *next_tile = 0;
        for (int s = 0; s < NS; ++s) { tiled::mbar_init(&slot_at(s).full, 1); armed[s] = -1; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int s = 0; s < NS; ++s)
            if (tile_of(s) < n_tiles) { arm_slot(slot_at(s), tile_of(s)); __threadfence_block(); armed[s] = 0; }

    // per-lane constants of the tree walk
    const int jb = 1 + 12 * arm;            // first arm joint of this lane's arm
    const int gb = jb + 6 + 3 * h;          // gripper sub-branch: gb (child of arm joint 6), gb+1 (child of gb), gb+2
    const int row_a = R.row_arm[arm];
    auto rowoff = [&](int i) { return PACKED ? i * (i + 1) / 2 : i * N; };
    auto ccol = [&](int i) { return i == 0 ? 0 : jb + i - 1; };      // C index -> joint

    // Lock-step groups: GW warps claim GW consecutive tiles and walk through the phases together
    // (named barrier per group).  ncu showed the straight-line code (~60 KB per tile) streaming
    // through the instruction caches once per warp, with the GPC-level cache at 68 % of its
    // request rate; in lock step one fetched line serves GW warps.
    static_assert(W % GW == 0, "W must be a multiple of GW");
    constexpr int kPhaseBarriers = 5;
    volatile int *group_k0 = reinterpret_cast<volatile int *>(smem_raw) + 8;
    const int group = warp / GW;
    auto gsync = [&]() {
        if constexpr (GW > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(GW * 32) : "memory");
    };
    while (true) {
        int k = 0;
        if constexpr (GW > 1) {
            if (warp % GW == 0 && lane == 0) group_k0[group] = atomicAdd(next_tile, GW);
            gsync();
            k = group_k0[group] + warp % GW;
            if (tile_of(k - warp % GW) >= n_tiles) break;           // the whole group is done
            if (tile_of(k) >= n_tiles) {                              // nothing left for this warp: keep the barriers balanced
                for (int i = 0; i < kPhaseBarriers; ++i) gsync();
                continue;
            }
        } else {
            if (lane == 0) k = atomicAdd(next_tile, 1);
            k = __shfl_sync(FULL, k, 0);
            if (tile_of(k) >= n_tiles) break;
        }
        const int64_t tile = tile_of(k);
        SLT &SL = slot_at(k % NS);
        // the slot must have been re-armed for THIS use before its barrier phase means anything:
        // a parity wait alone would fall through on the phase of two uses ago
        if (lane == 0)
            while (armed[k % NS] < k / NS) __nanosleep(64);
        __syncwarp();
        tiled::mbar_wait(&SL.full, (uint32_t)((k / NS) & 1));
        if (!tile_full(tile)) { manual_fill(SL, tile); __syncwarp(); }
        const int64_t inst = tile * WI + grp;
        const bool valid = inst < B;
        const double *Ms = SL.M + grp * MSZ;
        const double *Js = SL.J + grp * K * N;
        const double *dqs = SL.dq + grp * N;
        if (l == 0) PV.flags[grp] = 0;
        for (int e = l; e < K * (K + 1); e += G) (&PV.As[grp][0][0])[e] = 0.0;

        double dqc[7];
        sfor<0, 7>([&](auto ic) { dqc[decltype(ic)::value] = dqs[ccol(decltype(ic)::value)]; });
        // optional verification of the declared sparsity (irlosc_params.check_topology)
        bool sparse_bad = false;
        if (P.check_topology) {
            for (int i = 0; i < (QM ? 0 : N); ++i) {       // qM holds the tree's entries only: nothing to verify in M
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

        // ================================================== pull the tile into registers / scratch
        // C index 0 = stand joint, 1..6 = arm joints jb..jb+5.  c[i(i+1)/2 + j], i >= j.
        double c[28], uvC[7];
        sfor<0, 7>([&](auto ic) { uvC[decltype(ic)::value] = 0.0; });
        {
            const double m00 = Ms[0];
            c[0] = (l == 0) ? m00 : 0.0;                      // M[0][0] enters once (right arm, h = 0)
            uvC[0] = c[0] * dqc[0];
        }
        const int qa = arm ? qm_adr(13) : qm_adr(1);          // qM: first entry of this arm's chain rows
        sfor<1, 7>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            const int ro = QM ? qa + (i - 1) * (i + 2) / 2 : rowoff(jb + i - 1);
            sfor<0, i + 1>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                double v = 0.0;                               // the h = 1 copy accumulates updates only
                // qM row of arm joint i: [itself, arm joints i-1 .. 1, stand]
                if (h == 0) v = QM ? Ms[ro + (j == 0 ? i : i - j)] : Ms[ro + ccol(j)];
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
            // qM rows of a gripper half: [gb, arm 6..1, stand] (8), [gb+1, gb, arm 6..1, stand] (9), [gb+2, arm 6..1, stand] (8)
            constexpr int qs = (r == 1) ? 1 : 0;
            const int ro = QM ? qa + 27 + 25 * h + (r == 0 ? 0 : r == 1 ? 8 : 17) : rowoff(gj);
            dqg[r] = dqs[gj];
            dg[r] = QM ? Ms[ro] : Ms[ro + gj];
            uvg[r] = dg[r] * dqg[r];
            sfor<0, 7>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                const double v = QM ? Ms[ro + (i == 0 ? 7 + qs : 7 + qs - i)] : Ms[ro + ccol(i)];
                rg[r][i] = v;
                uvg[r] = fma(v, dqc[i], uvg[r]);
                uvC[i] = fma(v, dqg[r], uvC[i]);
            });
        });
        const double e10 = QM ? Ms[qa + 27 + 25 * h + 8 + 1] : Ms[rowoff(gb + 1) + gb];           // M[gb+1][gb]
        uvg[1] = fma(e10, dqg[0], uvg[1]);
        uvg[0] = fma(e10, dqg[1], uvg[0]);
        sfor<0, 3>([&](auto rc) { PV.uv[grp][gb + decltype(rc)::value] = uvg[decltype(rc)::value]; });
        // task rows of this arm on [stand, arm joints]; originals go to scratch for J^T w
        double jr[KD][7];
        sfor<0, KD>([&](auto cc) {
            constexpr int cr = decltype(cc)::value;
            double dxc = 0.0;                                  // dx = J dq on the sparse columns (osc.py:150)
            sfor<0, 7>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                const double v = Js[(row_a + cr) * N + ccol(i)];
                jr[cr][i] = v;
                dxc = fma(v, dqc[i], dxc);
                if (h == 0) PV.jc[grp][arm][cr][i] = v;
            });
            if (h == 0) PV.dx[grp][row_a + cr] = dxc;
        });
        if constexpr (HAS_BASE) {
            if (l == 0) {
                const double jb0 = Js[R.row_base * N + 0];
                PV.jst[grp] = jb0;
                PV.j0[grp][R.row_base] = jb0;
                PV.dx[grp][R.row_base] = jb0 * dqc[0];
            }
        }
        if (P.use_g)
            for (int j = l; j < N; j += G) PV.bias[grp][j] = SL.bias[grp * N + j];
        // everything the task law needs later: poses, targets, optional arrays
        for (int e = lane; e < WI * 3 * D; e += 32) { PV.ee_xyz[e] = SL.ee_xyz[e]; PV.t_xyz[e] = SL.t_xyz[e]; }
        for (int e = lane; e < WI * 4 * D; e += 32) { PV.ee_quat[e] = SL.ee_quat[e]; PV.t_quat[e] = SL.t_quat[e]; }
        for (int e = lane; e < R.opt_doubles; e += 32) PV.tail[e] = SL.tail[e];
        __syncwarp();
        // -------------------------------------------------- slot consumed: hand it to the tile NS ahead
        if (lane == 0) {
            const int64_t tn = tile_of(k + NS);
            if (tn < n_tiles) {
                tiled::fence_proxy_async();
                arm_slot(SL, tn);
                __threadfence_block();
                armed[k % NS] = k / NS + 1;
            }
        }

        gsync();
        // ================================================== phase A: gripper sub-branch
        bool m_bad = false;
        {   // eliminate gb+1 (leaf): touches the 7x7 block, row gb and pivot gb
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
        sfor<0, 2>([&](auto qc) {      // eliminate gb, then gb+2
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
        // merge the two halves of the arm (both lanes end up with the sum)
        sfor<0, 28>([&](auto ec) { c[decltype(ec)::value] += __shfl_xor_sync(FULL, c[decltype(ec)::value], 1); });
        sfor<0, 7>([&](auto ec) { uvC[decltype(ec)::value] += __shfl_xor_sync(FULL, uvC[decltype(ec)::value], 1); });
        m_bad = __shfl_xor_sync(FULL, m_bad ? 1 : 0, 1) != 0 || m_bad;

        gsync();
        // ================================================== phase B: arm joints 6..1 with the arm's task rows
        double ak[KD * (KD + 1) / 2];
        sfor<0, KD *(KD + 1) / 2>([&](auto ec) { ak[decltype(ec)::value] = 0.0; });
        sfor<0, 6>([&](auto kc) {
            constexpr int kk = 6 - decltype(kc)::value;        // C index of the pivot: 6, 5, ..., 1
            const double d = c[kk * (kk + 1) / 2 + kk];
            m_bad = m_bad || !(d > 0.0);
            const double inv = rcp_nr(d);
            double t[kk];
            sfor<0, kk>([&](auto jc) { t[decltype(jc)::value] = c[kk * (kk + 1) / 2 + decltype(jc)::value] * inv; });
            sfor<0, kk>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                sfor<0, i + 1>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    c[i * (i + 1) / 2 + j] = fma(-c[kk * (kk + 1) / 2 + i], t[j], c[i * (i + 1) / 2 + j]);
                });
            });
            sfor<0, KD>([&](auto cc) {
                constexpr int cr = decltype(cc)::value;
                const double jk = jr[cr][kk];
                sfor<0, kk>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    jr[cr][j] = fma(-jk, t[j], jr[cr][j]);
                });
                const double tc = jk * inv;
                sfor<cr, KD>([&](auto c2) {
                    constexpr int c2r = decltype(c2)::value;                 // rows c2r >= cr
                    ak[c2r * (c2r + 1) / 2 + cr] = fma(-jr[c2r][kk], tc, ak[c2r * (c2r + 1) / 2 + cr]);
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
                PV.j0[grp][row_a + cr] = jr[cr][0];
                sfor<0, cr + 1>([&](auto c2) {
                    constexpr int c2r = decltype(c2)::value;
                    const double v = -ak[cr * (cr + 1) / 2 + c2r];          // A block = -(Schur block)
                    PV.As[grp][row_a + cr][row_a + c2r] = v;
                    PV.As[grp][row_a + c2r][row_a + cr] = v;
                });
            });
            sfor<1, 7>([&](auto ic) { PV.uv[grp][jb + decltype(ic)::value - 1] = uvC[decltype(ic)::value]; });
        }
        if (l == 0) {
            PV.uv[grp][0] = uv0;
            PV.inv0[grp] = rcp_nr(d0);
        }
        __syncwarp();

        gsync();
        // -------------------------------------------------- per-device task signal (osc.py:156-181)
        if (l < D) {
            const int d = l;
            const KDevice &dv = P.dev[d];
            const int sd = grp * D + d;
            double mv[2] = {dv.max_vel[0], dv.max_vel[1]};
            if (has_mvel) { mv[0] = PV.tail[R.opt_mvel + sd * 2]; mv[1] = PV.tail[R.opt_mvel + sd * 2 + 1]; }
            double tv[6], u6[6];
            if (has_tvel)
                for (int i = 0; i < 6; ++i) tv[i] = PV.tail[R.opt_tvel + sd * 6 + i];
            bool oob = false;
            const bool tracking = device_task_signal(dv, &PV.ee_xyz[sd * 3], &PV.ee_quat[sd * 4], &PV.t_xyz[sd * 3],
                                                     &PV.t_quat[sd * 4], has_tvel ? tv : nullptr, mv, PV.dx[grp], K,
                                                     u6, &oob);
            PV.vel_zero[grp][d] = tracking ? 0 : 1;
            double ft[6] = {0, 0, 0, 0, 0, 0};
            if (adm) rotate_wrench(&PV.tail[R.opt_ftx + sd * 9], &PV.tail[R.opt_ftr + sd * 6], ft);
            int r = dv.row0;
            const double kvn = P.has_nullspace ? P.nullspace_kv : 0.0;
            for (int i = 0; i < 6; ++i)
                if (dv.dof[i]) {
                    const double v = adm ? u6[i] + ft[i] : u6[i];
                    PV.g[grp][r] = v - kvn * PV.dx[grp][r];
                    ++r;
                }
            const int fl = (tracking ? IRLOSC_ST_VEL_BRANCH : 0) | (oob ? IRLOSC_ST_DX_RANGE : 0);
            if (fl) atomicOr(&PV.flags[grp], fl);
        }

        __syncwarp();

        gsync();
        // -------------------------------------------------- dense A = blocks + j0 j0^T / d0, rows -> lanes
        double a[KS::TOT];
        double trA = 0.0;
        {
            const double inv0 = PV.inv0[grp];
            sfor<0, KS::RB>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                const int i = l + G * m;
                const double j0i = (i < K) ? PV.j0[grp][i] * inv0 : 0.0;
                sfor<0, KS::rl(m)>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    double v = 0.0;
                    if (i < K && j <= i) {
                        v = fma(j0i, PV.j0[grp][j], PV.As[grp][i][j]);
                        if (j == i) trA += v;
                    }
                    a[KS::roff(m) + j] = v;
                });
            });
            __syncwarp();
            sfor<0, KS::RB>([&](auto mc) {       // full A back to scratch for the (rare) eigen path
                constexpr int m = decltype(mc)::value;
                const int i = l + G * m;
                sfor<0, KS::rl(m)>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    if (i < K && j <= i) { PV.As[grp][i][j] = a[KS::roff(m) + j]; PV.As[grp][j][i] = a[KS::roff(m) + j]; }
                });
            });
        }
        double x[KS::XS][K];             // extra rows: rhs g and identity
        sfor<0, KS::XS>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            const int e = l + G * s;
            sfor<0, K>([&](auto cc) {
                constexpr int cr = decltype(cc)::value;
                x[s][cr] = (e == 0) ? PV.g[grp][cr] : ((e == cr + 1) ? 1.0 : 0.0);
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
            if (e >= 1 && e < KS::NX) { trAinv += tin[s]; PV.w[grp][e - 1] = wacc[s]; }
        });
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            trA += __shfl_xor_sync(FULL, trA, o);
            trAinv += __shfl_xor_sync(FULL, trAinv, o);
            sparse_bad = __shfl_xor_sync(FULL, sparse_bad ? 1 : 0, o) != 0 || sparse_bad;
        }
        const bool small_det = !(fabs(detinv) <= 1.0 / kDetThreshold);   // |det A| < 1e-4 (osc.py:52)
        const bool certified = (trA * trAinv < 1.0 / kPinvRcond);
        const bool hard = a_bad || (small_det && !certified);
        if (l == 0) {
            int fl = 0;
            if (m_bad) fl |= IRLOSC_ST_M_NOT_PD;
            if (small_det && !a_bad) fl |= IRLOSC_ST_PINV;
            if (sparse_bad) fl |= IRLOSC_ST_SPARSITY;
            if (fl) atomicOr(&PV.flags[grp], fl);
        }
        __syncwarp();
        unsigned hard_mask = __ballot_sync(FULL, hard && valid && (l == 0));
        while (hard_mask) {
            const int src = __ffs(hard_mask) - 1;
            hard_mask &= hard_mask - 1;
            const int gi = src / G;
            const bool gi_abad = __shfl_sync(FULL, a_bad ? 1 : 0, src) != 0;
            tiled::eigen_solve<K>(PV.As[gi], PV.Vs, PV.g[gi], PV.w[gi], PV.dx[gi], PV.j0[gi], !gi_abad, lane, &PV.flags[gi]);
        }
        __syncwarp();

        gsync();
        // -------------------------------------------------- joint-space assembly (osc.py:174,184-200)
        const int flg = PV.flags[grp];
        const bool poison = (flg & (IRLOSC_ST_M_NOT_PD | IRLOSC_ST_DX_RANGE | IRLOSC_ST_SPARSITY)) != 0;
        auto finish = [&](int j, double jt) {
            const double uvj = PV.uv[grp][j];
            double u = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (PV.vel_zero[grp][d] && ((P.dev[d].joint_mask >> j) & 1u)) u = -1.0 * P.dev[d].kv * uvj;
            u -= jt;
            if (P.use_g) u += PV.bias[grp][j];
            if (P.has_nullspace) u -= P.nullspace_kv * uvj;
            if (poison) u = nan("");
            PV.uv[grp][j] = u;                                  // u replaces uv in place
            if (io.u_all && valid) io.u_all[inst * N + j] = u;
        };
        sfor<0, 3>([&](auto rc) { finish(gb + decltype(rc)::value, 0.0); });   // gripper joints: J is zero there
        sfor<0, 3>([&](auto tc) {     // arm joints: h = 0 takes C indices 1..3, h = 1 takes 4..6
            const int ci = 1 + 3 * h + decltype(tc)::value;
            double jt = 0.0;
            sfor<0, KD>([&](auto cc) {
                constexpr int cr = decltype(cc)::value;
                jt = fma(PV.jc[grp][arm][cr][ci], PV.w[grp][row_a + cr], jt);
            });
            finish(jb + ci - 1, jt);
        });
        if (l == 0) {                                         // stand joint: every task row
            double jt = 0.0;
            sfor<0, 2>([&](auto ac) {
                constexpr int aa = decltype(ac)::value;
                sfor<0, KD>([&](auto cc) {
                    constexpr int cr = decltype(cc)::value;
                    jt = fma(PV.jc[grp][aa][cr][0], PV.w[grp][R.row_arm[aa] + cr], jt);
                });
            });
            if constexpr (HAS_BASE) jt = fma(PV.jst[grp], PV.w[grp][R.row_base], jt);
            finish(0, jt);
        }
        __syncwarp();
        // -------------------------------------------------- packing (osc.py:203-208)
        // The tile's rows are contiguous in every destination: stage them in the (now dead) bias
        // scratch and write 16-byte vectors, so the peer stores of the fused gather leave the SM as
        // full 512-byte requests instead of 32-byte pieces (NVLink packet efficiency).
        double *ct = &PV.bias[0][0];
#pragma unroll
        for (int t = 0; t < (32 + G - 1) / G; ++t) {
            const int cidx = l + G * t;
            if (cidx < P.n_ctrl) {
                int d = 0;
                while (d + 1 < D && cidx >= P.dev[d + 1].ctrl0) ++d;
                ct[grp * P.n_ctrl + cidx] = PV.uv[grp][P.dev[d].actuator[cidx - P.dev[d].ctrl0]];
            }
        }
        __syncwarp();
        if (tile_full(tile) && R.ctrl_vec) {
            const double2 *src = reinterpret_cast<const double2 *>(ct);
            const int64_t row0 = tile * WI * (int64_t)P.n_ctrl;
            double2 *dst = reinterpret_cast<double2 *>(io.ctrl + row0);
            for (int e = lane; e < (WI / 2) * P.n_ctrl; e += 32) {
                const double2 v = src[e];
                dst[e] = v;
                if (io.ctrl_mc)
                    multimem_st(reinterpret_cast<double2 *>(io.ctrl_mc + io.gather_offset * P.n_ctrl + row0) + e, v);
                else
                    for (int g = 0; g < io.n_gather; ++g)
                        reinterpret_cast<double2 *>(io.ctrl_gather[g] + io.gather_offset * P.n_ctrl + row0)[e] = v;
            }
        } else if (valid) {
            for (int cidx = l; cidx < P.n_ctrl; cidx += G) store_ctrl(io, P.n_ctrl, inst, cidx, ct[grp * P.n_ctrl + cidx]);
        }
        if (io.status && valid && l == 0) io.status[inst] = (uint8_t)flg;
        __syncwarp();
    }
}

}  // namespace tree
}  // namespace irlosc

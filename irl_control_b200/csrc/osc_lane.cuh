// Lane kernel (sm_100a): one THREAD per robot instance, state read from BATCH-INTERLEAVED TILES.
//
// The control law needs ~9 % of an SM's FP64 rate and ~1 % of its issue slots to keep up with HBM, so what
// decides the speed of this path is how the ~300 doubles of an instance reach the lane that owns it.  A
// thread-per-instance kernel has no shuffles and no idle lanes (osc_stream.cuh), but plain per-variable
// arrays make every lane walk its own 2.6 - 4.9 KB record: 8-byte gathers at a multi-KB stride, which DRAM
// serves at about a third of its rate and which cost an LDGSTS each (round 1: 0.36 of the roofline).
//
// So the batch layout is made for the machine (the "interleaved" / "compact" batch layout batched
// small-matrix libraries use): instances are grouped in TILES of 32, and inside a tile every scalar of the
// state is stored for the 32 instances side by side,
//
//     tiles[t][e][l] = entry e of instance 32 t + l,          e < E,  l < 32,
//
// with the entries in exactly the order the leaves-first elimination consumes them (stand / base data, then per
// arm: M over [stand, arm joints] + dq, gripper half 0, gripper half 1, task rows of J + bias, the device's
// poses and targets) and only the entries the kinematic tree makes non-zero (155 of 325 for M - MuJoCo's own qM
// set -, 43 of 175 for J at k = 7).  A warp-wide load of entry e is then ONE fully coalesced 256-byte request,
// a tile is one contiguous 76 - 91 KB block of HBM read front to back exactly once, there is no staging
// buffer, no shared-memory round trip and no copy-issue code: the lanes load straight into the registers the
// elimination works in, and the warps of an SM (12 - 16, no shared memory needed) hide each other's latency.
// An optional L2 prefetch runs a few groups ahead (cp.async.bulk.prefetch.L2, one instruction per group).
//
// `irlosc_pack_tiles` (pack_tiles_kernel below) produces the layout from the per-variable arrays of irlosc_io
// for callers that hold MuJoCo-style arrays; a caller that assembles its batch itself writes tiles directly
// (irlosc_tile_spec gives the entry table).
//
// Arithmetic: the tree-sparse elimination shared with osc_stream.cuh (consume_cc / consume_grip / consume_rows),
// then osc_tail.cuh.  The original Jacobian entries needed for J^T w at the very end are re-read from the tile
// (L2 hits) instead of being carried in registers.
//
// Reference restated: ir-lab/irl_control osc.py:41-68, 120-210; robot.py:44-72; device.py:115-170.
#pragma once
#include <cstring>
#include "irlosc_device.cuh"
#include "osc_fused_types.h"
#include "osc_tail.cuh"
#include "osc_stream.cuh"
#include "osc_tma.cuh"
#include "irlosc_build.h"

namespace irlosc {
namespace lane {

using fused::FRoles;
using fused::Debug;
using fused::kN;
using stream::kGroups;

constexpr int kTile = 32;                   // instances per tile
constexpr int kMaxEntries = 400;

// symbolic source of a tile entry (mirrors IRLOSC_ARR_* of irlosc.h)
enum : int { kArrPad = 0, kArrM, kArrJ, kArrDq, kArrBias, kArrEeXyz, kArrEeQuat, kArrTXyz, kArrTQuat, kArrMaxVel,
              kArrFtX, kArrFtRaw, kArrCount };

struct TileSpec {
    int32_t n_entries;
    int32_t gbase[kGroups + 1];             // first entry of every group; gbase[kGroups] == n_entries
    irlosc_tile_entry e[kMaxEntries];
};

// Entry table of a controller configuration: the stage layout of osc_stream.cuh materialised in HBM.
// max_vel is always present (the device default when the caller has no per-instance array).
inline bool build_tile_spec(const KParams &P, const FRoles &R, int kd, bool has_base, TileSpec &S) {
    using namespace stream;
    memset(&S, 0, sizeof S);
    const int devN = P.admittance ? kDevEntries : kMaxVel + 2;
    int size[kGroups];
    size[0] = kG0Dev + (has_base ? devN : 0);
    for (int arm = 0; arm < 2; ++arm) {
        const int g0 = 1 + 5 * arm;
        size[g0] = kG1Entries;
        size[g0 + 1] = size[g0 + 2] = kGripEntries;
        size[g0 + 3] = kd * 7 + 6;
        size[g0 + 4] = devN;
    }
    int at = 0;
    for (int g = 0; g < kGroups; ++g) { S.gbase[g] = at; at += size[g]; }
    S.gbase[kGroups] = at;
    S.n_entries = at;
    if (at > kMaxEntries) return false;
    auto put = [&](int g, int e, int arr, int i, int j) { S.e[S.gbase[g] + e] = irlosc_tile_entry{arr, i, j}; };
    auto device_block = [&](int g, int d, int e0) {
        for (int i = 0; i < 3; ++i) { put(g, e0 + kEeXyz + i, kArrEeXyz, d, i); put(g, e0 + kTXyz + i, kArrTXyz, d, i); }
        for (int i = 0; i < 4; ++i) { put(g, e0 + kEeQuat + i, kArrEeQuat, d, i); put(g, e0 + kTQuat + i, kArrTQuat, d, i); }
        for (int i = 0; i < 2; ++i) put(g, e0 + kMaxVel + i, kArrMaxVel, d, i);
        if (P.admittance) {
            for (int i = 0; i < 9; ++i) put(g, e0 + kFtX + i, kArrFtX, d, i);
            for (int i = 0; i < 6; ++i) put(g, e0 + kFtRaw + i, kArrFtRaw, d, i);
        }
    };
    put(0, kG0Bias0, kArrBias, 0, 0);
    if (has_base) {
        put(0, kG0Jbase, kArrJ, R.row_base, 0);
        device_block(0, R.dev_base, kG0Dev);
    }
    for (int arm = 0; arm < 2; ++arm) {
        const int jb = 1 + 12 * arm, g0 = 1 + 5 * arm;
        auto C = [&](int i) { return i == 0 ? 0 : jb + i - 1; };
        for (int i = 0; i < 7; ++i) {
            for (int j = 0; j <= i; ++j) put(g0, kCC + i * (i + 1) / 2 + j, kArrM, C(i), C(j));
            put(g0, kDqC + i, kArrDq, C(i), 0);
        }
        for (int half = 0; half < 2; ++half) {
            const int g = g0 + 1 + half, gj = jb + 6 + 3 * half;
            for (int i = 0; i < 7; ++i) {
                put(g, kRg1 + i, kArrM, gj + 1, C(i));
                put(g, kRg0 + i, kArrM, gj, C(i));
                put(g, kRg2 + i, kArrM, gj + 2, C(i));
            }
            put(g, kE10, kArrM, gj + 1, gj);
            put(g, kD1, kArrM, gj + 1, gj + 1);
            put(g, kD0, kArrM, gj, gj);
            put(g, kD2, kArrM, gj + 2, gj + 2);
            for (int r = 0; r < 3; ++r) { put(g, kDqG + r, kArrDq, gj + r, 0); put(g, kBiasG + r, kArrBias, gj + r, 0); }
        }
        for (int cr = 0; cr < kd; ++cr)
            for (int i = 0; i < 7; ++i) put(g0 + 3, cr * 7 + i, kArrJ, R.row_arm[arm] + cr, C(i));
        for (int i = 0; i < 6; ++i) put(g0 + 3, kd * 7 + i, kArrBias, jb + i, 0);
        device_block(g0 + 4, R.dev_arm[arm], 0);
    }
    return true;
}

// Where the pack step finds every entry: array, element offset inside an instance's record, instance stride.
struct PackTable {
    int32_t n_entries, pad_;
    uint64_t base[kArrCount];               // 0: the array is absent (default max_vel / zero is used instead)
    int64_t stride[kArrCount];              // doubles between instances
    int32_t off[kMaxEntries];               // doubles
    int8_t arr[kMaxEntries];
    double mv_default[IRLOSC_MAX_DEVICES][2];
};

inline int32_t build_pack_table(const KParams &P, const KIo &io, const TileSpec &S, PackTable &T) {
    memset(&T, 0, sizeof T);
    T.n_entries = S.n_entries;
    const int D = P.D, n = P.n;
    auto set = [&](int a, const double *p, int64_t stride) { T.base[a] = (uint64_t)(uintptr_t)p; T.stride[a] = stride; };
    set(kArrM, io.M, io.m_stride);
    set(kArrJ, io.J, io.j_stride);
    set(kArrDq, io.dq, n);
    set(kArrBias, io.bias, n);
    set(kArrEeXyz, io.ee_xyz, 3 * D);
    set(kArrEeQuat, io.ee_quat, 4 * D);
    set(kArrTXyz, io.target_xyz, 3 * D);
    set(kArrTQuat, io.target_quat, 4 * D);
    set(kArrMaxVel, io.max_vel, 2 * D);
    set(kArrFtX, io.ft_xmat, 9 * D);
    set(kArrFtRaw, io.ft_raw, 6 * D);
    for (int d = 0; d < D; ++d) { T.mv_default[d][0] = P.dev[d].max_vel[0]; T.mv_default[d][1] = P.dev[d].max_vel[1]; }
    for (int e = 0; e < S.n_entries; ++e) {
        const irlosc_tile_entry &s = S.e[e];
        T.arr[e] = (int8_t)s.array;
        int64_t off = 0;
        switch (s.array) {
            case kArrM:
                if (io.m_layout == IRLOSC_M_PACKED) off = (int64_t)s.i * (s.i + 1) / 2 + s.j;
                else if (io.m_layout == IRLOSC_M_QM) {
                    off = qm_offset(P, s.i, s.j);
                    if (off < 0) return fail(IRLOSC_ERR_INVALID, "IRLOSC_M_QM: tile entry (%d, %d) is outside the kinematic tree", s.i, s.j);
                } else off = (int64_t)s.i * io.ldm + s.j;
                break;
            case kArrJ: {
                const int64_t r = io.j_layout == IRLOSC_J_ROWS ? s.i : (int64_t)P.row_dev[s.i] * 6 + P.row_comp[s.i];
                off = r * io.ldj + s.j;
                break;
            }
            case kArrDq: case kArrBias: off = s.i; break;
            case kArrEeXyz: case kArrTXyz: off = s.i * 3 + s.j; break;
            case kArrEeQuat: case kArrTQuat: off = s.i * 4 + s.j; break;
            case kArrMaxVel: off = s.i * 2 + s.j; break;
            case kArrFtX: off = s.i * 9 + s.j; break;
            case kArrFtRaw: off = s.i * 6 + s.j; break;
            default: break;
        }
        if (off > INT32_MAX) return fail(IRLOSC_ERR_INVALID, "record too large for the tile packer");
        T.off[e] = (int32_t)off;
    }
    return IRLOSC_OK;
}

// One tile entry of one instance from the caller's arrays (host and device).
IRLOSC_HD double pack_fetch(const PackTable &T, int e, int64_t inst) {
    const int a = T.arr[e];
    const double *base = reinterpret_cast<const double *>(T.base[a]);
    if (a == kArrPad) return 0.0;
    if (base == nullptr) return a == kArrMaxVel ? T.mv_default[T.off[e] >> 1][T.off[e] & 1] : 0.0;
    return base[inst * T.stride[a] + T.off[e]];
}

// ---------------------------------------------------------------- per-instance function (host + device)
template <int KD, bool HAS_BASE>
struct LaneState {
    static constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    static constexpr int KT = KD * (KD + 1) / 2;
    double akA[2][KT];
    double j0[K], dxr[K], g[K];
    double base_arm[2][6];
    double base_st, inv0;
    double *u_all_row, *ctrl_row;
    uint8_t *status;
    bool force_pinv;
};

// Original Jacobian entries, re-read from the tile (the lane's column: p points at entry 0 of this lane).
template <int KD, bool HAS_BASE, class LD>
struct JTile {
    const double *p;
    int rows[2], g0;                        // first entry of the arms' task-row groups, of G0
    LD ld;
    IRLOSC_HD double stand(int canon) const {
        if (canon < KD) return ld(p + (size_t)(rows[0] + canon * 7) * kTile);
        if (canon < 2 * KD) return ld(p + (size_t)(rows[1] + (canon - KD) * 7) * kTile);
        return ld(p + (size_t)(g0 + stream::kG0Jbase) * kTile);
    }
    IRLOSC_HD double arm(int am, int i, int cr) const { return ld(p + (size_t)(rows[am] + cr * 7 + i + 1) * kTile); }
};

// One instance; `group(g)` returns the reader of group g.  Returns true when the warp must finish the task-space solve.
template <int KD, bool HAS_BASE, class GROUPS, class JA>
IRLOSC_HD bool lane_instance(const KParams &P, const FRoles &R, const double *target_vel, GROUPS &group, const JA &ja,
                             LaneState<KD, HAS_BASE> &T, const Debug *dbg) {
    using namespace stream;
    constexpr int KT = KD * (KD + 1) / 2;
    const int D = P.D;
    const double gb = P.use_g ? 1.0 : 0.0;
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
    bool m_ok = true;
    double d0 = 0.0, uv_st = 0.0, bias0;
    {   // ---- G0: stand / base device
        auto rd = group(0);
        bias0 = rd(kG0Bias0);
        if (HAS_BASE) {
            const double jb0 = rd(kG0Jbase);
            T.j0[R.row_base] = jb0;
            T.dxr[R.row_base] = jb0;                 // times dq[0], known after the first arm group
            device_signal_staged(P, R.dev_base, rd, kG0Dev, true, T.g);
        }
    }
#pragma unroll 1
    for (int arm = 0; arm < 2; ++arm) {
        const int jb = 1 + 12 * arm;
        const int row_a = R.row_arm[arm];
        ArmState<KD> S;
        {
            auto rd = group(1 + 5 * arm);
            consume_cc<KD>(rd, arm, S);
        }
        if (HAS_BASE && arm == 0) T.dxr[R.row_base] *= S.dqC[0];
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            auto rd = group(2 + 5 * arm + half);
            m_ok = consume_grip<KD>(rd, P, R, vel_zero, gb, jb + 6 + 3 * half, S, T.u_all_row, T.ctrl_row, dbg) && m_ok;
        }
        double ak[KT], j0r[KD], dxa[KD], c0, uv0;
        {
            auto rd = group(4 + 5 * arm);
            m_ok = consume_rows<KD>(rd, P, R, vel_zero, gb, jb, S, ak, j0r, (double *)nullptr, dxa, (double (*)[KD]) nullptr,
                                    T.base_arm[arm], &c0, &uv0, dbg) && m_ok;
        }
        d0 += c0;
        uv_st += uv0;
#pragma unroll
        for (int cr = 0; cr < KD; ++cr) { T.j0[row_a + cr] = j0r[cr]; T.dxr[row_a + cr] = dxa[cr]; }
#pragma unroll
        for (int e = 0; e < KT; ++e) T.akA[arm][e] = ak[e];
        {
            auto rd = group(5 + 5 * arm);
            device_signal_staged(P, R.dev_arm[arm], rd, 0, true, T.g);
        }
    }
    m_ok = m_ok && (d0 > 0.0);
    T.inv0 = fused::rcp64(d0);
    T.base_st = fma(fused::coef_uv(P, R, vel_zero, 0), uv_st, gb * bias0);
    if (dbg && dbg->uv) dbg->uv[0] = uv_st;
    return fused::osc_tail<KD, HAS_BASE>(P, R, target_vel, vel_zero, 0, m_ok, T.akA, T.j0, T.dxr, T.g, ja, T.base_arm, T.base_st,
                                         T.inv0, T.u_all_row, T.ctrl_row, T.status, &T.force_pinv, dbg);
}

#if defined(__CUDACC__) && !defined(IRLOSC_FUSED_NO_KERNELS)
// ---------------------------------------------------------------- kernels
struct LaneArgs {
    const double *tiles;
    int32_t n_entries;                      // doubles per instance (E); a tile is 32 E doubles
    int32_t pf;                             // L2 prefetch distance in groups, 0 = off
    int32_t gbase[kGroups + 1];
    const double *target_vel;               // [B][D][6] plain array, optional
    double *u_all, *ctrl;
    uint8_t *status;
    int *sched;                             // ticket counter of the launch (pair kernel), null = static assignment
};

struct LdStream {       // streaming load: read once, keep out of L1
    __device__ __forceinline__ double operator()(const double *p) const {
        double v;
        asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
        return v;
    }
};
struct LdCached {       // second read of the Jacobian entries: served by L2 / L1
    __device__ __forceinline__ double operator()(const double *p) const { return __ldg(p); }
};

__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// STAGED = false: lanes load their entries straight from HBM / L2 into registers (LDG, one coalesced 256-byte request
//                 per entry), an optional bulk L2 prefetch runs A.pf groups ahead.
// STAGED = true : a group of a tile is ONE contiguous block of HBM, so one elected lane brings it into the warp's
//                 shared-memory ring with a single TMA bulk copy (cp.async.bulk + mbarrier complete_tx) NS - 1 groups
//                 ahead of its use; the lanes then read conflict-free LDS.64 ([entry][lane]).  ncu on the LDG form:
//                 47 % of the stall cycles are long_scoreboard (waiting for the loads of the group just started).
template <int KD, bool HAS_BASE, int NT, bool STAGED>
__global__ void __launch_bounds__(NT, 1)
osc_step_lane(const __grid_constant__ KParams P, const __grid_constant__ LaneArgs A, const int64_t B,
              const __grid_constant__ FRoles R, const __grid_constant__ stream::Gather G, const int warp_bytes,
              const int stage_bytes, const int n_stages) {
    extern __shared__ __align__(16) unsigned char lane_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = NT / 32;
    unsigned char *wbase = lane_smem + (size_t)warp * warp_bytes;
    // per warp: [ring of n_stages x stage_bytes (STAGED)][mbarriers][packed ctrl tile][warp-finish scratch]
    unsigned char *ring = wbase;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(wbase + (STAGED ? (size_t)n_stages * stage_bytes : 0));
    unsigned char *rest = reinterpret_cast<unsigned char *>(bars) + (STAGED ? 64 : 0);
    double *ctile = reinterpret_cast<double *>(rest);                                   // [32][n_ctrl]
    fused::WarpFix<KD, HAS_BASE> &wfix =
        *reinterpret_cast<fused::WarpFix<KD, HAS_BASE> *>(rest + ((32 * P.n_ctrl * 8 + 15) & ~15));
    const int64_t n_tiles = (B + kTile - 1) / kTile;
    const int64_t tile_doubles = (int64_t)A.n_entries * kTile;
    // warps of a CTA take neighbouring tiles, CTAs stride over the batch
    const int64_t gw = (int64_t)blockIdx.x * W + warp, gstride = (int64_t)gridDim.x * W;
    const int64_t my_tiles = gw < n_tiles ? (n_tiles - gw + gstride - 1) / gstride : 0;
    // ---- ring state (warp-uniform, plain 32-bit counters: no divisions on the per-group path)
    const uint32_t ring_u32 = tiled::smem_u32(ring), bars_u32 = tiled::smem_u32(bars);
    int cs = 0;                                          // stage the next group is consumed from
    uint32_t par = 0;                                    // bit s: parity of the next completion of stage s
    int is = 0, ig = 0;                                  // stage / group of the next copy to issue
    const double *itile = A.tiles + gw * tile_doubles;   // tile of the next copy to issue
    int64_t ileft = my_tiles * kGroups;                  // copies still to issue
    auto issue_next = [&]() {                            // lane 0 issues the bulk copy; every lane advances the counters
        if (ileft > 0) {
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)(A.gbase[ig + 1] - A.gbase[ig]) * (kTile * 8u);
                const uint32_t bar = bars_u32 + 8u * is;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(ring_u32 + (uint32_t)is * (uint32_t)stage_bytes), "l"(itile + (size_t)A.gbase[ig] * kTile), "r"(bytes),
                               "r"(bar) : "memory");
            }
            --ileft;
            if (++ig == kGroups) { ig = 0; itile += gstride * tile_doubles; }
            if (++is == n_stages) is = 0;
        }
    };
    if (STAGED) {
        if (lane == 0) {
            for (int s = 0; s < n_stages; ++s) tiled::mbar_init(&bars[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        for (int s = 0; s < n_stages - 1; ++s) issue_next();
    }
    for (int64_t k = 0; k < my_tiles; ++k) {
        const int64_t tile = gw + k * gstride;
        const double *tb = A.tiles + tile * tile_doubles;
        const double *tl = tb + lane;
        const int64_t inst = tile * kTile + lane;
        const bool valid = inst < B;
        const int64_t inst_c = valid ? inst : B - 1;
        if (!STAGED && A.pf > 0 && lane == 0) bulk_prefetch_l2(tb, (uint32_t)((A.gbase[A.pf < kGroups ? A.pf : kGroups]) * kTile * 8));
        auto group = [&](int g) {
            const double *p;
            if (STAGED) {
                __syncwarp();                            // every lane is done with the stage that is refilled now
                if (lane == 0) tiled::fence_proxy_async();
                issue_next();
                const uint32_t bar = bars_u32 + 8u * cs;
                const uint32_t parity = (par >> cs) & 1u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "WAIT_%=:\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                    "@p bra DONE_%=;\n\t"
                    "bra WAIT_%=;\n\t"
                    "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
                p = reinterpret_cast<const double *>(ring + (size_t)cs * stage_bytes) + lane;
                par ^= 1u << cs;
                if (++cs == n_stages) cs = 0;
            } else {
                if (A.pf > 0 && lane == 0 && g + A.pf < kGroups)
                    bulk_prefetch_l2(tb + (size_t)A.gbase[g + A.pf] * kTile,
                                     (uint32_t)((A.gbase[g + A.pf + 1] - A.gbase[g + A.pf]) * kTile * 8));
                p = tl + (size_t)A.gbase[g] * kTile;
            }
            return [p](int e) { return STAGED ? p[e * kTile] : LdStream{}(p + e * kTile); };
        };
        const JTile<KD, HAS_BASE, LdCached> ja{tl, {A.gbase[4], A.gbase[9]}, A.gbase[0], LdCached{}};
        LaneState<KD, HAS_BASE> T;
        T.u_all_row = (A.u_all && valid) ? A.u_all + inst * kN : nullptr;
        T.ctrl_row = ctile + lane * P.n_ctrl;
        T.status = (A.status && valid) ? A.status + inst : nullptr;
        const bool hard = lane_instance<KD, HAS_BASE>(P, R, A.target_vel ? A.target_vel + inst_c * P.D * 6 : nullptr, group, ja, T,
                                                     nullptr);
        fused::tail_warp_finish<KD, HAS_BASE>(
            wfix, hard && valid, lane,
            [&](double *rec) { fused::tail_record<KD, HAS_BASE>(R, T.akA, T.j0, T.g, ja, T.base_arm, T.base_st, T.inv0, T.force_pinv, rec); },
            [&](const double *rec, const double *w, int fl) {
                fused::fixup_finish<KD, HAS_BASE>(R, T.u_all_row, T.ctrl_row, rec, w, 0, 1);
                if (T.status) *T.status = (uint8_t)(*T.status | fl);
            });
        __syncwarp();
        stream::write_ctrl_tile(ctile, A.ctrl, G, P.n_ctrl, tile, B, lane);
        __syncwarp();
    }
}

// arrays of irlosc_io -> tiles.  A CTA per tile; a warp writes one entry of the tile per iteration (one coalesced
// 256-byte row), reading it from the 32 instances' records.
static __global__ void __launch_bounds__(256)      // (static: this header is part of two translation units)
pack_tiles_kernel(const __grid_constant__ PackTable T, double *tiles, const int64_t B) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int64_t n_tiles = (B + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int64_t inst = tile * kTile + lane;
        inst = inst < B ? inst : B - 1;             // ragged tile: padding lanes repeat the last instance
        double *dst = tiles + tile * (int64_t)T.n_entries * kTile + lane;
        for (int e = warp; e < T.n_entries; e += W) dst[(size_t)e * kTile] = pack_fetch(T, e, inst);
    }
}
#endif

}  // namespace lane
}  // namespace irlosc

// Streaming DualUR5 OSC step (sm_100a): one THREAD per robot instance, state read from HBM.
//
// osc_tree.cuh spends 4 lanes on an instance, stages whole 4.9 KB records and needs ~460 warp
// instructions per instance (shuffles, duplicated phases, 25 % lane use in the task law); ncu showed
// it latency / instruction-delivery bound at a third of the HBM roofline.  Here a warp owns 32
// instances and every lane runs the complete scalar elimination of ONE instance, so there are no
// shuffles and no idle lanes (~90 warp instructions per instance).  What makes that possible is
// the staging: a thread walking its own 4.9 KB record would be an uncoalesced access pattern, so
// the warp copies the record cooperatively in the order the elimination consumes it:
//
//   * the host flattens the controller + array layouts into a PLAN: 11 groups per 32-instance
//     tile (stand/base data; per arm: M over [stand, arm joints] + dq, gripper half 0, gripper
//     half 1, task rows of J + bias, the device's poses/targets), each a list of 8-entry chunks
//     with per-lane source offsets.  Only the entries the kinematic tree makes non-zero are ever
//     copied (155 of 325 for M, 43 of 175 for J): structural zeros cost neither instructions nor,
//     where a 32-byte sector holds nothing else, HBM traffic;
//   * a chunk is moved by 8 cp.async (LDGSTS, 8 bytes per lane): 8 consecutive lanes read 8
//     neighbouring doubles of one instance, 4 instances per instruction, straight into shared
//     memory in [entry][instance] order (row pitch 33 doubles: conflict-free on both sides);
//   * each warp double-buffers groups (cp.async.commit_group / wait_group): while the lanes
//     eliminate group g from registers, group g + 1 is in flight.  No CTA-wide barriers at all.
//
// The arithmetic is the tree-sparse elimination of osc_tree.cuh (same order of pivots: gripper
// leaves, arm joints 6..1 with the arm's task rows, stand joint), then osc_tail.cuh: dense k x k
// LDL^T in registers, inverse-vs-pinv certificate, deferred eigen fix-up, assembly, packing.
// Packed ctrl rows are staged per warp and written as 16-byte vectors (local array, peer arrays or
// the NVSwitch multicast mapping, like osc_tree.cuh).
//
// Reference restated: ir-lab/irl_control osc.py:41-68, 120-210; robot.py:44-72; device.py:115-170.
#pragma once
#include "irlosc_device.cuh"
#include "osc_fused_types.h"
#include "osc_tail.cuh"

namespace irlosc {
namespace stream {

using fused::FRoles;
using fused::Debug;
using fused::kN;

constexpr int kGroups = 11;             // G0, then per arm: CC, half 0, half 1, task rows, device data
constexpr int kMaxChunks = 72;
constexpr int kPitch = 33;              // doubles between entries of a stage ([entry][instance])

struct Chunk {
    uint64_t base;                      // array base pointer
    int32_t stride;                     // bytes between instances
    int32_t pad_;
    int32_t off[8];                     // per lane (lane & 7): byte offset inside the instance's record
    int16_t dst[8];                     // per lane: stage entry the double goes to
};
constexpr int kMaxArrays = 12;
struct Plan {
    int32_t n_chunks, stage_entries;    // stage_entries includes the trash entry padding lanes write to
    int32_t first[kGroups + 1];         // chunk range of every group
    int32_t has_mvel, n_arrays;
    uint64_t arr_base[kMaxArrays];      // the input arrays (for the whole-tile L2 prefetch)
    int32_t arr_stride[kMaxArrays];     // bytes between instances
    Chunk ch[kMaxChunks];
};

// entry layouts ----------------------------------------------------------------------------------
// device data block (G0 from kDev0, per arm group 4 from 0)
constexpr int kEeXyz = 0, kEeQuat = 3, kTXyz = 7, kTQuat = 10, kMaxVel = 14, kFtX = 16, kFtRaw = 25, kDevEntries = 31;
// G0
constexpr int kG0Jbase = 0, kG0Bias0 = 1, kG0Dev = 2, kG0Entries = kG0Dev + 16;
// arm group 0: c[28] then dqC[7]
constexpr int kCC = 0, kDqC = 28, kG1Entries = 35;
// gripper half: rows g1 (7 + e10 + d), g0 (7 + d), g2 (7 + d), dq x3, bias x3
constexpr int kRg1 = 0, kE10 = 7, kD1 = 8, kRg0 = 9, kD0 = 16, kRg2 = 17, kD2 = 24, kDqG = 25, kBiasG = 28, kGripEntries = 31;

struct Outputs {
    double *u_all, *ctrl;
    uint8_t *status;
    const double *target_vel;
};

// ---------------------------------------------------------------- per-instance consumers (host + device)
template <int KD>
struct ArmState {
    double c[28], uvC[7], dqC[7];
};

// osc.py:159-168,179-181 for one device from staged values: the unmasked six task components.
template <class RD>
IRLOSC_HD void device_u6_staged(const KParams &P, int d, const RD &rd, int e0, bool has_mvel, double *u6) {
    const KDevice &dv = P.dev[d];
    double ee_p[3], ee_q[4], txyz[3], tquat[4];
#pragma unroll
    for (int i = 0; i < 3; ++i) { ee_p[i] = rd(e0 + kEeXyz + i); txyz[i] = rd(e0 + kTXyz + i); }
#pragma unroll
    for (int i = 0; i < 4; ++i) { ee_q[i] = rd(e0 + kEeQuat + i); tquat[i] = rd(e0 + kTQuat + i); }
    double mv[2] = {dv.max_vel[0], dv.max_vel[1]};
    if (has_mvel) { mv[0] = rd(e0 + kMaxVel); mv[1] = rd(e0 + kMaxVel + 1); }
    bool oob = false;
    device_task_signal(dv, ee_p, ee_q, txyz, tquat, nullptr, mv, nullptr, 0, u6, &oob);
    if (P.admittance) {
        double Rm[9], raw[6], ft[6];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rm[i] = rd(e0 + kFtX + i);
#pragma unroll
        for (int i = 0; i < 6; ++i) raw[i] = rd(e0 + kFtRaw + i);
        rotate_wrench(Rm, raw, ft);
#pragma unroll
        for (int i = 0; i < 6; ++i) u6[i] += ft[i];
    }
}
// ... masked by the device's controlled DoF into its task rows of gpre.
template <class RD>
IRLOSC_HD void device_signal_staged(const KParams &P, int d, const RD &rd, int e0, bool has_mvel, double *gpre) {
    const KDevice &dv = P.dev[d];
    double u6[6];
    device_u6_staged(P, d, rd, e0, has_mvel, u6);
    int r = dv.row0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (dv.dof[i]) gpre[r++] = u6[i];
}

// arm group 0: M over [stand, arm joints] and the matching dq; uv = M dq on that block.
template <int KD, class RD>
IRLOSC_HD void consume_cc(const RD &rd, int arm, ArmState<KD> &S) {
#pragma unroll
    for (int i = 0; i < 7; ++i) { S.dqC[i] = rd(kDqC + i); S.uvC[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double v = rd(kCC + i * (i + 1) / 2 + j);
            if (i == 0 && arm != 0) v = 0.0;               // M[0][0] enters once
            S.c[i * (i + 1) / 2 + j] = v;
            S.uvC[i] = fma(v, S.dqC[j], S.uvC[i]);
            if (j != i) S.uvC[j] = fma(v, S.dqC[i], S.uvC[j]);
        }
}

// gripper half: joints g0 (child of arm joint 6), g1 (child of g0), g2 (child of arm joint 6).
template <int KD, class RD>
IRLOSC_HD bool consume_grip(const RD &rd, const KParams &P, const FRoles &R, unsigned vel_zero, double gb, int gj,
                            ArmState<KD> &S, double *u_all_row, double *ctrl_row, const Debug *dbg) {
    bool ok = true;
    double rg0[7], rg1[7], rg2[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) { rg1[i] = rd(kRg1 + i); rg0[i] = rd(kRg0 + i); rg2[i] = rd(kRg2 + i); }
    const double e10 = rd(kE10);
    double d1 = rd(kD1), d0 = rd(kD0), d2 = rd(kD2);
    const double q0 = rd(kDqG), q1 = rd(kDqG + 1), q2 = rd(kDqG + 2);
    double uv0 = fma(e10, q1, d0 * q0), uv1 = fma(e10, q0, d1 * q1), uv2 = d2 * q2;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        uv0 = fma(rg0[i], S.dqC[i], uv0);
        uv1 = fma(rg1[i], S.dqC[i], uv1);
        uv2 = fma(rg2[i], S.dqC[i], uv2);
        S.uvC[i] = fma(rg0[i], q0, fma(rg1[i], q1, fma(rg2[i], q2, S.uvC[i])));
    }
    double cg[3];
    fused::coef_uv_group<3>(P, R, vel_zero, gj, cg);
    fused::put_joint(R, u_all_row, ctrl_row, gj, fma(cg[0], uv0, gb * rd(kBiasG)));
    fused::put_joint(R, u_all_row, ctrl_row, gj + 1, fma(cg[1], uv1, gb * rd(kBiasG + 1)));
    fused::put_joint(R, u_all_row, ctrl_row, gj + 2, fma(cg[2], uv2, gb * rd(kBiasG + 2)));
    if (dbg && dbg->uv) { dbg->uv[gj] = uv0; dbg->uv[gj + 1] = uv1; dbg->uv[gj + 2] = uv2; }
    {   // eliminate g1 (leaf): touches the C block, row g0 and pivot g0
        ok = ok && (d1 > 0.0);
        const double inv = fused::rcp64(d1);
        const double t7 = e10 * inv;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const double ti = rg1[i] * inv;
#pragma unroll
            for (int j = 0; j <= i; ++j) S.c[i * (i + 1) / 2 + j] = fma(-ti, rg1[j], S.c[i * (i + 1) / 2 + j]);
            rg0[i] = fma(-rg1[i], t7, rg0[i]);
        }
        d0 = fma(-e10, t7, d0);
    }
    {
        ok = ok && (d0 > 0.0);
        const double inv = fused::rcp64(d0);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const double ti = rg0[i] * inv;
#pragma unroll
            for (int j = 0; j <= i; ++j) S.c[i * (i + 1) / 2 + j] = fma(-ti, rg0[j], S.c[i * (i + 1) / 2 + j]);
        }
    }
    {
        ok = ok && (d2 > 0.0);
        const double inv = fused::rcp64(d2);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const double ti = rg2[i] * inv;
#pragma unroll
            for (int j = 0; j <= i; ++j) S.c[i * (i + 1) / 2 + j] = fma(-ti, rg2[j], S.c[i * (i + 1) / 2 + j]);
        }
    }
    return ok;
}

// arm task rows: dx, original J entries, then arm joints 6..1 are eliminated together with the rows.
// Leaves: ak (the arm's block of A), j0r (stand column of the reduced rows), c0 (the arm's part of the
// stand pivot), uv0 (its part of (M dq)_stand).
template <int KD, class RD>
IRLOSC_HD bool consume_rows(const RD &rd, const KParams &P, const FRoles &R, unsigned vel_zero, double gb, int jb, ArmState<KD> &S,
                            double *ak, double *j0r, double *jst, double *dxr, double (*jarm)[KD], double *base_arm,
                            double *c0, double *uv0, const Debug *dbg) {
    bool ok = true;
    double jr[KD][7];
#pragma unroll
    for (int cr = 0; cr < KD; ++cr) {
        double dx = 0.0;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const double v = rd(cr * 7 + i);
            jr[cr][i] = v;
            dx = fma(v, S.dqC[i], dx);
            if (jarm != nullptr && i > 0) jarm[i - 1][cr] = v;
        }
        dxr[cr] = dx;
        if (jst != nullptr) jst[cr] = jr[cr][0];
    }
    double ca[6];
    fused::coef_uv_group<6>(P, R, vel_zero, jb, ca);
#pragma unroll
    for (int i = 1; i < 7; ++i) {
        base_arm[i - 1] = fma(ca[i - 1], S.uvC[i], gb * rd(KD * 7 + i - 1));
        if (dbg && dbg->uv) dbg->uv[jb + i - 1] = S.uvC[i];
    }
    *uv0 = S.uvC[0];
#pragma unroll
    for (int e = 0; e < KD * (KD + 1) / 2; ++e) ak[e] = 0.0;
#pragma unroll
    for (int kk = 6; kk >= 1; --kk) {
        const double d = S.c[kk * (kk + 1) / 2 + kk];
        ok = ok && (d > 0.0);
        const double inv = fused::rcp64(d);
        double t[6];
#pragma unroll
        for (int j = 0; j < kk; ++j) t[j] = S.c[kk * (kk + 1) / 2 + j] * inv;
#pragma unroll
        for (int i = 0; i < kk; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j)
                S.c[i * (i + 1) / 2 + j] = fma(-S.c[kk * (kk + 1) / 2 + i], t[j], S.c[i * (i + 1) / 2 + j]);
#pragma unroll
        for (int cr = 0; cr < KD; ++cr) {
            const double jk = jr[cr][kk];
            const double tc = jk * inv;
#pragma unroll
            for (int c2 = cr; c2 < KD; ++c2) ak[c2 * (c2 + 1) / 2 + cr] = fma(jr[c2][kk], tc, ak[c2 * (c2 + 1) / 2 + cr]);
#pragma unroll
            for (int j = 0; j < kk; ++j) jr[cr][j] = fma(-jk, t[j], jr[cr][j]);
        }
    }
#pragma unroll
    for (int cr = 0; cr < KD; ++cr) j0r[cr] = jr[cr][0];
    *c0 = S.c[0];
    return ok;
}

// One instance, all groups available through `group(g)` (returns the reader of group g; on the device
// this is where the warp waits for the copy and issues the next one).
// Returns true when the task-space solve must be finished by the warp (state_warp_finish on T).
template <int KD, bool HAS_BASE, class GROUPS>
IRLOSC_HD bool stream_instance(const KParams &P, const FRoles &R, const Plan &plan, const Outputs &out, int64_t inst,
                               GROUPS &group, double *ctrl_row, fused::TailState<KD, HAS_BASE> &T, const Debug *dbg) {
    constexpr int KT = KD * (KD + 1) / 2;
    const int D = P.D;
    const double gb = P.use_g ? 1.0 : 0.0;
    const bool has_mvel = plan.has_mvel != 0;
    double *u_all_row = out.u_all ? out.u_all + inst * kN : nullptr;
    unsigned vel_zero = 0;
#pragma unroll
    for (int d = 0; d < IRLOSC_MAX_DEVICES; ++d) {
        bool tracking = false;
        if (d < D && out.target_vel != nullptr) {
            tracking = true;
            for (int i = 0; i < 6; ++i) tracking = tracking && (out.target_vel[(inst * D + d) * 6 + i] != 0.0);
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
            T.jst[R.row_base] = jb0;
            T.dxr[R.row_base] = jb0;                 // times dq[0], known after the first arm group
            device_signal_staged(P, R.dev_base, rd, kG0Dev, has_mvel, T.g);
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
            m_ok = consume_grip<KD>(rd, P, R, vel_zero, gb, jb + 6 + 3 * half, S, u_all_row, ctrl_row, dbg) && m_ok;
        }
        double ak[KT], j0r[KD], jstr[KD], dxa[KD], c0, uv0;
        {
            auto rd = group(4 + 5 * arm);
            m_ok = consume_rows<KD>(rd, P, R, vel_zero, gb, jb, S, ak, j0r, jstr, dxa, T.jarm[arm], T.base_arm[arm], &c0, &uv0, dbg) && m_ok;
        }
        d0 += c0;
        uv_st += uv0;
#pragma unroll
        for (int cr = 0; cr < KD; ++cr) { T.j0[row_a + cr] = j0r[cr]; T.jst[row_a + cr] = jstr[cr]; T.dxr[row_a + cr] = dxa[cr]; }
#pragma unroll
        for (int e = 0; e < KT; ++e) T.akA[arm][e] = ak[e];
        {
            auto rd = group(5 + 5 * arm);
            device_signal_staged(P, R.dev_arm[arm], rd, 0, has_mvel, T.g);
        }
    }
    m_ok = m_ok && (d0 > 0.0);
    T.inv0 = fused::rcp64(d0);
    T.base_st = fma(fused::coef_uv(P, R, vel_zero, 0), uv_st, gb * bias0);
    if (dbg && dbg->uv) dbg->uv[0] = uv_st;
    T.u_all_row = u_all_row;
    T.ctrl_row = ctrl_row;
    T.status = out.status ? out.status + inst : nullptr;
    return fused::state_tail<KD, HAS_BASE>(P, R, out.target_vel ? out.target_vel + inst * D * 6 : nullptr, vel_zero, 0, m_ok, T, dbg);
}

#if defined(__CUDACC__) && !defined(IRLOSC_FUSED_NO_KERNELS)
// ---------------------------------------------------------------- kernel
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8_l2(uint32_t dst, const void *src) {      // L2 fetches the whole 128-byte line
    asm volatile("cp.async.ca.shared.global.L2::128B [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Gather {       // fused result gather, same contract as KIo
    int32_t n_gather, ctrl_vec;
    int64_t gather_offset;
    double *ctrl_gather[IRLOSC_MAX_PEERS];
    double *ctrl_mc;
};

// Packed ctrl rows of a 32-instance tile are contiguous in every destination: local array, peer-mapped gathered
// arrays or the NVSwitch multicast mapping (16-byte vectors when everything is aligned).
__device__ __forceinline__ void write_ctrl_tile(const double *ctile, double *ctrl, const Gather &G, int n_ctrl, int64_t tile,
                                                int64_t B, int lane) {
    const int64_t row0 = tile * 32 * (int64_t)n_ctrl;
    const int n_valid = (int)((B - tile * 32) < 32 ? (B - tile * 32) : 32);
    if (n_valid == 32 && G.ctrl_vec) {
        const double2 *src = reinterpret_cast<const double2 *>(ctile);
        double2 *dst = reinterpret_cast<double2 *>(ctrl + row0);
        for (int e = lane; e < 16 * n_ctrl; e += 32) {
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
osc_step_stream(const __grid_constant__ KParams P, const __grid_constant__ Plan plan_in, const __grid_constant__ Outputs out,
                const int64_t B, const __grid_constant__ FRoles R,
                const __grid_constant__ Gather G, const int stage_bytes, const int warp_bytes, const int mode) {
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Plan &plan = *reinterpret_cast<Plan *>(smem_raw);
    {   // plan tables to shared memory: per-lane offsets are read with lane-dependent indices
        const int *src = reinterpret_cast<const int *>(&plan_in);
        int *dst = reinterpret_cast<int *>(smem_raw);
        for (int i = threadIdx.x; i < (int)(sizeof(Plan) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int l8 = lane & 7, sub = lane >> 3;
    unsigned char *wbase = smem_raw + ((sizeof(Plan) + 15) & ~size_t(15)) + (size_t)warp * warp_bytes;
    double *ctile = reinterpret_cast<double *>(wbase + 2 * (size_t)stage_bytes);      // [32][n_ctrl]
    // scratch of the warp-cooperative finish (osc_tail.cuh), behind the packed ctrl tile
    fused::WarpFix<KD, HAS_BASE> &wfix =
        *reinterpret_cast<fused::WarpFix<KD, HAS_BASE> *>(wbase + 2 * (size_t)stage_bytes + ((32 * P.n_ctrl * 8 + 15) & ~15));
    const uint32_t stage_u32 = (uint32_t)__cvta_generic_to_shared(wbase);

    const int64_t n_tiles = (B + 31) / 32;
    const int64_t gw = (int64_t)blockIdx.x * W + warp, gstride = (int64_t)gridDim.x * W;
    const int64_t my_tiles = gw < n_tiles ? (n_tiles - gw + gstride - 1) / gstride : 0;
    const int64_t total = my_tiles * kGroups;
    int64_t seq_issue = 0, seq_wait = 0;

    // mode bit 1: a tile's records are contiguous byte ranges of every input array - ask L2 for them as
    // whole 128-byte lines one tile ahead, so that DRAM sees long sequential reads and the scattered
    // 8-byte gathers below hit L2
    auto prefetch_tile = [&](int64_t tile) {
        if (!(mode & 2) || tile >= n_tiles) return;
        for (int a = 0; a < plan.n_arrays; ++a) {
            const unsigned char *p0 = reinterpret_cast<const unsigned char *>(plan.arr_base[a]) + tile * 32 * (int64_t)plan.arr_stride[a];
            int64_t n_inst = B - tile * 32;
            n_inst = n_inst < 32 ? n_inst : 32;
            const int64_t bytes = n_inst * plan.arr_stride[a];
            const unsigned char *lo = reinterpret_cast<const unsigned char *>(reinterpret_cast<uintptr_t>(p0) & ~uintptr_t(127));
            for (const unsigned char *p = lo + lane * 128; p < p0 + bytes; p += 32 * 128) prefetch_l2(p);
        }
    };
    auto issue_next = [&]() {
        if (seq_issue < total) {
            const int64_t tile = gw + (seq_issue / kGroups) * gstride;
            const int g = (int)(seq_issue % kGroups);
            if (g == 0) prefetch_tile(tile + gstride);
            const uint32_t st = stage_u32 + (uint32_t)(seq_issue & 1) * (uint32_t)stage_bytes;
            const int64_t i0 = tile * 32 + sub;
            const bool full = (tile + 1) * 32 <= B;
            for (int c = plan.first[g]; c < plan.first[g + 1]; ++c) {
                const Chunk &ch = plan.ch[c];
                const int64_t stride = ch.stride;
                const unsigned char *src = reinterpret_cast<const unsigned char *>(ch.base) + ch.off[l8];
                const uint32_t dst = st + (uint32_t)(ch.dst[l8] * kPitch + sub) * 8u;
                if (full && (mode & 1)) {
                    const unsigned char *s0 = src + i0 * stride;
#pragma unroll
                    for (int t = 0; t < 8; ++t) cp_async8_l2(dst + t * 32, s0 + (int64_t)(4 * t) * stride);
                } else if (full) {
                    const unsigned char *s0 = src + i0 * stride;
#pragma unroll
                    for (int t = 0; t < 8; ++t) cp_async8(dst + t * 32, s0 + (int64_t)(4 * t) * stride);
                } else {
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        int64_t ii = i0 + 4 * t;
                        ii = ii < B ? ii : B - 1;          // ragged tile: re-read the last instance
                        cp_async8(dst + t * 32, src + ii * stride);
                    }
                }
            }
        }
        cp_async_commit();
        ++seq_issue;
    };
    prefetch_tile(gw);
    issue_next();
    for (int64_t k = 0; k < my_tiles; ++k) {
        const int64_t tile = gw + k * gstride;
        const int64_t inst = tile * 32 + lane;
        const bool valid = inst < B;
        const int64_t inst_c = valid ? inst : B - 1;
        auto group = [&](int) {
            __syncwarp();                   // everyone is done with the stage the next copy overwrites
            issue_next();
            cp_async_wait<1>();
            __syncwarp();
            const double *st = reinterpret_cast<const double *>(wbase + (size_t)(seq_wait & 1) * stage_bytes) + lane;
            ++seq_wait;
            return [st](int e) { return st[e * kPitch]; };
        };
        Outputs o = out;
        if (!valid) { o.u_all = nullptr; o.status = nullptr; }
        fused::TailState<KD, HAS_BASE> T;
        const bool hard = stream_instance<KD, HAS_BASE>(P, R, plan, o, inst_c, group, ctile + lane * P.n_ctrl, T, nullptr);
        fused::state_warp_finish<KD, HAS_BASE>(wfix, R, T, hard && valid, lane);
        __syncwarp();
        write_ctrl_tile(ctile, out.ctrl, G, P.n_ctrl, tile, B, lane);
        __syncwarp();
    }
    cp_async_wait<0>();
    (void)FULL;
}
#endif

}  // namespace stream
}  // namespace irlosc

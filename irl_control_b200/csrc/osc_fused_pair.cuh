// Fused state provider + OSC step with TWO lanes per instance, one per arm (sm_100a).
//
// osc_step_fused gives an instance to one thread: 255 registers, ~1 M local-memory instructions per launch, two
// arms walked one after the other.  The arms are independent subtrees of the stand joint: here a lane runs
// fused_arm (osc_fused.cuh: forward kinematics, momenta, gripper leaves, articulated sweep with the arm's task rows)
// on ITS arm, the pair adds the arms' subtree sums at the stand (momenta, bias wrenches, articulated inertia: 33
// shuffles), and the task-space solve and joint-space assembly are pair_tail of osc_pair.cuh - the same code the
// tile pair kernel ends with, the Jacobian entries coming from registers instead of the tile.
// A warp owns 16 instances; lane = 2 * (instance in the warp) + arm.  (The action-sequence instantiations stay on the
// thread-per-instance kernel.)
//
// Reference restated: ir-lab/irl_control osc.py:41-68, 120-210; robot.py:44-72; device.py:87-170.
#pragma once
#include "osc_fused.cuh"
#include "osc_pair.cuh"

#if defined(__CUDACC__) && !defined(IRLOSC_FUSED_NO_KERNELS)
namespace irlosc {
namespace fused {

template <int KD>
struct PairJRegs {                     // J[row cr of the lane's arm][0: stand, 1..6: the arm's joints]
    const double *jsta;
    const double (*jarm_a)[KD];
    __device__ __forceinline__ double operator()(int cr, int i) const { return i == 0 ? jsta[cr] : jarm_a[i - 1][cr]; }
};

template <int KD, bool HAS_BASE, int NT>
__global__ void __launch_bounds__(NT, 1)
osc_step_fused_pair(const __grid_constant__ KParams P, const __grid_constant__ KModel Mdl, const __grid_constant__ FIo io,
                    const int64_t B, const __grid_constant__ FRoles R) {
    constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    constexpr int N = kN;
    const int W = blockDim.x >> 5;                       // <= NT / 32
    extern __shared__ __align__(16) double fused_pair_smem[];
    const Scratch scr{fused_pair_smem + threadIdx.x, (int)blockDim.x};
    WarpFix<KD, HAS_BASE> &wfix = reinterpret_cast<WarpFix<KD, HAS_BASE> *>(fused_pair_smem + kScratchDoubles * blockDim.x)[threadIdx.x >> 5];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int arm = lane & 1, li = lane >> 1;
    const int D = P.D;
    const double gb = P.use_g ? 1.0 : 0.0;
    pair::PairSys<KD, HAS_BASE> S;
    S.duo.mask = 3u << (lane & ~1);
    S.duo.arm = arm;
    const pair::Duo duo = S.duo;
    const int64_t n_half = (B + pair::kHalf - 1) / pair::kHalf;
    const int64_t gw = (int64_t)blockIdx.x * W + warp, gstride = (int64_t)gridDim.x * W;
    for (int64_t ht = gw; ht < n_half; ht += gstride) {
        // lanes past the end of the batch repeat its last instance (same values to the same addresses)
        const int64_t at = ht * pair::kHalf + li, inst = at < B ? at : B - 1;
        const double *q = io.q + inst * N, *dq = io.dq + inst * N;
        const double *target_vel = io.target_vel ? io.target_vel + inst * D * 6 : nullptr;
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
        // ---- stand joint (both lanes)
        Body S0;
        double s0[6], IA0[21], Htot[6], FBtot[6];
        {
            Body Wd;
#pragma unroll
            for (int e = 0; e < 9; ++e) Wd.R[e] = (e % 4 == 0) ? 1.0 : 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) { Wd.o[i] = 0.0; Wd.v[i] = Wd.v[3 + i] = 0.0; Wd.a[i] = 0.0; Wd.a[3 + i] = -Mdl.gravity[i]; }
            double r0[3], Iw0[6];
            joint_down(Mdl.stand, Wd, q[0], dq[0], S0, s0, r0, Iw0);
            body_wrench(Mdl.stand.mass, r0, Iw0, S0.v, S0.a, Htot, FBtot);
#pragma unroll
            for (int e = 0; e < 21; ++e) IA0[e] = 0.0;
            add_rigid(IA0, Mdl.stand.mass, r0, Iw0);
        }
        double g[K], jb0 = 0.0, dxb = 0.0;
        if (HAS_BASE) {
            const int d = R.dev_base;
            const KFrame &F = Mdl.ee[d];
            double t[3], Re[9], p[3], eq[4];
            mat3_vec(S0.R, F.pos, t);
#pragma unroll
            for (int i = 0; i < 3; ++i) p[i] = S0.o[i] + t[i];
            mat3_mul(S0.R, F.R, Re);
            mat_to_quat(Re, eq);
            device_signal_early(P, io, inst, d, p, eq, Re, false, -1.0, g);
            double e6[6];
            task_force(P.row_comp[R.row_base], p, e6);
            jb0 = dot6(s0, e6);
            dxb = dot6(e6, S0.v);
        }
        // ---- the lane's arm
        double jsta[KD], dxa[KD], jarm_a[6][KD], base_a[6], Hl[6], FBl[6], IAl[21];
#pragma unroll
        for (int e = 0; e < 6; ++e) { Hl[e] = 0.0; FBl[e] = 0.0; }
#pragma unroll
        for (int e = 0; e < 21; ++e) IAl[e] = 0.0;
        SeqResult seq{false, 0.0};
        bool m_ok = fused_arm<KD, false>(P, Mdl, R, io, inst, scr, arm, vel_zero, gb, S0, s0, q, dq, g, S.D, S.v, jsta, dxa, jarm_a,
                                         base_a, Hl, FBl, IAl, seq, nullptr, nullptr);
        // ---- the stand joint couples the arms
#pragma unroll
        for (int e = 0; e < 6; ++e) { Htot[e] += duo.sum(Hl[e]); FBtot[e] += duo.sum(FBl[e]); }
#pragma unroll
        for (int e = 0; e < 21; ++e) IA0[e] += duo.sum(IAl[e]);
        double f0[6], inv0;
        const bool ok0 = joint_up(IA0, s0, f0, &inv0);
        m_ok = duo.all(m_ok) && ok0;
        const double base_st = fma(coef_uv(P, R, vel_zero, 0), dot6(s0, Htot), gb * dot6(s0, FBtot));
        S.inv0 = inv0;
        S.d0 = rcp64(inv0);
        S.vb = jb0;
        double gc[KD], gcb = 0.0;
        {
            const int row_a = R.row_arm[arm];
#pragma unroll
            for (int cr = 0; cr < KD; ++cr) gc[cr] = g[row_a + cr];
            if (HAS_BASE) gcb = g[R.row_base];
        }
        const PairJRegs<KD> ja{jsta, jarm_a};
        pair::pair_tail<KD, HAS_BASE>(P, R, S, gc, gcb, dxa, dxb, base_a, base_st, m_ok, ja, target_vel, vel_zero, 0,
                                      io.u_all ? io.u_all + inst * N : nullptr, io.ctrl + inst * P.n_ctrl,
                                      io.status ? io.status + inst : nullptr, true, wfix, lane);
        __syncwarp();
    }
}

}  // namespace fused
}  // namespace irlosc
#endif

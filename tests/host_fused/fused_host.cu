// TEST INFRASTRUCTURE ONLY.  Runs the per-instance function of the fused CUDA kernel
// (irl_control_b200/csrc/osc_fused.cuh, __host__ __device__) on the CPU so that the CPU test
// suite can check the kernel's arithmetic against the oracle without a GPU.  Nothing in the
// package loads this library; the product path is libirlosc.so on a GPU.
#define IRLOSC_FUSED_NO_KERNELS 1
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../irl_control_b200/csrc/irlosc_internal.h"
#include "../../irl_control_b200/csrc/irlosc_build.h"
#include "../../irl_control_b200/csrc/osc_fused.cuh"
#include "../../irl_control_b200/csrc/osc_stream.cuh"

static std::string g_err;
int32_t irlosc::fail(int32_t rc, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return rc;
}
int32_t irlosc::ensure_cap(Staging &, int, size_t) { return IRLOSC_OK; }

using namespace irlosc;
using namespace irlosc::fused;

// optional tap: which path resolved the task-space solve of every instance (TailHow bits), set by the test
static int *g_how = nullptr;
extern "C" void host_set_how(int *how) { g_how = how; }

// serial cyclic Jacobi + the truncation rule of tiled::eigen_solve (osc.py:52-55), test-only
template <int K>
static int host_eigen_solve(const double *A0, const double *g, bool force_pinv, double *w) {
    double A[K][K], V[K][K];
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) { A[i][j] = A0[i * K + j]; V[i][j] = i == j; }
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0, dia = 0;
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) (i == j ? dia : off) += A[i][j] * A[i][j];
        if (off <= 1e-30 * dia) break;
        for (int p = 0; p < K; ++p)
            for (int q = p + 1; q < K; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int i = 0; i < K; ++i) {
                    const double aip = A[i][p], aiq = A[i][q];
                    A[i][p] = c * aip - s * aiq; A[i][q] = s * aip + c * aiq;
                    const double vip = V[i][p], viq = V[i][q];
                    V[i][p] = c * vip - s * viq; V[i][q] = s * vip + c * viq;
                }
                for (int j = 0; j < K; ++j) {
                    const double apj = A[p][j], aqj = A[q][j];
                    A[p][j] = c * apj - s * aqj; A[q][j] = s * apj + c * aqj;
                }
            }
    }
    double lmax = 0, det = 1;
    for (int i = 0; i < K; ++i) { lmax = fmax(lmax, fabs(A[i][i])); det *= A[i][i]; }
    const bool pinv = force_pinv || !(fabs(det) >= kDetThreshold);
    double c[K];
    for (int e = 0; e < K; ++e) {
        double proj = 0;
        for (int i = 0; i < K; ++i) proj += V[i][e] * g[i];
        const bool keep = pinv ? (fabs(A[e][e]) > kPinvRcond * lmax) : true;
        c[e] = keep ? proj / A[e][e] : 0.0;
    }
    for (int i = 0; i < K; ++i) {
        double acc = 0;
        for (int e = 0; e < K; ++e) acc += V[i][e] * c[e];
        w[i] = acc;
    }
    return IRLOSC_ST_EIGEN | (pinv ? IRLOSC_ST_PINV : 0);
}

template <int KD, bool HB>
static int64_t run(const KParams &P, const KModel &M, const FRoles &R, const FIo &io, int64_t B, const Debug *dbg0,
                   const KSeq *Q = nullptr) {
    using RC = Rec<KD, HB>;
    constexpr int K = RC::K;
    std::vector<double> scratch(kScratchDoubles), rec(RC::SIZE);
    int64_t n_hard = 0;
    for (int64_t i = 0; i < B; ++i) {
        Debug d, *dp = nullptr;
        if (dbg0) {
            d = *dbg0;
            if (d.A) d.A += i * K * K;
            if (d.g) d.g += i * K;
            if (d.dx) d.dx += i * K;
            if (d.uv) d.uv += i * kN;
            if (d.bias) d.bias += i * kN;
            if (d.J) d.J += i * K * kN;
            dp = &d;
        }
        Debug dh{};
        if (g_how) {
            if (!dp) { dp = &dh; }
            dp->how = g_how + i;
        }
        const Scratch scr{scratch.data(), 1};
        TailState<KD, HB> T;
        const bool hard = Q ? fused_instance<KD, HB, true>(P, M, R, io, i, scr, T, dp, Q)
                            : fused_instance<KD, HB, false>(P, M, R, io, i, scr, T, dp);
        if (hard) {
            ++n_hard;
            double w[K];
            state_record<KD, HB>(R, T, rec.data());
            const int fl = host_eigen_solve<K>(&rec[RC::A], &rec[RC::G], rec[RC::ABAD] == 0.0, w);
            fixup_finish<KD, HB>(R, T.u_all_row, T.ctrl_row, rec.data(), w, 0, 1);
            if (T.status) *T.status = (uint8_t)(*T.status | fl);
        }
    }
    return n_hard;
}

extern "C" const char *fused_host_error(void) { return g_err.c_str(); }

// Returns the number of instances that took the eigen path, or -1 on error.
extern "C" int64_t fused_host_run(const irlosc_params *params, const irlosc_model *model, int64_t B,
                                  const irlosc_fused_io *io, double *dbg_A, double *dbg_g, double *dbg_uv,
                                  double *dbg_bias, double *dbg_dx, double *dbg_J) {
    KParams P;
    if (build_kparams(*params, P) != IRLOSC_OK) return -1;
    FRoles R;
    int kd = 0;
    bool hb = false;
    if (!fused_roles(P, R, kd, hb)) { irlosc::fail(1, "not the DualUR5 topology"); return -1; }
    KModel M;
    if (build_kmodel(P, *model, M) != IRLOSC_OK) return -1;
    FIo k;
    k.q = io->q; k.dq = io->dq; k.target_xyz = io->target_xyz; k.target_quat = io->target_quat;
    k.target_vel = io->target_vel; k.max_vel = io->max_vel; k.ft_raw = io->ft_raw;
    k.ctrl = io->ctrl; k.u_all = io->u_all; k.status = io->status; k.ee_xyz = io->ee_xyz; k.ee_quat = io->ee_quat;
    k.wp_xyz = k.wp_quat = nullptr;
    k.seq_action = k.seq_entered = k.seq_timer = nullptr;
    k.seq_err = k.seq_mv0 = k.seq_tgt_xyz = k.seq_tgt_quat = nullptr;
    k.wps = nullptr; k.wp_idx = nullptr;
    Debug d{dbg_A, dbg_g, dbg_uv, dbg_bias, dbg_dx, dbg_J, nullptr};
    const Debug *dp = (dbg_A || dbg_uv || dbg_J) ? &d : nullptr;
    if (kd == 3 && hb) return run<3, true>(P, M, R, k, B, dp);
    if (kd == 3 && !hb) return run<3, false>(P, M, R, k, B, dp);
    if (kd == 6 && hb) return run<6, true>(P, M, R, k, B, dp);
    return run<6, false>(P, M, R, k, B, dp);
}

// One control step of an action sequence (irlosc_step_sequence) on the CPU.
extern "C" int64_t sequence_host_step(const irlosc_params *params, const irlosc_model *model, int64_t B,
                                      const irlosc_fused_io *io, const irlosc_sequence *seq,
                                      const irlosc_sequence_io *sio) {
    KParams P;
    if (build_kparams(*params, P) != IRLOSC_OK) return -1;
    FRoles R;
    int kd = 0;
    bool hb = false;
    if (!fused_roles(P, R, kd, hb)) { irlosc::fail(1, "not the DualUR5 topology"); return -1; }
    KModel M;
    if (build_kmodel(P, *model, M) != IRLOSC_OK) return -1;
    KSeq Q;
    if (build_kseq(P, R, *seq, Q) != IRLOSC_OK) return -1;
    FIo k;
    k.q = io->q; k.dq = io->dq; k.target_xyz = sio->target_xyz; k.target_quat = sio->target_quat;
    k.target_vel = io->target_vel; k.max_vel = io->max_vel; k.ft_raw = io->ft_raw;
    k.ctrl = io->ctrl; k.u_all = io->u_all; k.status = io->status; k.ee_xyz = io->ee_xyz; k.ee_quat = io->ee_quat;
    k.wp_xyz = sio->wp_xyz; k.wp_quat = sio->wp_quat;
    k.seq_action = sio->action; k.seq_entered = sio->entered; k.seq_timer = sio->timer;
    k.seq_err = sio->err; k.seq_mv0 = sio->max_vel0; k.seq_tgt_xyz = sio->target_xyz; k.seq_tgt_quat = sio->target_quat;
    k.wps = nullptr; k.wp_idx = nullptr;
    if (kd == 3 && hb) return run<3, true>(P, M, R, k, B, nullptr, &Q);
    if (kd == 3 && !hb) return run<3, false>(P, M, R, k, B, nullptr, &Q);
    if (kd == 6 && hb) return run<6, true>(P, M, R, k, B, nullptr, &Q);
    return run<6, false>(P, M, R, k, B, nullptr, &Q);
}

// ------------------------------------------------------------------ streaming step (state given)
// Emulates the staging of osc_step_stream on the CPU: every group of the host-built copy plan is
// gathered into a stage exactly as the cp.async chunks would, then the per-instance consumers
// (the same __host__ __device__ code the kernel runs) read it.
namespace {
struct HostGroups {
    const stream::Plan &plan;
    int64_t inst;
    std::vector<double> stage;
    struct Reader {
        const double *st;
        __host__ __device__ double operator()(int e) const { return st[e]; }
    };
    __host__ __device__ Reader operator()(int g) {
#ifndef __CUDA_ARCH__
        std::fill(stage.begin(), stage.end(), std::nan(""));      // anything not copied must not be used
        for (int c = plan.first[g]; c < plan.first[g + 1]; ++c) {
            const stream::Chunk &ch = plan.ch[c];
            for (int l = 0; l < 8; ++l) {
                const unsigned char *src = reinterpret_cast<const unsigned char *>(ch.base) + inst * (int64_t)ch.stride + ch.off[l];
                stage[ch.dst[l]] = *reinterpret_cast<const double *>(src);
            }
        }
        return Reader{stage.data()};
#else
        return Reader{nullptr};
#endif
    }
};

template <int KD, bool HB>
int64_t run_stream(const KParams &P, const FRoles &R, const stream::Plan &plan, const stream::Outputs &out, int64_t B,
                   const Debug *dbg0) {
    using RC = Rec<KD, HB>;
    constexpr int K = RC::K;
    std::vector<double> rec(RC::SIZE);
    int64_t n_hard = 0;
    for (int64_t i = 0; i < B; ++i) {
        Debug d, *dp = nullptr;
        if (dbg0) {
            d = *dbg0;
            if (d.A) d.A += i * K * K;
            if (d.g) d.g += i * K;
            if (d.dx) d.dx += i * K;
            if (d.uv) d.uv += i * kN;
            if (d.bias) d.bias = nullptr;
            if (d.J) d.J += i * K * kN;
            dp = &d;
        }
        Debug dh{};
        if (g_how) {
            if (!dp) { dp = &dh; }
            dp->how = g_how + i;
        }
        HostGroups groups{plan, i, std::vector<double>(plan.stage_entries)};
        double *ctrl_row = out.ctrl + i * P.n_ctrl;
        TailState<KD, HB> T;
        const bool hard = stream::stream_instance<KD, HB>(P, R, plan, out, i, groups, ctrl_row, T, dp);
        if (hard) {
            ++n_hard;
            double w[K];
            state_record<KD, HB>(R, T, rec.data());
            const int fl = host_eigen_solve<K>(&rec[RC::A], &rec[RC::G], rec[RC::ABAD] == 0.0, w);
            fixup_finish<KD, HB>(R, T.u_all_row, T.ctrl_row, rec.data(), w, 0, 1);
            if (T.status) *T.status = (uint8_t)(*T.status | fl);
        }
    }
    return n_hard;
}
}  // namespace

extern "C" int64_t stream_host_run(const irlosc_params *params, int64_t B, const irlosc_io *io, double *dbg_A,
                                   double *dbg_g, double *dbg_uv, double *dbg_dx, double *dbg_J, int32_t *n_chunks) {
    KParams P;
    if (build_kparams(*params, P) != IRLOSC_OK) return -1;
    FRoles R;
    int kd = 0;
    bool hb = false;
    if (!fused_roles(P, R, kd, hb)) { irlosc::fail(1, "not the DualUR5 topology"); return -1; }
    KIo k;
    memset(&k, 0, sizeof k);
    if (resolve_m_layout(P, *io, k) != IRLOSC_OK) return -1;
    k.J = io->J; k.j_layout = io->j_layout; k.ldj = io->ldj ? io->ldj : P.n;
    k.j_stride = io->j_stride ? io->j_stride : (io->j_layout == IRLOSC_J_ROWS ? (int64_t)k.ldj * P.k : (int64_t)k.ldj * 6 * P.D);
    k.dq = io->dq; k.bias = io->bias; k.ee_xyz = io->ee_xyz; k.ee_quat = io->ee_quat;
    k.target_xyz = io->target_xyz; k.target_quat = io->target_quat; k.target_vel = io->target_vel;
    k.max_vel = io->max_vel; k.ft_xmat = io->ft_xmat; k.ft_raw = io->ft_raw;
    stream::Plan plan;
    if (build_stream_plan(P, k, R, kd, hb, plan) != IRLOSC_OK) return -1;
    if (n_chunks) *n_chunks = plan.n_chunks;
    stream::Outputs out{io->u_all, io->ctrl, io->status, io->target_vel};
    Debug d{dbg_A, dbg_g, dbg_uv, nullptr, dbg_dx, dbg_J, nullptr};
    const Debug *dp = (dbg_A || dbg_uv || dbg_J) ? &d : nullptr;
    if (kd == 3 && hb) return run_stream<3, true>(P, R, plan, out, B, dp);
    if (kd == 3 && !hb) return run_stream<3, false>(P, R, plan, out, B, dp);
    if (kd == 6 && hb) return run_stream<6, true>(P, R, plan, out, B, dp);
    return run_stream<6, false>(P, R, plan, out, B, dp);
}

// One control step of gain_test-style waypoint cycling (irlosc_step_waypoints) on the CPU.
extern "C" int64_t waypoints_host_step(const irlosc_params *params, const irlosc_model *model, int64_t B,
                                       const irlosc_fused_io *io, const irlosc_waypoints_io *wio) {
    KParams P;
    if (build_kparams(*params, P) != IRLOSC_OK) return -1;
    FRoles R;
    int kd = 0;
    bool hb = false;
    if (!fused_roles(P, R, kd, hb)) { irlosc::fail(1, "not the DualUR5 topology"); return -1; }
    KModel M;
    if (build_kmodel(P, *model, M) != IRLOSC_OK) return -1;
    KSeq Q;
    if (build_kseq_waypoints(P, R, *wio, Q) != IRLOSC_OK) return -1;
    FIo k;
    memset(&k, 0, sizeof k);
    k.q = io->q; k.dq = io->dq; k.target_xyz = wio->target_xyz; k.target_quat = wio->target_quat;
    k.target_vel = io->target_vel; k.max_vel = io->max_vel; k.ft_raw = io->ft_raw;
    k.ctrl = io->ctrl; k.u_all = io->u_all; k.status = io->status; k.ee_xyz = io->ee_xyz; k.ee_quat = io->ee_quat;
    k.wps = wio->wps; k.wp_idx = wio->wp_idx; k.seq_tgt_xyz = wio->target_xyz; k.seq_tgt_quat = wio->target_quat;
    if (kd == 3 && hb) return run<3, true>(P, M, R, k, B, nullptr, &Q);
    if (kd == 3 && !hb) return run<3, false>(P, M, R, k, B, nullptr, &Q);
    if (kd == 6 && hb) return run<6, true>(P, M, R, k, B, nullptr, &Q);
    return run<6, false>(P, M, R, k, B, nullptr, &Q);
}


// ------------------------------------------------------------------ lane step (state in batch-interleaved tiles)
// Packs the caller's arrays into tiles with the product's own pack table (osc_lane.cuh: build_tile_spec,
// build_pack_table, pack_fetch - the functions irlosc_pack_tiles[_host] run) and runs the lane kernel's
// per-instance function on them.
#include "../../irl_control_b200/csrc/osc_lane.cuh"
namespace {
struct LdHost {
    double operator()(const double *p) const { return *p; }
};
struct TileGroups {
    const double *tl;          // entry 0 of this lane
    const int32_t *gbase;
    struct Reader {
        const double *p;
        double operator()(int e) const { return p[(size_t)e * lane::kTile]; }
    };
    Reader operator()(int g) const { return Reader{tl + (size_t)gbase[g] * lane::kTile}; }
};

template <int KD, bool HB>
int64_t run_lane(const KParams &P, const FRoles &R, const lane::TileSpec &S, const double *tiles, const irlosc_io *io, int64_t B) {
    using RC = Rec<KD, HB>;
    constexpr int K = RC::K;
    std::vector<double> rec(RC::SIZE);
    int64_t n_hard = 0;
    for (int64_t i = 0; i < B; ++i) {
        const double *tl = tiles + (i / lane::kTile) * (int64_t)S.n_entries * lane::kTile + (i % lane::kTile);
        TileGroups groups{tl, S.gbase};
        const lane::JTile<KD, HB, LdHost> ja{tl, {S.gbase[4], S.gbase[9]}, S.gbase[0], LdHost{}};
        lane::LaneState<KD, HB> T;
        T.u_all_row = io->u_all ? io->u_all + i * kN : nullptr;
        T.ctrl_row = io->ctrl + i * P.n_ctrl;
        T.status = io->status ? io->status + i : nullptr;
        Debug dh{};
        if (g_how) dh.how = g_how + i;
        const bool hard = lane::lane_instance<KD, HB>(P, R, io->target_vel ? io->target_vel + i * P.D * 6 : nullptr, groups, ja, T,
                                                      g_how ? &dh : nullptr);
        if (hard) {
            ++n_hard;
            double w[K];
            tail_record<KD, HB>(R, T.akA, T.j0, T.g, ja, T.base_arm, T.base_st, T.inv0, T.force_pinv, rec.data());
            const int fl = host_eigen_solve<K>(&rec[RC::A], &rec[RC::G], rec[RC::ABAD] == 0.0, w);
            fixup_finish<KD, HB>(R, T.u_all_row, T.ctrl_row, rec.data(), w, 0, 1);
            if (T.status) *T.status = (uint8_t)(*T.status | fl);
        }
    }
    return n_hard;
}
}  // namespace

// Entry table of the tile layout; returns E (0: no tile layout for this controller).
extern "C" int32_t lane_host_spec(const irlosc_params *params, irlosc_tile_entry *out, int32_t capacity, int32_t *gbase_out) {
    KParams P;
    if (build_kparams(*params, P) != IRLOSC_OK) return -1;
    FRoles R;
    int kd = 0;
    bool hb = false;
    if (!fused_roles(P, R, kd, hb)) return 0;
    static lane::TileSpec S;
    if (!lane::build_tile_spec(P, R, kd, hb, S)) return 0;
    for (int e = 0; out && e < S.n_entries && e < capacity; ++e) out[e] = S.e[e];
    for (int g = 0; gbase_out && g <= lane::kGroups; ++g) gbase_out[g] = S.gbase[g];
    return S.n_entries;
}

// arrays -> tiles (tiles_out: [ceil(B / 32)][E][32]) -> lane step.  Returns the number of instances finished by the
// eigen-solver, or -1.
extern "C" int64_t lane_host_run(const irlosc_params *params, int64_t B, const irlosc_io *io, double *tiles_out) {
    KParams P;
    if (build_kparams(*params, P) != IRLOSC_OK) return -1;
    FRoles R;
    int kd = 0;
    bool hb = false;
    if (!fused_roles(P, R, kd, hb)) { irlosc::fail(1, "not the DualUR5 topology"); return -1; }
    static lane::TileSpec S;
    if (!lane::build_tile_spec(P, R, kd, hb, S)) { irlosc::fail(1, "no tile layout"); return -1; }
    KIo k;
    memset(&k, 0, sizeof k);
    if (resolve_m_layout(P, *io, k) != IRLOSC_OK) return -1;
    k.J = io->J; k.j_layout = io->j_layout; k.ldj = io->ldj ? io->ldj : P.n;
    k.j_stride = io->j_stride ? io->j_stride : (io->j_layout == IRLOSC_J_ROWS ? (int64_t)k.ldj * P.k : (int64_t)k.ldj * 6 * P.D);
    k.dq = io->dq; k.bias = io->bias; k.ee_xyz = io->ee_xyz; k.ee_quat = io->ee_quat;
    k.target_xyz = io->target_xyz; k.target_quat = io->target_quat; k.target_vel = io->target_vel;
    k.max_vel = io->max_vel; k.ft_xmat = io->ft_xmat; k.ft_raw = io->ft_raw;
    static lane::PackTable T;
    if (lane::build_pack_table(P, k, S, T) != IRLOSC_OK) return -1;
    const int64_t n_tiles = (B + lane::kTile - 1) / lane::kTile;
    for (int64_t t = 0; t < n_tiles; ++t)
        for (int l = 0; l < lane::kTile; ++l) {
            const int64_t inst = std::min<int64_t>(t * lane::kTile + l, B - 1);
            for (int e = 0; e < S.n_entries; ++e) tiles_out[(t * S.n_entries + e) * lane::kTile + l] = lane::pack_fetch(T, e, inst);
        }
    if (kd == 3 && hb) return run_lane<3, true>(P, R, S, tiles_out, io, B);
    if (kd == 3 && !hb) return run_lane<3, false>(P, R, S, tiles_out, io, B);
    if (kd == 6 && hb) return run_lane<6, true>(P, R, S, tiles_out, io, B);
    return run_lane<6, false>(P, R, S, tiles_out, io, B);
}

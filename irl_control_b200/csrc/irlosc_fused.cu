// C ABI of the fused state provider + OSC step (include/irlosc.h: irlosc_set_model,
// irlosc_step_fused, irlosc_step_fused_host).  Kernels: osc_fused.cuh.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <cuda_runtime.h>

#include "irlosc_internal.h"
#include "irlosc_build.h"
#include "osc_fused.cuh"
#include "osc_fused_pair.cuh"
#include "osc_stream.cuh"
#include <cstdlib>

using namespace irlosc;
using namespace irlosc::fused;

namespace {

using fused_build::kDualUr5Parent;

struct FusedEntry {
    int kd;
    bool has_base;
    int variant;              // 0 = default; others for A/B measurements (irlosc_set_kernel(h, 2 + variant))
    int threads;
    size_t smem;
    const void *step;
    const char *name;
    const void *step_seq;     // same kernel with the action-sequence state machine compiled in (variant 0 only)
};

template <int KD, bool HB, int NT, bool SMEM>
FusedEntry entry(int variant, const char *name) {
    const void *seq = nullptr;
    if constexpr (SMEM && (NT == 256 || NT == 224)) seq = (const void *)osc_step_fused<KD, HB, NT, SMEM, true>;
    // chain scratch (SMEM variants) + one warp-finish scratch per warp
    const size_t smem = (SMEM ? (size_t)kScratchDoubles * NT * sizeof(double) : 0) + (size_t)(NT / 32) * sizeof(WarpFix<KD, HB>);
    return FusedEntry{KD, HB, variant, NT, smem, (const void *)osc_step_fused<KD, HB, NT, SMEM>, name, seq};
}

// Variant 0 is the default and exists with 8 and with 7 warps per CTA (one CTA per SM): a thread owns an
// instance, so a batch is ceil(B / 32 / (SMs * warps)) passes, and with 148 SMs the headline batch of
// 65 536 is 1.73 passes at 8 warps but 1.98 at 7.  launch_fused picks the cheaper of the two for B
// (measured pass times: 8 warps 92.5 us, 7 warps 82.5 us; profiles/r01_fused_variants_final.jsonl).
const FusedEntry *fused_table(int *count) {
    static const FusedEntry t[] = {
        entry<3, true, 256, true>(0, "osc_step_fused<kd3,base,t256,smem>"),
        entry<6, false, 256, true>(0, "osc_step_fused<kd6,t256,smem>"),
        entry<6, true, 256, true>(0, "osc_step_fused<kd6,base,t256,smem>"),
        entry<3, false, 256, true>(0, "osc_step_fused<kd3,t256,smem>"),
        entry<3, true, 224, true>(0, "osc_step_fused<kd3,base,t224,smem>"),
        entry<6, false, 224, true>(0, "osc_step_fused<kd6,t224,smem>"),
        entry<6, true, 224, true>(0, "osc_step_fused<kd6,base,t224,smem>"),
        entry<3, false, 224, true>(0, "osc_step_fused<kd3,t224,smem>"),
    };
    *count = (int)(sizeof t / sizeof t[0]);
    return t;
}

const FusedEntry *fused_find(int kd, bool has_base, int variant = 0, int threads = 0) {
    int cnt = 0;
    const FusedEntry *t = fused_table(&cnt);
    for (int i = 0; i < cnt; ++i)
        if (t[i].kd == kd && t[i].has_base == has_base && t[i].variant == variant && (threads == 0 || t[i].threads == threads))
            return &t[i];
    return nullptr;
}

// Two lanes per instance (osc_fused_pair.cuh), 8 warps x 255 registers like the tile pair kernel.
struct FusedPairEntry {
    int kd;
    bool has_base;
    int threads;
    size_t smem;
    const void *step;
    const char *name;
};
template <int KD, bool HB, int NT>
FusedPairEntry fpentry(const char *name) {
    return FusedPairEntry{KD, HB, NT, (size_t)kScratchDoubles * NT * sizeof(double) + (size_t)(NT / 32) * sizeof(WarpFix<KD, HB>),
                          (const void *)osc_step_fused_pair<KD, HB, NT>, name};
}
const FusedPairEntry *fused_pair_table(int *count) {
    static const FusedPairEntry t[] = {
        fpentry<3, true, 256>("osc_step_fused_pair<kd3,base,t256>"), fpentry<3, false, 256>("osc_step_fused_pair<kd3,t256>"),
        fpentry<6, true, 256>("osc_step_fused_pair<kd6,base,t256>"), fpentry<6, false, 256>("osc_step_fused_pair<kd6,t256>"),
    };
    *count = (int)(sizeof t / sizeof t[0]);
    return t;
}

// 8 or 7 warps per CTA for a batch of B (see fused_table)
int fused_threads_for(int64_t B, int sms) {
    const int64_t tiles = (B + 31) / 32;
    const double p8 = (double)((tiles + (int64_t)sms * 8 - 1) / ((int64_t)sms * 8)) * 92.5;
    const double p7 = (double)((tiles + (int64_t)sms * 7 - 1) / ((int64_t)sms * 7)) * 82.5;
    return p7 < p8 ? 224 : 256;
}

int32_t check_fio(const irlosc_handle *h, const irlosc_fused_io *io, FIo &k) {
    if (!io) return fail(IRLOSC_ERR_INVALID, "io is null");
    if (!io->q || !io->dq || !io->target_xyz || !io->target_quat || !io->ctrl)
        return fail(IRLOSC_ERR_INVALID, "a required array (q, dq, target_xyz, target_quat, ctrl) is null");
    if (h->kp.admittance && !io->ft_raw) return fail(IRLOSC_ERR_INVALID, "admittance is set but ft_raw is null");
    k.q = io->q; k.dq = io->dq; k.target_xyz = io->target_xyz; k.target_quat = io->target_quat;
    k.target_vel = io->target_vel; k.max_vel = io->max_vel; k.ft_raw = io->ft_raw;
    k.ctrl = io->ctrl; k.u_all = io->u_all; k.status = io->status; k.ee_xyz = io->ee_xyz; k.ee_quat = io->ee_quat;
    k.wp_xyz = k.wp_quat = nullptr;
    k.seq_action = k.seq_entered = k.seq_timer = nullptr;
    k.seq_err = k.seq_mv0 = k.seq_tgt_xyz = k.seq_tgt_quat = nullptr;
    k.wps = nullptr;
    k.wp_idx = nullptr;
    return IRLOSC_OK;
}

// B_whole: the batch the caller handed in (the host entry point launches chunks of it); the kernel choice follows it.
int32_t launch_fused(irlosc_handle *h, int64_t B, const FIo &k, cudaStream_t st, const KSeq *seq = nullptr, int64_t B_whole = -1) {
    const int variant = (h->kernel_choice >= 2 && h->kernel_choice < 9) ? h->kernel_choice - 2 : 0;
    const int sms = std::max(1, h->sm_count - h->sm_margin);
    int want_threads = variant == 0 ? fused_threads_for(B, sms) : 0;
    if (const char *t = getenv("IRLOSC_FUSED_THREADS")) want_threads = atoi(t);                         // experiments only
    if (!seq && variant == 0) {
        // one wave of half tiles or less: the pair kernel's latency is one arm instead of two (measured -23 % at k = 7,
        // -33 % at k = 12 for B <= 16 384); beyond that the thread-per-instance kernel is as fast or faster
        int use_pair = ((B_whole < 0 ? B : B_whole) + 15) / 16 <= (int64_t)sms * 8 ? 1 : 0;
        if (const char *t = getenv("IRLOSC_FUSED_PAIR")) use_pair = atoi(t);                             // experiments only
        if (use_pair) {
            int pc = 0;
            const FusedPairEntry *pt = fused_pair_table(&pc), *pe = nullptr;
            for (int i = 0; i < pc; ++i)
                if (pt[i].kd == h->fused_kd && pt[i].has_base == h->fused_base) pe = &pt[i];
            if (!pe) return fail(IRLOSC_ERR_INVALID, "no fused pair kernel for kd=%d base=%d", h->fused_kd, (int)h->fused_base);
            static bool ready = false;
            if (!ready) {
                for (int i = 0; i < pc; ++i)
                    CUDA_TRY(cudaFuncSetAttribute(pt[i].step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pt[i].smem));
                ready = true;
            }
            const int64_t n_half = (B + 15) / 16;
            // spread a small batch over the SMs: CTAs of 1 .. 8 warps (see launch_lane)
            const int pw = (int)std::max<int64_t>(1, std::min<int64_t>(pe->threads / 32, (n_half + sms - 1) / sms));
            const int pgrid = (int)std::min<int64_t>((n_half + pw - 1) / pw, (int64_t)sms);
            const size_t psmem = pe->smem / (pe->threads / 32) * pw;
            void *pargs[] = {(void *)&h->kp, (void *)&h->km, (void *)&k, (void *)&B, (void *)&h->fr};
            cudaError_t perr = cudaLaunchKernel(pe->step, dim3(pgrid), dim3(pw * 32), pargs, psmem, st);
            if (perr != cudaSuccess) return fail(IRLOSC_ERR_CUDA, "fused pair kernel launch: %s", cudaGetErrorString(perr));
            h->launches += 1;
            h->last_kernel = pe->name;
            return IRLOSC_OK;
        }
    }
    const FusedEntry *e = fused_find(h->fused_kd, h->fused_base, variant, want_threads);
    if (!e) e = fused_find(h->fused_kd, h->fused_base, variant);
    if (!e) return fail(IRLOSC_ERR_INVALID, "no fused kernel for kd=%d base=%d variant=%d", h->fused_kd, (int)h->fused_base, variant);
    const int grid = (int)std::min<int64_t>((B + e->threads - 1) / e->threads, (int64_t)sms);
    static const KSeq no_seq = {};
    const void *fn = e->step;
    if (seq) {
        if (!e->step_seq) return fail(IRLOSC_ERR_INVALID, "the action-sequence step exists for fused variant 0 only");
        fn = e->step_seq;
    }
    void *args[] = {(void *)&h->kp, (void *)&h->km, (void *)&k, (void *)&B, (void *)&h->fr, (void *)(seq ? seq : &no_seq)};
    cudaError_t err = cudaLaunchKernel(fn, dim3(grid), dim3(e->threads), args, e->smem, st);
    if (err != cudaSuccess) return fail(IRLOSC_ERR_CUDA, "fused kernel launch: %s", cudaGetErrorString(err));
    h->launches += 1;
    h->last_kernel = e->name;
    return IRLOSC_OK;
}

// ------------------------------------------------------------------ streaming step (state in HBM)
struct StreamEntry {
    int kd;
    bool has_base;
    int max_threads;
    const void *step;
    int fix_bytes;            // per-warp scratch of the warp finish
    const char *name;
};

template <int KD, bool HB, int NT>
StreamEntry sentry(const char *name) {
    return StreamEntry{KD, HB, NT, (const void *)stream::osc_step_stream<KD, HB, NT>, (int)((sizeof(WarpFix<KD, HB>) + 15) & ~size_t(15)), name};
}

const StreamEntry *stream_table(int *count) {
    static const StreamEntry t[] = {
        sentry<3, true, 256>("osc_step_stream<kd3,base>"),
        sentry<6, false, 256>("osc_step_stream<kd6>"),
        sentry<6, true, 256>("osc_step_stream<kd6,base>"),
        sentry<3, false, 256>("osc_step_stream<kd3>"),
    };
    *count = (int)(sizeof t / sizeof t[0]);
    return t;
}

constexpr size_t kSmemLimit = 227 * 1024;
bool g_stream_ready = false;

}  // namespace

bool irlosc::stream_supported(const irlosc_handle *h, const KIo &) {
    FRoles R;
    int kd = 0;
    bool hb = false;
    return !h->kp.check_topology && fused_roles(h->kp, R, kd, hb);
}

bool irlosc::stream_preferred(const irlosc_handle *h) {
    FRoles R;
    int kd = 0;
    bool hb = false;
    return fused_roles(h->kp, R, kd, hb) && kd == 6;
}

int32_t irlosc::stream_launch(irlosc_handle *h, int64_t B, const KIo &io, cudaStream_t st) {
    FRoles R;
    int kd = 0;
    bool has_base = false;
    if (!fused_roles(h->kp, R, kd, has_base)) return fail(IRLOSC_ERR_INVALID, "streaming kernel: not the DualUR5 topology");
    int cnt = 0;
    const StreamEntry *t = stream_table(&cnt), *e = nullptr;
    for (int i = 0; i < cnt; ++i)
        if (t[i].kd == kd && t[i].has_base == has_base) e = &t[i];
    if (!e) return fail(IRLOSC_ERR_INVALID, "no streaming kernel for kd=%d base=%d", kd, (int)has_base);
    if (!g_stream_ready) {
        for (int i = 0; i < cnt; ++i)
            CUDA_TRY(cudaFuncSetAttribute(t[i].step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
        g_stream_ready = true;
    }
    stream::Plan plan;
    int32_t rc = build_stream_plan(h->kp, io, R, kd, has_base, plan);
    if (rc != IRLOSC_OK) return rc;
    // shared-memory plan: tables, then per warp two stages, the packed ctrl tile and the warp-finish scratch
    const int stage_bytes = (plan.stage_entries * stream::kPitch * 8 + 15) & ~15;
    const int warp_bytes = 2 * stage_bytes + ((32 * h->kp.n_ctrl * 8 + 15) & ~15) + e->fix_bytes;
    const size_t head = (sizeof(stream::Plan) + 15) & ~size_t(15);
    int warps = (int)std::min<size_t>(e->max_threads / 32, (kSmemLimit - head) / warp_bytes);
    {   // a warp owns 32 instances: pick the warp count whose number of passes over the batch is cheapest
        // (pass times measured at 4 / 6 / 7 / 8 warps, profiles/r01_stream_vs_tree.log)
        static const double pass_us[9] = {0, 40, 42, 44, 45.5, 47, 55, 66.5, 79};
        const int64_t tiles_ = (B + 31) / 32;
        const int sms_ = std::max(1, h->sm_count - h->sm_margin);
        int best = warps;
        double best_t = 1e300;
        for (int w = std::min(warps, 8); w >= 4; --w) {
            const double t_ = (double)((tiles_ + (int64_t)sms_ * w - 1) / ((int64_t)sms_ * w)) * pass_us[w];
            if (t_ < best_t - 1e-9) { best_t = t_; best = w; }
        }
        warps = best;
    }
    if (const char *w = getenv("IRLOSC_STREAM_WARPS")) warps = std::max(1, std::min(warps, atoi(w)));   // experiments only
    if (warps < 1) return fail(IRLOSC_ERR_INVALID, "streaming kernel: a stage does not fit in shared memory");
    const size_t smem = head + (size_t)warps * warp_bytes;
    stream::Outputs out{io.u_all, io.ctrl, io.status, io.target_vel};
    auto al16 = [](const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    stream::Gather G;
    memset(&G, 0, sizeof G);
    G.n_gather = io.n_gather;
    G.gather_offset = io.gather_offset;
    G.ctrl_mc = io.ctrl_mc;
    G.ctrl_vec = al16(io.ctrl) && ((io.gather_offset * (int64_t)h->kp.n_ctrl) % 2 == 0) && al16(io.ctrl_mc);
    for (int g = 0; g < io.n_gather; ++g) { G.ctrl_gather[g] = io.ctrl_gather[g]; G.ctrl_vec = G.ctrl_vec && al16(io.ctrl_gather[g]); }
    const int sms = std::max(1, h->sm_count - h->sm_margin);
    const int64_t n_tiles = (B + 31) / 32;
    const int grid = (int)std::min<int64_t>((n_tiles + warps - 1) / warps, (int64_t)sms);
    int mode = 0;
    if (const char *m = getenv("IRLOSC_STREAM_MODE")) mode = atoi(m);                                   // experiments only
    void *args[] = {(void *)&h->kp, (void *)&plan, (void *)&out, (void *)&B, (void *)&R, (void *)&G,
                    (void *)&stage_bytes, (void *)&warp_bytes, (void *)&mode};
    cudaError_t err = cudaLaunchKernel(e->step, dim3(grid), dim3(warps * 32), args, smem, st);
    if (err != cudaSuccess) return fail(IRLOSC_ERR_CUDA, "streaming kernel launch: %s", cudaGetErrorString(err));
    h->launches += 1;
    h->last_kernel = e->name;
    return IRLOSC_OK;
}

extern "C" int32_t irlosc_set_model(irlosc_handle *h, const irlosc_model *m) {
    if (!h || !m) return fail(IRLOSC_ERR_INVALID, "null argument");
    h->has_model = false;
    if (m->n_joints != h->kp.n) return fail(IRLOSC_ERR_INVALID, "model has %d joints, controller n=%d", m->n_joints, h->kp.n);
    int kd = 0;
    bool has_base = false;
    if (!fused_roles(h->kp, h->fr, kd, has_base))
        return fail(IRLOSC_ERR_INVALID, "the fused step needs the DualUR5 topology (has_topology, two arm devices with 3 or 6 "
                                        "rows each, optionally the base with one row)");
    for (int j = 0; j < kN; ++j)
        if (m->joint[j].parent != kDualUr5Parent[j])
            return fail(IRLOSC_ERR_INVALID, "model joint %d has parent %d, the DualUR5 tree has %d", j, m->joint[j].parent,
                        kDualUr5Parent[j]);
    if (!fused_find(kd, has_base)) return fail(IRLOSC_ERR_INVALID, "no fused kernel for kd=%d base=%d", kd, (int)has_base);
    int32_t rc = build_kmodel(h->kp, *m, h->km);
    if (rc != IRLOSC_OK) return rc;
    int cnt = 0;
    const FusedEntry *t = fused_table(&cnt);
    for (int i = 0; i < cnt; ++i) {
        CUDA_TRY(cudaFuncSetAttribute(t[i].step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t[i].smem));
        if (t[i].step_seq)
            CUDA_TRY(cudaFuncSetAttribute(t[i].step_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t[i].smem));
        if (t[i].smem == 0)    // scratch in local memory: give the whole SM to L1
            CUDA_TRY(cudaFuncSetAttribute(t[i].step, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
    }
    h->fused_kd = kd;
    h->fused_base = has_base;
    h->has_model = true;
    return IRLOSC_OK;
}

extern "C" int32_t irlosc_step_fused(irlosc_handle *h, int64_t B, const irlosc_fused_io *io, void *cuda_stream) {
    if (!h) return fail(IRLOSC_ERR_INVALID, "handle is null");
    if (!h->has_model) return fail(IRLOSC_ERR_INVALID, "irlosc_set_model has not been called");
    if (B < 0) return fail(IRLOSC_ERR_INVALID, "B=%lld is negative", (long long)B);
    if (B == 0) return IRLOSC_OK;
    FIo k;
    int32_t rc = check_fio(h, io, k);
    if (rc != IRLOSC_OK) return rc;
    return launch_fused(h, B, k, (cudaStream_t)cuda_stream);
}

extern "C" int32_t irlosc_step_sequence(irlosc_handle *h, int64_t B, const irlosc_fused_io *io, const irlosc_sequence *seq,
                                        const irlosc_sequence_io *sio, void *cuda_stream) {
    if (!h) return fail(IRLOSC_ERR_INVALID, "handle is null");
    if (!h->has_model) return fail(IRLOSC_ERR_INVALID, "irlosc_set_model has not been called");
    if (!io || !seq || !sio) return fail(IRLOSC_ERR_INVALID, "null argument");
    if (B < 0) return fail(IRLOSC_ERR_INVALID, "B=%lld is negative", (long long)B);
    if (B == 0) return IRLOSC_OK;
    if (!sio->wp_xyz || !sio->wp_quat || !sio->action || !sio->entered || !sio->timer || !sio->err || !sio->max_vel0 ||
        !sio->target_xyz || !sio->target_quat)
        return fail(IRLOSC_ERR_INVALID, "every array of irlosc_sequence_io is required");
    KSeq Q;
    int32_t rc = build_kseq(h->kp, h->fr, *seq, Q);
    if (rc != IRLOSC_OK) return rc;
    irlosc_fused_io io2 = *io;
    io2.target_xyz = sio->target_xyz;            // the targets live in the episode state
    io2.target_quat = sio->target_quat;
    FIo k;
    rc = check_fio(h, &io2, k);
    if (rc != IRLOSC_OK) return rc;
    k.wp_xyz = sio->wp_xyz; k.wp_quat = sio->wp_quat;
    k.seq_action = sio->action; k.seq_entered = sio->entered; k.seq_timer = sio->timer;
    k.seq_err = sio->err; k.seq_mv0 = sio->max_vel0; k.seq_tgt_xyz = sio->target_xyz; k.seq_tgt_quat = sio->target_quat;
    return launch_fused(h, B, k, (cudaStream_t)cuda_stream, &Q);
}

extern "C" int32_t irlosc_step_waypoints(irlosc_handle *h, int64_t B, const irlosc_fused_io *io,
                                         const irlosc_waypoints_io *wio, void *cuda_stream) {
    if (!h) return fail(IRLOSC_ERR_INVALID, "handle is null");
    if (!h->has_model) return fail(IRLOSC_ERR_INVALID, "irlosc_set_model has not been called");
    if (!io || !wio) return fail(IRLOSC_ERR_INVALID, "null argument");
    if (B < 0) return fail(IRLOSC_ERR_INVALID, "B=%lld is negative", (long long)B);
    if (B == 0) return IRLOSC_OK;
    KSeq Q;
    int32_t rc = build_kseq_waypoints(h->kp, h->fr, *wio, Q);
    if (rc != IRLOSC_OK) return rc;
    irlosc_fused_io io2 = *io;
    io2.target_xyz = wio->target_xyz;
    io2.target_quat = wio->target_quat;
    FIo k;
    rc = check_fio(h, &io2, k);
    if (rc != IRLOSC_OK) return rc;
    k.wps = wio->wps; k.wp_idx = wio->wp_idx; k.seq_tgt_xyz = wio->target_xyz; k.seq_tgt_quat = wio->target_quat;
    return launch_fused(h, B, k, (cudaStream_t)cuda_stream, &Q);
}

extern "C" int32_t irlosc_step_fused_host(irlosc_handle *h, int64_t B, const irlosc_fused_io *io) {
    if (!h) return fail(IRLOSC_ERR_INVALID, "handle is null");
    if (!h->has_model) return fail(IRLOSC_ERR_INVALID, "irlosc_set_model has not been called");
    if (B < 0) return fail(IRLOSC_ERR_INVALID, "B is negative");
    if (B == 0) return IRLOSC_OK;
    FIo hk;
    int32_t rc = check_fio(h, io, hk);
    if (rc != IRLOSC_OK) return rc;
    const KParams &P = h->kp;
    CUDA_TRY(cudaSetDevice(h->device));
    for (int s = 0; s < kPipeDepth; ++s)
        if (!h->fstage[s].stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->fstage[s].stream, cudaStreamNonBlocking));
    const size_t D = P.D, n = P.n;
    struct In { const double *src; size_t per; } ins[7] = {
        {hk.q, n}, {hk.dq, n}, {hk.target_xyz, 3 * D}, {hk.target_quat, 4 * D},
        {hk.target_vel, 6 * D}, {hk.max_vel, 2 * D}, {hk.ft_raw, 6 * D}};
    // Chunks rotate over kPipeDepth stages, each with its own stream, staging buffers and queue of
    // eigen-path instances, so the H2D copy of chunk i + 1 overlaps kernel and D2H copy of chunk i.
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(4 * h->host_chunk, B));   // 2 x measured slower (per-chunk call overhead)
    int turn = 0;
    int32_t result = IRLOSC_OK;
    for (int64_t b0 = 0; b0 < B && result == IRLOSC_OK; b0 += chunk, ++turn) {
        const int64_t nb = std::min<int64_t>(chunk, B - b0);
        Staging &S = h->fstage[turn % kPipeDepth];
        const double *dptr[7];
        for (int i = 0; i < 7 && result == IRLOSC_OK; ++i) {
            dptr[i] = nullptr;
            if (!ins[i].src) continue;
            result = ensure_cap(S, i, ins[i].per * (size_t)chunk * sizeof(double));
            if (result != IRLOSC_OK) break;
            CUDA_TRY(cudaMemcpyAsync(S.buf[i], ins[i].src + ins[i].per * (size_t)b0, ins[i].per * (size_t)nb * sizeof(double),
                                     cudaMemcpyHostToDevice, S.stream));
            dptr[i] = (const double *)S.buf[i];
        }
        if (result == IRLOSC_OK) result = ensure_cap(S, 8, (size_t)chunk * P.n_ctrl * sizeof(double));
        if (result == IRLOSC_OK && hk.u_all) result = ensure_cap(S, 9, (size_t)chunk * n * sizeof(double));
        if (result == IRLOSC_OK && hk.status) result = ensure_cap(S, 10, (size_t)chunk);
        if (result == IRLOSC_OK && hk.ee_xyz) result = ensure_cap(S, 11, (size_t)chunk * 3 * D * sizeof(double));
        if (result == IRLOSC_OK && hk.ee_quat) result = ensure_cap(S, 12, (size_t)chunk * 4 * D * sizeof(double));
        if (result != IRLOSC_OK) break;
        FIo dk;
        dk.q = dptr[0]; dk.dq = dptr[1]; dk.target_xyz = dptr[2]; dk.target_quat = dptr[3];
        dk.target_vel = dptr[4]; dk.max_vel = dptr[5]; dk.ft_raw = dptr[6];
        dk.ctrl = (double *)S.buf[8];
        dk.u_all = hk.u_all ? (double *)S.buf[9] : nullptr;
        dk.status = hk.status ? (uint8_t *)S.buf[10] : nullptr;
        dk.ee_xyz = hk.ee_xyz ? (double *)S.buf[11] : nullptr;
        dk.ee_quat = hk.ee_quat ? (double *)S.buf[12] : nullptr;
        result = launch_fused(h, nb, dk, S.stream, nullptr, B);
        if (result != IRLOSC_OK) break;
        CUDA_TRY(cudaMemcpyAsync(hk.ctrl + (size_t)b0 * P.n_ctrl, dk.ctrl, (size_t)nb * P.n_ctrl * sizeof(double),
                                 cudaMemcpyDeviceToHost, S.stream));
        if (hk.u_all)
            CUDA_TRY(cudaMemcpyAsync(hk.u_all + (size_t)b0 * n, dk.u_all, (size_t)nb * n * sizeof(double),
                                     cudaMemcpyDeviceToHost, S.stream));
        if (hk.status) CUDA_TRY(cudaMemcpyAsync(hk.status + b0, dk.status, (size_t)nb, cudaMemcpyDeviceToHost, S.stream));
        if (hk.ee_xyz)
            CUDA_TRY(cudaMemcpyAsync(hk.ee_xyz + (size_t)b0 * 3 * D, dk.ee_xyz, (size_t)nb * 3 * D * sizeof(double),
                                     cudaMemcpyDeviceToHost, S.stream));
        if (hk.ee_quat)
            CUDA_TRY(cudaMemcpyAsync(hk.ee_quat + (size_t)b0 * 4 * D, dk.ee_quat, (size_t)nb * 4 * D * sizeof(double),
                                     cudaMemcpyDeviceToHost, S.stream));
    }
    for (int s = 0; s < kPipeDepth; ++s) {
        cudaError_t e = cudaStreamSynchronize(h->fstage[s].stream);
        if (e != cudaSuccess && result == IRLOSC_OK) result = fail(IRLOSC_ERR_CUDA, "stream sync: %s", cudaGetErrorString(e));
    }
    return result;
}

// C ABI of the tiled step (include/irlosc.h: irlosc_tile_*, irlosc_pack_tiles[_host], irlosc_step_tiles[_host]).
// Kernels: osc_lane.cuh.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <cuda_runtime.h>

#include "irlosc_internal.h"
#include "irlosc_build.h"
#include "osc_lane.cuh"
#include "osc_pair.cuh"

using namespace irlosc;
using namespace irlosc::lane;

namespace {

struct LaneCtx {
    bool ok = false;              // the configuration has a tile layout
    int kd = 0;
    bool has_base = false;
    fused::FRoles R;
    TileSpec spec;
    Staging stage[kPipeDepth];    // host pipeline: 0 tiles, 1 target_vel, 2 ctrl, 3 u_all, 4 status
    // ticket counters of the pair kernel, one per stream that steps of this handle are issued on (a launch rewinds its
    // counter itself; launches on one stream run in order, launches on different streams must not share one)
    static constexpr int kSchedSlots = 16;
    int *sched = nullptr;
    cudaStream_t sched_stream[kSchedSlots] = {};
    int sched_used = 0;
};

LaneCtx *ctx_of(irlosc_handle *h) {
    if (!h->lane_ctx) {
        LaneCtx *c = new (std::nothrow) LaneCtx();
        if (!c) return nullptr;
        c->ok = fused_roles(h->kp, c->R, c->kd, c->has_base) && build_tile_spec(h->kp, c->R, c->kd, c->has_base, c->spec);
        if (c->ok) {     // 64 bytes of device memory, here so that no step ever allocates; without them: static assignment
            if (cudaMalloc(&c->sched, LaneCtx::kSchedSlots * sizeof(int)) != cudaSuccess ||
                cudaMemset(c->sched, 0, LaneCtx::kSchedSlots * sizeof(int)) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
                cudaGetLastError();
                if (c->sched) cudaFree(c->sched);
                c->sched = nullptr;
            }
        }
        h->lane_ctx = c;
    }
    return static_cast<LaneCtx *>(h->lane_ctx);
}
const LaneCtx *ctx_of(const irlosc_handle *h) { return ctx_of(const_cast<irlosc_handle *>(h)); }

struct LaneEntry {
    int kd;
    bool has_base;
    int threads;
    bool staged;
    const void *fn;
    int fix_bytes;
    const char *name;
};

template <int KD, bool HB, int NT, bool ST>
LaneEntry lentry(const char *name) {
    return LaneEntry{KD, HB, NT, ST, (const void *)osc_step_lane<KD, HB, NT, ST>, (int)((sizeof(fused::WarpFix<KD, HB>) + 15) & ~size_t(15)), name};
}

// A thread owns an instance, so the warps per SM are bounded by registers: 65 536 / threads per thread.  More warps
// hide more latency, fewer registers spill more; the table keeps the candidates that were measured (profiles/).
// "tma": groups staged through a per-warp shared-memory ring by TMA bulk copies; "ldg": direct coalesced loads.
const LaneEntry *lane_table(int *count) {
    static const LaneEntry t[] = {
        lentry<3, true, 224, true>("osc_step_lane<kd3,base,t224,tma>"),   lentry<3, true, 256, true>("osc_step_lane<kd3,base,t256,tma>"),
        lentry<3, false, 224, true>("osc_step_lane<kd3,t224,tma>"),       lentry<3, false, 256, true>("osc_step_lane<kd3,t256,tma>"),
        lentry<3, true, 224, false>("osc_step_lane<kd3,base,t224,ldg>"),  lentry<3, true, 256, false>("osc_step_lane<kd3,base,t256,ldg>"),
        lentry<3, false, 256, false>("osc_step_lane<kd3,t256,ldg>"),
        lentry<6, false, 256, false>("osc_step_lane<kd6,t256,ldg>"),
        lentry<6, true, 256, false>("osc_step_lane<kd6,base,t256,ldg>"),
    };
    *count = (int)(sizeof t / sizeof t[0]);
    return t;
}

constexpr size_t kLaneSmem = 227 * 1024;

// Pair kernels (osc_pair.cuh): two lanes per instance.  Registers per thread = 65 536 / threads.
struct PairEntry {
    int kd;
    bool has_base;
    int threads;
    const void *fn;
    int fix_bytes;
    const char *name;
};
template <int KD, bool HB, int NT>
PairEntry pentry(const char *name) {
    return PairEntry{KD, HB, NT, (const void *)pair::osc_step_pair<KD, HB, NT>, (int)((sizeof(fused::WarpFix<KD, HB>) + 15) & ~size_t(15)), name};
}
const PairEntry *pair_table(int *count) {
    // 8 warps of 255 registers: measured against 10, 12 and 16 warps (204 / 168 / 128 registers, spills) at every batch
    // size and layout, 8 won everywhere (profiles/r02_summary.md).
    static const PairEntry t[] = {
        pentry<3, true, 256>("osc_step_pair<kd3,base,t256>"), pentry<3, false, 256>("osc_step_pair<kd3,t256>"),
        pentry<6, true, 256>("osc_step_pair<kd6,base,t256>"), pentry<6, false, 256>("osc_step_pair<kd6,t256>"),
    };
    *count = (int)(sizeof t / sizeof t[0]);
    return t;
}

int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Threads per CTA (one CTA per SM).  Measured (profiles/): 8 warps at the full 255 registers beat 10 - 14 warps with
// spills in every configuration.  A warp owns a 32-instance tile, so a batch is ceil(tiles / (SMs x warps)) waves;
// 7 warps are taken when that needs no more waves than 8 (B = 65 536 on 148 SMs: 1.73 waves of 8 or 1.98 of 7).
// IRLOSC_LANE_THREADS overrides for experiments.
int lane_threads_for(int64_t B, int sms) {
    const int64_t tiles = (B + kTile - 1) / kTile;
    const int64_t w8 = (tiles + (int64_t)sms * 8 - 1) / ((int64_t)sms * 8), w7 = (tiles + (int64_t)sms * 7 - 1) / ((int64_t)sms * 7);
    return (w7 <= w8 && tiles > (int64_t)sms * 7) ? 224 : 256;
}

// B_whole: the batch the caller handed in (the host entry point launches it in chunks): the kernel choice follows it,
// so that a batch gives the same bits whichever entry point it came through.
int32_t launch_lane(irlosc_handle *h, LaneCtx &c, int64_t B, const irlosc_tiles_io &io, cudaStream_t st, int64_t B_whole = -1) {
    const KParams &P = h->kp;
    int cnt = 0;
    const LaneEntry *t = lane_table(&cnt), *e = nullptr;
    const int want = env_int("IRLOSC_LANE_THREADS", lane_threads_for(B, std::max(1, h->sm_count - h->sm_margin)));
    // 3-row arm devices: groups staged by TMA bulk copies (measured -11 % vs direct loads); 6-row arm devices: direct
    // loads (their groups are larger: the ring leaves room for fewer warps, measured +4 .. +14 %)
    bool staged = env_int("IRLOSC_LANE_STAGED", c.kd == 3 ? 1 : 0) != 0;
    int stage_bytes = 0;
    for (int g = 0; g < kGroups; ++g) stage_bytes = std::max(stage_bytes, (c.spec.gbase[g + 1] - c.spec.gbase[g]) * kTile * 8);
    int n_stages = 0, warp_bytes = 0;
    for (int pass = 0; pass < 2 && !e; ++pass, staged = false) {
        for (int i = 0; i < cnt; ++i)
            if (t[i].kd == c.kd && t[i].has_base == c.has_base && t[i].staged == staged &&
                (e == nullptr || abs(t[i].threads - want) < abs(e->threads - want))) e = &t[i];
        if (!e) continue;
        const int warps_ = e->threads / 32;
        const int rest_bytes = ((32 * P.n_ctrl * 8 + 15) & ~15) + e->fix_bytes;
        n_stages = 0;
        if (e->staged) {      // ring depth: 3 preferred (two groups in flight while one is consumed), 2 if that is all that fits
            n_stages = std::min(4, std::max(2, env_int("IRLOSC_LANE_STAGES", 3)));
            while (n_stages > 2 && (size_t)warps_ * ((size_t)n_stages * stage_bytes + 64 + rest_bytes) > kLaneSmem) --n_stages;
        }
        warp_bytes = (e->staged ? n_stages * stage_bytes + 64 : 0) + rest_bytes;
        if ((size_t)warps_ * warp_bytes > kLaneSmem) e = nullptr;      // the ring does not fit: direct loads
    }
    if (!e) return fail(IRLOSC_ERR_INVALID, "no lane kernel for kd=%d base=%d", c.kd, (int)c.has_base);
    LaneArgs A;
    memset(&A, 0, sizeof A);
    A.tiles = io.tiles;
    A.n_entries = c.spec.n_entries;
    A.pf = env_int("IRLOSC_LANE_PREFETCH", 2);      // L2 prefetch two groups ahead (measured: -7 % at k = 7)
    for (int g = 0; g <= kGroups; ++g) A.gbase[g] = c.spec.gbase[g];
    A.target_vel = io.target_vel;
    A.u_all = io.u_all; A.ctrl = io.ctrl; A.status = io.status;
    auto al16 = [](const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    stream::Gather G;
    memset(&G, 0, sizeof G);
    if (io.n_gather < 0 || io.n_gather > IRLOSC_MAX_PEERS) return fail(IRLOSC_ERR_INVALID, "n_gather=%d outside 0..%d", io.n_gather, IRLOSC_MAX_PEERS);
    G.n_gather = io.n_gather;
    G.gather_offset = io.gather_offset;
    G.ctrl_mc = io.ctrl_multicast;
    G.ctrl_vec = al16(io.ctrl) && ((io.gather_offset * (int64_t)P.n_ctrl) % 2 == 0) && al16(io.ctrl_multicast);
    for (int g = 0; g < io.n_gather; ++g) {
        if (!io.ctrl_gather[g]) return fail(IRLOSC_ERR_INVALID, "ctrl_gather[%d] is null", g);
        G.ctrl_gather[g] = io.ctrl_gather[g];
        G.ctrl_vec = G.ctrl_vec && al16(io.ctrl_gather[g]);
    }
    const int sms = std::max(1, h->sm_count - h->sm_margin);
    // Which kernel.  The pair kernel walks half the dependent chain per instance with half the state per thread:
    // it wins whenever the batch is one wave of half tiles or less (latency of ONE half tile instead of one tile:
    // B <= 16 x 8 warps x SMs), and for 6-row arm devices at every size (no spills; measured -13 % at B = 65 536,
    // -27 % at 262 144).  3-row layouts above one wave: the TMA-staged lane kernel (measured 0.050 vs 0.060 ms).
    // IRLOSC_PAIR (experiments) overrides: threads per CTA of the pair kernel, 0 = lane kernel.
    const int64_t n_half = (B + pair::kHalf - 1) / pair::kHalf;
    const int64_t n_half_whole = ((B_whole < 0 ? B : B_whole) + pair::kHalf - 1) / pair::kHalf;
    int pair_threads = 0;
    if (h->tile_kernel == IRLOSC_TILES_PAIR || (h->tile_kernel == IRLOSC_TILES_AUTO && (c.kd > 3 || n_half_whole <= (int64_t)sms * 8)))
        pair_threads = 256;
    if (h->tile_kernel == IRLOSC_TILES_AUTO) pair_threads = env_int("IRLOSC_PAIR", pair_threads);
    if (pair_threads > 0) {
        int pc = 0;
        const PairEntry *pt = pair_table(&pc), *pe = nullptr;
        for (int i = 0; i < pc; ++i)
            if (pt[i].kd == c.kd && pt[i].has_base == c.has_base && (pe == nullptr || abs(pt[i].threads - pair_threads) < abs(pe->threads - pair_threads)))
                pe = &pt[i];
        if (!pe) return fail(IRLOSC_ERR_INVALID, "no pair kernel for kd=%d base=%d", c.kd, (int)c.has_base);
        static bool pair_ready = false;
        if (!pair_ready) {
            for (int i = 0; i < pc; ++i)
                CUDA_TRY(cudaFuncSetAttribute(pt[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLaneSmem));
            pair_ready = true;
        }
        // One wave or less: spread the half tiles over the SMs as CTAs of 1 .. 8 warps (a warp alone on its
        // scheduler walks its chain faster than two sharing one) instead of filling a few SMs with 8 warps each.
        int pw = pe->threads / 32;
        if (env_int("IRLOSC_PAIR_SPREAD", 1) != 0) {
            const int64_t per_sm = (n_half + sms - 1) / sms;
            if (per_sm < pw) pw = (int)std::max<int64_t>(1, per_sm);
        }
        const int pthreads = pw * 32;
        // tickets pay when a warp has several half tiles to walk (IRLOSC_PAIR_TICKETS: experiments)
        if (env_int("IRLOSC_PAIR_TICKETS", 1) != 0 && n_half > (int64_t)sms * pw && c.sched) {
            int slot = -1;
            for (int i = 0; i < c.sched_used && slot < 0; ++i)
                if (c.sched_stream[i] == st) slot = i;
            if (slot < 0 && c.sched_used < LaneCtx::kSchedSlots) {
                slot = c.sched_used++;
                c.sched_stream[slot] = st;
            }
            if (slot >= 0) A.sched = c.sched + slot;     // more streams than counters: this launch assigns statically
        }
        int pair_warp_bytes = ((pair::kHalf * P.n_ctrl * 8 + 15) & ~15) + pe->fix_bytes;
        const int pgrid = (int)std::min<int64_t>((n_half + pw - 1) / pw, (int64_t)sms);
        void *pargs[] = {(void *)&P, (void *)&A, (void *)&B, (void *)&c.R, (void *)&G, (void *)&pair_warp_bytes};
        cudaError_t perr = cudaLaunchKernel(pe->fn, dim3(pgrid), dim3(pthreads), pargs, (size_t)pw * pair_warp_bytes, st);
        if (perr != cudaSuccess) return fail(IRLOSC_ERR_CUDA, "pair kernel launch: %s", cudaGetErrorString(perr));
        h->launches += 1;
        h->last_kernel = pe->name;
        return IRLOSC_OK;
    }
    const int warps = e->threads / 32;
    const size_t smem = (size_t)warps * warp_bytes;
    static bool ready = false;
    if (!ready) {
        for (int i = 0; i < cnt; ++i)
            CUDA_TRY(cudaFuncSetAttribute(t[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLaneSmem));
        ready = true;
    }
    const int64_t n_tiles = (B + kTile - 1) / kTile;
    const int grid = (int)std::min<int64_t>((n_tiles + warps - 1) / warps, (int64_t)sms);
    void *args[] = {(void *)&P, (void *)&A, (void *)&B, (void *)&c.R, (void *)&G, (void *)&warp_bytes, (void *)&stage_bytes,
                    (void *)&n_stages};
    cudaError_t err = cudaLaunchKernel(e->fn, dim3(grid), dim3(e->threads), args, smem, st);
    if (err != cudaSuccess) return fail(IRLOSC_ERR_CUDA, "lane kernel launch: %s", cudaGetErrorString(err));
    h->launches += 1;
    h->last_kernel = e->name;
    return IRLOSC_OK;
}

int32_t check_tiles_io(const irlosc_tiles_io *io) {
    if (!io) return fail(IRLOSC_ERR_INVALID, "io is null");
    if (!io->tiles || !io->ctrl) return fail(IRLOSC_ERR_INVALID, "a required array (tiles, ctrl) is null");
    return IRLOSC_OK;
}

}  // namespace

void irlosc::lane_destroy(irlosc_handle *h) {
    LaneCtx *c = static_cast<LaneCtx *>(h->lane_ctx);
    if (!c) return;
    for (int s = 0; s < kPipeDepth; ++s) {
        for (int i = 0; i < 16; ++i)
            if (c->stage[s].buf[i]) cudaFree(c->stage[s].buf[i]);
        if (c->stage[s].stream) cudaStreamDestroy(c->stage[s].stream);
    }
    if (c->sched) cudaFree(c->sched);
    delete c;
    h->lane_ctx = nullptr;
}

extern "C" int32_t irlosc_tile_entries(const irlosc_handle *h) {
    if (!h) return 0;
    const LaneCtx *c = ctx_of(h);
    return (c && c->ok) ? c->spec.n_entries : 0;
}

extern "C" int32_t irlosc_tile_spec(const irlosc_handle *h, irlosc_tile_entry *out, int32_t capacity) {
    if (!h) return 0;
    const LaneCtx *c = ctx_of(h);
    if (!c || !c->ok) return 0;
    for (int e = 0; out && e < c->spec.n_entries && e < capacity; ++e) out[e] = c->spec.e[e];
    return c->spec.n_entries;
}

extern "C" int64_t irlosc_tiles_doubles(const irlosc_handle *h, int64_t B) {
    const int E = irlosc_tile_entries(h);
    return B <= 0 ? 0 : ((B + kTile - 1) / kTile) * (int64_t)kTile * E;
}

static int32_t pack_table_for(irlosc_handle *h, const irlosc_io *io, const LaneCtx *&c, PackTable &T) {
    if (!h) return fail(IRLOSC_ERR_INVALID, "handle is null");
    c = ctx_of(h);
    if (!c) return fail(IRLOSC_ERR_NOMEM, "out of host memory");
    if (!c->ok) return fail(IRLOSC_ERR_INVALID, "this controller has no tile layout (needs the declared DualUR5 topology with two "
                                                "equally masked arm devices, optionally the base)");
    KIo k;
    int32_t rc = resolve_io(h, io, k, false);
    if (rc != IRLOSC_OK) return rc;
    return build_pack_table(h->kp, k, c->spec, T);
}

extern "C" int32_t irlosc_pack_tiles(irlosc_handle *h, int64_t B, const irlosc_io *io, double *tiles, void *cuda_stream) {
    const LaneCtx *c = nullptr;
    PackTable T;
    int32_t rc = pack_table_for(h, io, c, T);
    if (rc != IRLOSC_OK) return rc;
    if (B < 0) return fail(IRLOSC_ERR_INVALID, "B=%lld is negative", (long long)B);
    if (B == 0) return IRLOSC_OK;
    if (!tiles) return fail(IRLOSC_ERR_INVALID, "tiles is null");
    const int64_t n_tiles = (B + kTile - 1) / kTile;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)h->sm_count * 8);
    pack_tiles_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(T, tiles, B);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(IRLOSC_ERR_CUDA, "pack kernel launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    h->last_kernel = "pack_tiles_kernel";
    return IRLOSC_OK;
}

extern "C" int32_t irlosc_pack_tiles_host(irlosc_handle *h, int64_t B, const irlosc_io *io, double *tiles) {
    const LaneCtx *c = nullptr;
    PackTable T;
    int32_t rc = pack_table_for(h, io, c, T);
    if (rc != IRLOSC_OK) return rc;
    if (B < 0) return fail(IRLOSC_ERR_INVALID, "B=%lld is negative", (long long)B);
    if (B == 0) return IRLOSC_OK;
    if (!tiles) return fail(IRLOSC_ERR_INVALID, "tiles is null");
    const int E = T.n_entries;
    const int64_t n_tiles = (B + kTile - 1) / kTile;
    for (int64_t t = 0; t < n_tiles; ++t) {
        double *dst = tiles + t * (int64_t)E * kTile;
        for (int l = 0; l < kTile; ++l) {
            const int64_t inst = std::min<int64_t>(t * kTile + l, B - 1);
            for (int e = 0; e < E; ++e) dst[(size_t)e * kTile + l] = pack_fetch(T, e, inst);
        }
    }
    return IRLOSC_OK;
}

extern "C" int32_t irlosc_step_tiles(irlosc_handle *h, int64_t B, const irlosc_tiles_io *io, void *cuda_stream) {
    if (!h) return fail(IRLOSC_ERR_INVALID, "handle is null");
    if (B < 0) return fail(IRLOSC_ERR_INVALID, "B=%lld is negative", (long long)B);
    if (B == 0) return IRLOSC_OK;
    int32_t rc = check_tiles_io(io);
    if (rc != IRLOSC_OK) return rc;
    LaneCtx *c = ctx_of(h);
    if (!c) return fail(IRLOSC_ERR_NOMEM, "out of host memory");
    if (!c->ok) return fail(IRLOSC_ERR_INVALID, "this controller has no tile layout");
    return launch_lane(h, *c, B, *io, (cudaStream_t)cuda_stream);
}

extern "C" int32_t irlosc_step_tiles_host(irlosc_handle *h, int64_t B, const irlosc_tiles_io *io) {
    if (!h) return fail(IRLOSC_ERR_INVALID, "handle is null");
    if (B < 0) return fail(IRLOSC_ERR_INVALID, "B is negative");
    if (B == 0) return IRLOSC_OK;
    int32_t rc = check_tiles_io(io);
    if (rc != IRLOSC_OK) return rc;
    if (io->n_gather != 0 || io->ctrl_multicast) return fail(IRLOSC_ERR_INVALID, "the fused gather is only available with irlosc_step_tiles (device pointers)");
    LaneCtx *c = ctx_of(h);
    if (!c) return fail(IRLOSC_ERR_NOMEM, "out of host memory");
    if (!c->ok) return fail(IRLOSC_ERR_INVALID, "this controller has no tile layout");
    const KParams &P = h->kp;
    CUDA_TRY(cudaSetDevice(h->device));
    for (int s = 0; s < kPipeDepth; ++s)
        if (!c->stage[s].stream) CUDA_TRY(cudaStreamCreateWithFlags(&c->stage[s].stream, cudaStreamNonBlocking));
    const int E = c->spec.n_entries;
    const size_t D = P.D, n = P.n;
    const int64_t chunk = std::max<int64_t>(kTile, std::min<int64_t>((2 * h->host_chunk) / kTile * kTile, (B + kTile - 1) / kTile * kTile));
    int turn = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, ++turn) {
        const int64_t nb = std::min<int64_t>(chunk, B - b0);
        const int64_t nt = (nb + kTile - 1) / kTile;
        Staging &S = c->stage[turn % kPipeDepth];
        rc = ensure_cap(S, 0, (size_t)(chunk / kTile) * E * kTile * sizeof(double));
        if (rc == IRLOSC_OK && io->target_vel) rc = ensure_cap(S, 1, (size_t)chunk * 6 * D * sizeof(double));
        if (rc == IRLOSC_OK) rc = ensure_cap(S, 2, (size_t)chunk * P.n_ctrl * sizeof(double));
        if (rc == IRLOSC_OK && io->u_all) rc = ensure_cap(S, 3, (size_t)chunk * n * sizeof(double));
        if (rc == IRLOSC_OK && io->status) rc = ensure_cap(S, 4, (size_t)chunk);
        if (rc != IRLOSC_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(S.buf[0], io->tiles + (b0 / kTile) * (int64_t)E * kTile, (size_t)nt * E * kTile * sizeof(double),
                                 cudaMemcpyHostToDevice, S.stream));
        if (io->target_vel)
            CUDA_TRY(cudaMemcpyAsync(S.buf[1], io->target_vel + (size_t)b0 * 6 * D, (size_t)nb * 6 * D * sizeof(double),
                                     cudaMemcpyHostToDevice, S.stream));
        irlosc_tiles_io dk;
        memset(&dk, 0, sizeof dk);
        dk.tiles = (const double *)S.buf[0];
        dk.target_vel = io->target_vel ? (const double *)S.buf[1] : nullptr;
        dk.ctrl = (double *)S.buf[2];
        dk.u_all = io->u_all ? (double *)S.buf[3] : nullptr;
        dk.status = io->status ? (uint8_t *)S.buf[4] : nullptr;
        rc = launch_lane(h, *c, nb, dk, S.stream, B);
        if (rc != IRLOSC_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(io->ctrl + (size_t)b0 * P.n_ctrl, dk.ctrl, (size_t)nb * P.n_ctrl * sizeof(double),
                                 cudaMemcpyDeviceToHost, S.stream));
        if (io->u_all)
            CUDA_TRY(cudaMemcpyAsync(io->u_all + (size_t)b0 * n, dk.u_all, (size_t)nb * n * sizeof(double), cudaMemcpyDeviceToHost, S.stream));
        if (io->status) CUDA_TRY(cudaMemcpyAsync(io->status + b0, dk.status, (size_t)nb, cudaMemcpyDeviceToHost, S.stream));
    }
    for (int s = 0; s < kPipeDepth; ++s) CUDA_TRY(cudaStreamSynchronize(c->stage[s].stream));
    return IRLOSC_OK;
}

// C ABI of the batched OSC (include/irlosc.h): parameter validation / flattening,
// kernel dispatch, and the host-buffer pipeline.  No torch, no Python.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#include "../../include/irlosc.h"
#include "irlosc_internal.h"
#include "irlosc_build.h"
#include "osc_generic.cuh"
#include "osc_dispatch.cuh"

using namespace irlosc;

// ------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

int32_t irlosc::fail(int32_t rc, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return rc;
}

extern "C" const char *irlosc_last_error(void) { return g_last_error.c_str(); }
extern "C" int32_t irlosc_abi_version(void) { return IRLOSC_ABI_VERSION; }

extern "C" int32_t irlosc_create(const irlosc_params *params, irlosc_handle **out) {
    if (!params || !out) return fail(IRLOSC_ERR_INVALID, "null argument");
    *out = nullptr;
    irlosc_handle *h = new (std::nothrow) irlosc_handle();
    if (!h) return fail(IRLOSC_ERR_NOMEM, "out of host memory");
    h->user = *params;
    int32_t rc = build_kparams(*params, h->kp);
    if (rc != IRLOSC_OK) { delete h; return rc; }
    cudaError_t e = cudaGetDevice(&h->device);
    if (e != cudaSuccess) {
        delete h;
        return fail(IRLOSC_ERR_CUDA, "cudaGetDevice failed: %s (is a CUDA device present?)", cudaGetErrorString(e));
    }
    e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    if (e != cudaSuccess) { delete h; return fail(IRLOSC_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e)); }
    e = cudaFuncSetAttribute(osc_step_generic, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(GenericSmem) * kGenericWarps));
    if (e != cudaSuccess) { delete h; return fail(IRLOSC_ERR_CUDA, "cudaFuncSetAttribute(generic): %s", cudaGetErrorString(e)); }
    rc = tiled_prepare();
    if (rc != IRLOSC_OK) { delete h; return fail(IRLOSC_ERR_CUDA, "tiled kernel setup failed"); }
    *out = h;
    return IRLOSC_OK;
}

extern "C" int32_t irlosc_destroy(irlosc_handle *h) {
    if (!h) return IRLOSC_OK;
    for (int s = 0; s < kPipeDepth; ++s) {
        for (int i = 0; i < 16; ++i) {
            if (h->stage[s].buf[i]) cudaFree(h->stage[s].buf[i]);
            if (h->fstage[s].buf[i]) cudaFree(h->fstage[s].buf[i]);
        }
        if (h->stage[s].stream) cudaStreamDestroy(h->stage[s].stream);
        if (h->fstage[s].stream) cudaStreamDestroy(h->fstage[s].stream);
    }
    if (h->small_host) cudaFreeHost(h->small_host);
    if (h->small_dev) cudaFree(h->small_dev);
    lane_destroy(h);
    delete h;
    return IRLOSC_OK;
}

extern "C" int32_t irlosc_num_task_rows(const irlosc_handle *h) { return h ? h->kp.k : -1; }
extern "C" int32_t irlosc_num_ctrl(const irlosc_handle *h) { return h ? h->kp.n_ctrl : -1; }
extern "C" int64_t irlosc_kernel_launches(const irlosc_handle *h) { return h ? h->launches : -1; }
extern "C" const char *irlosc_last_kernel(const irlosc_handle *h) { return h ? h->last_kernel : "none"; }

extern "C" int32_t irlosc_set_sm_margin(irlosc_handle *h, int32_t sms) {
    if (!h || sms < 0 || sms >= h->sm_count) return fail(IRLOSC_ERR_INVALID, "sm margin must be in 0..%d", h ? h->sm_count - 1 : 0);
    h->sm_margin = sms;
    return IRLOSC_OK;
}

extern "C" int32_t irlosc_set_kernel(irlosc_handle *h, int32_t which) {
    if (!h || which < 0 || which > 9) return fail(IRLOSC_ERR_INVALID, "kernel selector must be 0 (auto), 1 (generic) or 2+v (tiled variant v)");
    h->kernel_choice = which;
    return IRLOSC_OK;
}

extern "C" int32_t irlosc_set_tile_kernel(irlosc_handle *h, int32_t which) {
    if (!h || which < IRLOSC_TILES_AUTO || which > IRLOSC_TILES_PAIR)
        return fail(IRLOSC_ERR_INVALID, "tile kernel selector must be IRLOSC_TILES_AUTO, _LANE or _PAIR");
    h->tile_kernel = which;
    return IRLOSC_OK;
}

// ------------------------------------------------------------------ step (device pointers)
int32_t irlosc::resolve_io(const irlosc_handle *h, const irlosc_io *io, KIo &k, bool need_outputs) {
    const KParams &P = h->kp;
    if (!io) return fail(IRLOSC_ERR_INVALID, "io is null");
    if (!io->M || !io->J || !io->dq || !io->ee_xyz || !io->ee_quat || !io->target_xyz || !io->target_quat ||
        (need_outputs && !io->ctrl))
        return fail(IRLOSC_ERR_INVALID, "a required array (M, J, dq, ee_xyz, ee_quat, target_xyz, target_quat, ctrl) is null");
    if (P.use_g && !io->bias) return fail(IRLOSC_ERR_INVALID, "use_g is set but bias is null");
    if (P.admittance && (!io->ft_xmat || !io->ft_raw))
        return fail(IRLOSC_ERR_INVALID, "admittance is set but ft_xmat / ft_raw is null");
    const int32_t mrc = resolve_m_layout(P, *io, k);
    if (mrc != IRLOSC_OK) return mrc;
    k.J = io->J; k.j_layout = io->j_layout;
    k.ldj = io->ldj ? io->ldj : P.n;
    if (k.ldj < P.n) return fail(IRLOSC_ERR_INVALID, "ldj=%d < n=%d", k.ldj, P.n);
    if (io->j_layout == IRLOSC_J_ROWS) k.j_stride = io->j_stride ? io->j_stride : (int64_t)k.ldj * P.k;
    else if (io->j_layout == IRLOSC_J_FULL6) k.j_stride = io->j_stride ? io->j_stride : (int64_t)k.ldj * 6 * P.D;
    else return fail(IRLOSC_ERR_INVALID, "unknown j_layout %d", io->j_layout);
    k.dq = io->dq; k.bias = io->bias; k.ee_xyz = io->ee_xyz; k.ee_quat = io->ee_quat;
    k.target_xyz = io->target_xyz; k.target_quat = io->target_quat; k.target_vel = io->target_vel;
    k.max_vel = io->max_vel; k.ft_xmat = io->ft_xmat; k.ft_raw = io->ft_raw;
    k.u_all = io->u_all; k.ctrl = io->ctrl; k.status = io->status;
    if (io->n_gather < 0 || io->n_gather > IRLOSC_MAX_PEERS) return fail(IRLOSC_ERR_INVALID, "n_gather=%d outside 0..%d", io->n_gather, IRLOSC_MAX_PEERS);
    k.n_gather = io->n_gather; k.gather_offset = io->gather_offset;
    for (int g = 0; g < IRLOSC_MAX_PEERS; ++g) {
        k.ctrl_gather[g] = g < io->n_gather ? io->ctrl_gather[g] : nullptr;
        if (g < io->n_gather && !k.ctrl_gather[g]) return fail(IRLOSC_ERR_INVALID, "ctrl_gather[%d] is null", g);
    }
    k.ctrl_mc = io->ctrl_multicast;
    return IRLOSC_OK;
}

// Kernel selection for per-variable arrays (irlosc_set_kernel): 0 auto, 1 generic, 2 + v tree-table variant v,
// 9 streaming.  (State held as batch-interleaved tiles goes to the lane kernel, irlosc_step_tiles.)
//   auto, DualUR5 topology declared: 3-row arm devices (gain_test) -> tree-sparse 4-lane kernel when
//   the layout is one it stages (tight packed / dense / qM, row-stacked J), 6-row arm devices -> streaming
//   kernel; anything the tree kernel does not stage (strided views, full-6 J) -> streaming kernel;
//   check_topology -> the tree kernel (it reads every entry) or generic.  No topology: generic.
constexpr int kKernelStream = 9;

static int32_t launch_step(irlosc_handle *h, int64_t B, const KIo &k, cudaStream_t st) {
    if (B == 0) return IRLOSC_OK;
    const bool stream_ok = stream_supported(h, k);
    if (k.m_layout == IRLOSC_M_QM) {
        // MuJoCo's sparse qM: the tree-sparse kernel's qM instantiations when it stages this layout (tight stride,
        // row-stacked J) and the arm devices have 3 rows; otherwise the streaming kernel's copy plan
        const int v = (h->kernel_choice >= 2 && h->kernel_choice != kKernelStream) ? h->kernel_choice - 2 : 0;
        const bool tree_ok = h->kernel_choice != 1 && h->kernel_choice != kKernelStream && tiled_supported(h->kp, k, v);
        if (tree_ok && (h->kernel_choice >= 2 || !stream_preferred(h))) {
            cudaError_t e = tiled_launch(h->kp, k, B, h->sm_count - h->sm_margin, st, &h->last_kernel, v);
            if (e != cudaSuccess) return fail(IRLOSC_ERR_CUDA, "tiled kernel launch: %s", cudaGetErrorString(e));
            h->launches += 1;
            return IRLOSC_OK;
        }
        if (h->kernel_choice >= 2 && h->kernel_choice != kKernelStream)
            return fail(IRLOSC_ERR_INVALID, "no tree-sparse kernel variant %d for IRLOSC_M_QM with this controller / stride", v);
        if (!stream_ok || h->kernel_choice == 1)
            return fail(IRLOSC_ERR_INVALID, "IRLOSC_M_QM needs the declared DualUR5 topology (check_topology off) and "
                                            "kernel selector 0, 9 or 2 + v");
        const int32_t rc = stream_launch(h, B, k, st);
        return rc == kErrPlanTooLarge ? IRLOSC_ERR_INVALID : rc;
    }
    if (h->kernel_choice == kKernelStream) {
        if (!stream_ok) return fail(IRLOSC_ERR_INVALID, "streaming kernel requested but the DualUR5 topology is not declared (or check_topology is set)");
        const int32_t rc = stream_launch(h, B, k, st);
        return rc == kErrPlanTooLarge ? IRLOSC_ERR_INVALID : rc;
    }
    bool use_tiled = false;
    const int variant = h->kernel_choice >= 2 ? h->kernel_choice - 2 : 0;
    if (h->kernel_choice != 1) use_tiled = tiled_supported(h->kp, k, variant);
    if (h->kernel_choice == 0 && stream_ok && (stream_preferred(h) || !use_tiled)) {
        const int32_t rc = stream_launch(h, B, k, st);
        if (rc != kErrPlanTooLarge) return rc;
        // this configuration's copy plan does not fit (three 6-row devices with admittance): the kernels below stage
        // whole records instead
    }
    if (h->kernel_choice >= 2 && !use_tiled)
        return fail(IRLOSC_ERR_INVALID, "tiled kernel requested but this shape/layout is not supported (n=%d k=%d)", h->kp.n, h->kp.k);
    if (use_tiled) {
        cudaError_t e = tiled_launch(h->kp, k, B, h->sm_count - h->sm_margin, st, &h->last_kernel, variant);
        if (e != cudaSuccess) return fail(IRLOSC_ERR_CUDA, "tiled kernel launch: %s", cudaGetErrorString(e));
    } else {
        const int64_t blocks_needed = (B + kGenericWarps - 1) / kGenericWarps;
        const int grid = (int)std::min<int64_t>(blocks_needed, (int64_t)(h->sm_count - h->sm_margin) * 8);
        osc_step_generic<<<grid, kGenericWarps * 32, sizeof(GenericSmem) * kGenericWarps, st>>>(h->kp, k, B);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(IRLOSC_ERR_CUDA, "generic kernel launch: %s", cudaGetErrorString(e));
        h->last_kernel = "osc_step_generic";
    }
    h->launches += 1;
    return IRLOSC_OK;
}

extern "C" int32_t irlosc_step(irlosc_handle *h, int64_t B, const irlosc_io *io, void *cuda_stream) {
    if (!h) return fail(IRLOSC_ERR_INVALID, "handle is null");
    if (B < 0) return fail(IRLOSC_ERR_INVALID, "B=%lld is negative", (long long)B);
    if (B == 0) return IRLOSC_OK;   // empty batch: nothing to read or write (array pointers may be null)
    KIo k;
    int32_t rc = resolve_io(h, io, k, true);
    if (rc != IRLOSC_OK) return rc;
    return launch_step(h, B, k, (cudaStream_t)cuda_stream);
}

// ------------------------------------------------------------------ calc_error
__global__ void calc_error_kernel(const KParams P, const double *ee_xyz, const double *ee_quat,
                                  const double *t_xyz, const double *t_quat, double *err, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int d = (int)(i % P.D);
    double ee[3], eq[4], tx[3], tq[4], u[6];
    for (int c = 0; c < 3; ++c) { ee[c] = ee_xyz[i * 3 + c]; tx[c] = t_xyz[i * 3 + c]; }
    for (int c = 0; c < 4; ++c) { eq[c] = ee_quat[i * 4 + c]; tq[c] = t_quat[i * 4 + c]; }
    device_pose_error(P.dev[d], ee, eq, tx, tq, u);
    for (int c = 0; c < 6; ++c) err[i * 6 + c] = u[c];
}

extern "C" int32_t irlosc_calc_error(irlosc_handle *h, int64_t B, const double *ee_xyz, const double *ee_quat,
                                     const double *target_xyz, const double *target_quat, double *err,
                                     void *cuda_stream) {
    if (!h) return fail(IRLOSC_ERR_INVALID, "handle is null");
    if (B < 0) return fail(IRLOSC_ERR_INVALID, "B is negative");
    if (!ee_xyz || !ee_quat || !target_xyz || !target_quat || !err) return fail(IRLOSC_ERR_INVALID, "null array");
    if (B == 0) return IRLOSC_OK;
    const int64_t total = B * h->kp.D;
    const int threads = 128;
    calc_error_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        h->kp, ee_xyz, ee_quat, target_xyz, target_quat, err, total);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(IRLOSC_ERR_CUDA, "calc_error launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    h->last_kernel = "calc_error_kernel";
    return IRLOSC_OK;
}

// ------------------------------------------------------------------ host-buffer pipeline
extern "C" int32_t irlosc_host_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(IRLOSC_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaHostAlloc(ptr, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault));
    return IRLOSC_OK;
}

extern "C" int32_t irlosc_host_free(void *ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return IRLOSC_OK;
}

int32_t irlosc::ensure_cap(Staging &s, int slot, size_t bytes) {
    if (bytes <= s.cap[slot]) return IRLOSC_OK;
    if (s.buf[slot]) { CUDA_TRY(cudaFree(s.buf[slot])); s.buf[slot] = nullptr; s.cap[slot] = 0; }
    CUDA_TRY(cudaMalloc(&s.buf[slot], bytes));
    s.cap[slot] = bytes;
    return IRLOSC_OK;
}

// Small batches (the drop-in OSC.generate calls this with B = 1 every millisecond, examples/gain_test.py:143-147): the
// latency is the number of driver calls, so all inputs travel in ONE pinned staging block with one H2D copy, the
// outputs come back in one D2H copy, one stream synchronisation.
constexpr int64_t kSmallBatch = 64;
struct HostIn { const double *src; size_t per; };

static int32_t step_host_small(irlosc_handle *h, int64_t B, const KIo &hk, const HostIn *ins) {
    const KParams &P = h->kp;
    Staging &S = h->stage[0];
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    size_t in_off[12], at = 0;
    for (int i = 0; i < 12; ++i) {
        in_off[i] = at;
        if (ins[i].src) at += up(ins[i].per * (size_t)B * sizeof(double));
    }
    const size_t in_bytes = at;
    const size_t o_ctrl = at; at += up((size_t)B * P.n_ctrl * sizeof(double));
    const size_t o_uall = at; if (hk.u_all) at += up((size_t)B * P.n * sizeof(double));
    const size_t o_stat = at; if (hk.status) at += up((size_t)B);
    const size_t total = at;
    if (total > h->small_cap) {
        if (h->small_host) { CUDA_TRY(cudaFreeHost(h->small_host)); h->small_host = nullptr; }
        if (h->small_dev) { CUDA_TRY(cudaFree(h->small_dev)); h->small_dev = nullptr; }
        h->small_cap = 0;
        CUDA_TRY(cudaHostAlloc(&h->small_host, total, cudaHostAllocDefault));
        CUDA_TRY(cudaMalloc(&h->small_dev, total));
        h->small_cap = total;
    }
    unsigned char *hb = static_cast<unsigned char *>(h->small_host), *db = static_cast<unsigned char *>(h->small_dev);
    KIo dk = hk;
    const double *dptr[12];
    for (int i = 0; i < 12; ++i) {
        dptr[i] = nullptr;
        if (!ins[i].src) continue;
        memcpy(hb + in_off[i], ins[i].src, ins[i].per * (size_t)B * sizeof(double));
        dptr[i] = reinterpret_cast<const double *>(db + in_off[i]);
    }
    CUDA_TRY(cudaMemcpyAsync(db, hb, in_bytes, cudaMemcpyHostToDevice, S.stream));
    dk.M = dptr[0]; dk.J = dptr[1]; dk.dq = dptr[2]; dk.bias = dptr[3];
    dk.ee_xyz = dptr[4]; dk.ee_quat = dptr[5]; dk.target_xyz = dptr[6]; dk.target_quat = dptr[7];
    dk.target_vel = dptr[8]; dk.max_vel = dptr[9]; dk.ft_xmat = dptr[10]; dk.ft_raw = dptr[11];
    dk.ctrl = reinterpret_cast<double *>(db + o_ctrl);
    dk.u_all = hk.u_all ? reinterpret_cast<double *>(db + o_uall) : nullptr;
    dk.status = hk.status ? db + o_stat : nullptr;
    int32_t rc = launch_step(h, B, dk, S.stream);
    if (rc != IRLOSC_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(hb + o_ctrl, db + o_ctrl, total - o_ctrl, cudaMemcpyDeviceToHost, S.stream));
    CUDA_TRY(cudaStreamSynchronize(S.stream));
    memcpy(hk.ctrl, hb + o_ctrl, (size_t)B * P.n_ctrl * sizeof(double));
    if (hk.u_all) memcpy(hk.u_all, hb + o_uall, (size_t)B * P.n * sizeof(double));
    if (hk.status) memcpy(hk.status, hb + o_stat, (size_t)B);
    return IRLOSC_OK;
}

extern "C" int32_t irlosc_step_host(irlosc_handle *h, int64_t B, const irlosc_io *io) {
    if (!h) return fail(IRLOSC_ERR_INVALID, "handle is null");
    if (B < 0) return fail(IRLOSC_ERR_INVALID, "B is negative");
    if (B == 0) return IRLOSC_OK;
    KIo hk;   // host-pointer view with resolved strides
    int32_t rc = resolve_io(h, io, hk, true);
    if (rc != IRLOSC_OK) return rc;
    if (B == 0) return IRLOSC_OK;
    if (hk.n_gather != 0 || hk.ctrl_mc) return fail(IRLOSC_ERR_INVALID, "the fused gather is only available with irlosc_step (device pointers)");
    const KParams &P = h->kp;
    CUDA_TRY(cudaSetDevice(h->device));
    for (int s = 0; s < kPipeDepth; ++s)
        if (!h->stage[s].stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->stage[s].stream, cudaStreamNonBlocking));

    // per-instance element counts of every array, in the slot order used below
    const size_t D = P.D, n = P.n;
    HostIn ins[12] = {
        {hk.M, (size_t)hk.m_stride}, {hk.J, (size_t)hk.j_stride}, {hk.dq, n}, {hk.bias, n},
        {hk.ee_xyz, 3 * D}, {hk.ee_quat, 4 * D}, {hk.target_xyz, 3 * D}, {hk.target_quat, 4 * D},
        {hk.target_vel, 6 * D}, {hk.max_vel, 2 * D}, {hk.ft_xmat, 9 * D}, {hk.ft_raw, 6 * D}};
    if (B <= kSmallBatch) return step_host_small(h, B, hk, ins);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(h->host_chunk, B));
    int turn = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, ++turn) {
        const int64_t nb = std::min<int64_t>(chunk, B - b0);
        Staging &S = h->stage[turn % kPipeDepth];
        const double *dptr[12];
        for (int i = 0; i < 12; ++i) {
            dptr[i] = nullptr;
            if (!ins[i].src) continue;
            const size_t bytes = ins[i].per * (size_t)nb * sizeof(double);
            rc = ensure_cap(S, i, ins[i].per * (size_t)chunk * sizeof(double));
            if (rc != IRLOSC_OK) return rc;
            CUDA_TRY(cudaMemcpyAsync(S.buf[i], ins[i].src + ins[i].per * (size_t)b0, bytes,
                                     cudaMemcpyHostToDevice, S.stream));
            dptr[i] = (const double *)S.buf[i];
        }
        rc = ensure_cap(S, 12, (size_t)chunk * P.n_ctrl * sizeof(double));
        if (rc == IRLOSC_OK && hk.u_all) rc = ensure_cap(S, 13, (size_t)chunk * n * sizeof(double));
        if (rc == IRLOSC_OK && hk.status) rc = ensure_cap(S, 14, (size_t)chunk);
        if (rc != IRLOSC_OK) return rc;
        KIo dk = hk;
        dk.M = dptr[0]; dk.J = dptr[1]; dk.dq = dptr[2]; dk.bias = dptr[3];
        dk.ee_xyz = dptr[4]; dk.ee_quat = dptr[5]; dk.target_xyz = dptr[6]; dk.target_quat = dptr[7];
        dk.target_vel = dptr[8]; dk.max_vel = dptr[9]; dk.ft_xmat = dptr[10]; dk.ft_raw = dptr[11];
        dk.ctrl = (double *)S.buf[12];
        dk.u_all = hk.u_all ? (double *)S.buf[13] : nullptr;
        dk.status = hk.status ? (uint8_t *)S.buf[14] : nullptr;
        rc = launch_step(h, nb, dk, S.stream);
        if (rc != IRLOSC_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(hk.ctrl + (size_t)b0 * P.n_ctrl, dk.ctrl, (size_t)nb * P.n_ctrl * sizeof(double),
                                 cudaMemcpyDeviceToHost, S.stream));
        if (hk.u_all)
            CUDA_TRY(cudaMemcpyAsync(hk.u_all + (size_t)b0 * n, dk.u_all, (size_t)nb * n * sizeof(double),
                                     cudaMemcpyDeviceToHost, S.stream));
        if (hk.status)
            CUDA_TRY(cudaMemcpyAsync(hk.status + b0, dk.status, (size_t)nb, cudaMemcpyDeviceToHost, S.stream));
    }
    for (int s = 0; s < kPipeDepth; ++s) CUDA_TRY(cudaStreamSynchronize(h->stage[s].stream));
    return IRLOSC_OK;
}

// Handle layout and error plumbing shared by the translation units of libirlosc.so.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/irlosc.h"
#include "irlosc_device.cuh"
#include "osc_fused_types.h"

namespace irlosc {

// internal (never returned through the C ABI): the streaming kernel's copy plan of this configuration needs more than
// stream::kMaxChunks chunks (three 6-row devices with admittance) - auto dispatch then takes a record-staging kernel
constexpr int32_t kErrPlanTooLarge = 100;


constexpr int kPipeDepth = 3;   // chunks in flight in the *_host entry points

struct Staging {
    cudaStream_t stream = nullptr;
    void *buf[16] = {nullptr};
    size_t cap[16] = {0};
};

int32_t fail(int32_t rc, const char *fmt, ...);
int32_t ensure_cap(Staging &s, int slot, size_t bytes);

}  // namespace irlosc

struct irlosc_handle;
namespace irlosc {
// streaming step (irlosc_fused.cu, osc_stream.cuh): serves every layout of the DualUR5 topology;
// `stream_preferred` says whether it is also the faster choice (6-row arm devices).
bool stream_supported(const irlosc_handle *h, const KIo &io);
bool stream_preferred(const irlosc_handle *h);
int32_t stream_launch(irlosc_handle *h, int64_t B, const KIo &io, cudaStream_t st);
// tiled step (irlosc_lane.cu, osc_lane.cuh)
void lane_destroy(irlosc_handle *h);
// resolves the input part of an irlosc_io (irlosc.cu)
int32_t resolve_io(const irlosc_handle *h, const irlosc_io *io, KIo &k, bool need_outputs);
}  // namespace irlosc

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return irlosc::fail(IRLOSC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                __FILE__, __LINE__);                                          \
    } while (0)

struct irlosc_handle {
    irlosc_params user;
    irlosc::KParams kp;
    int device = 0;
    int sm_count = 0;
    int kernel_choice = 0;    // 0 auto, 1 generic, 2 + v specialised variant v
    int tile_kernel = 0;      // IRLOSC_TILES_*
    int sm_margin = 0;        // SMs left free for overlapping collectives
    int64_t launches = 0;
    const char *last_kernel = "none";
    irlosc::Staging stage[irlosc::kPipeDepth];
    int64_t host_chunk = 8192;
    // fused state provider (irlosc_set_model)
    bool has_model = false;
    irlosc::fused::KModel km;
    irlosc::fused::FRoles fr;
    int fused_kd = 0;
    bool fused_base = false;
    irlosc::Staging fstage[irlosc::kPipeDepth];
    void *small_host = nullptr, *small_dev = nullptr;   // one-block staging of irlosc_step_host for small batches
    size_t small_cap = 0;
    void *lane_ctx = nullptr;     // tile layout + host pipeline of the tiled step (irlosc_lane.cu), created on first use
};

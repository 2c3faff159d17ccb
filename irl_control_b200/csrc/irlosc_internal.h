// Handle layout and error plumbing shared by the translation units of libirlosc.so.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/irlosc.h"
#include "irlosc_device.cuh"
#include "osc_fused_types.h"

namespace irlosc {

constexpr int kPipeDepth = 3;   // chunks in flight in the *_host entry points

struct Staging {
    cudaStream_t stream = nullptr;
    void *buf[16] = {nullptr};
    size_t cap[16] = {0};
};

int32_t fail(int32_t rc, const char *fmt, ...);
int32_t ensure_cap(Staging &s, int slot, size_t bytes);

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
    int *hard_count = nullptr;        // device counter of queued instances
    int64_t *hard_inst = nullptr;
    double *hard_rec = nullptr;
    int64_t hard_cap = 0;             // instances the queue buffers can hold
    int hard_rec_doubles = 0;
    irlosc::Staging fstage[irlosc::kPipeDepth];
};

// Register-tiled DualUR5 OSC kernel (placeholder until the tiled kernel lands).
#pragma once
#include "irlosc_device.cuh"

namespace irlosc {

inline int32_t tiled_prepare() { return IRLOSC_OK; }
inline bool tiled_supported(const KParams &, const KIo &) { return false; }
inline cudaError_t tiled_launch(const KParams &, const KIo &, int64_t, int, cudaStream_t, const char **) {
    return cudaErrorNotSupported;
}

}  // namespace irlosc

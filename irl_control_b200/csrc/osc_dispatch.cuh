// Kernel table and dispatch for the register-tiled OSC kernels.
#pragma once
#include "osc_tiled.cuh"
#include "osc_rows.cuh"

namespace irlosc {

// ---- host side: instantiations and dispatch --------------------------------------------
struct TiledEntry {
    int n, k, d;
    bool packed;
    int variant;          // 0 = default choice for this shape; others are selectable for experiments
    int ctas_per_sm;
    const void *fn;
    size_t smem_per_cta;
    const char *name;
};

template <int N, int K, int D, int G, bool PACKED, int MINB>
inline TiledEntry tiled_entry(int variant, const char *name) {
    return TiledEntry{N, K, D, PACKED, variant, MINB,
                      (const void *)tiled::osc_step_tiled<N, K, D, G, PACKED, MINB>,
                      sizeof(tiled::WarpSmem<N, K, D, G, PACKED>) * tiled::kWarpsPerCta, name};
}

template <int N, int K, int D, int G, bool PACKED, int MINB>
inline TiledEntry rows_entry(int variant, const char *name) {
    return TiledEntry{N, K, D, PACKED, variant, MINB,
                      (const void *)rows::osc_step_rows<N, K, D, G, PACKED, MINB>,
                      sizeof(rows::RowSmem<N, K, D, G, PACKED>) * tiled::kWarpsPerCta, name};
}

// variant 0 is what irlosc_step dispatches to; the others exist for A/B measurements
// (irlosc_set_kernel(h, 2 + variant)).
inline const TiledEntry *tiled_table(int *count) {
    static const TiledEntry table[] = {
        rows_entry<25, 7, 3, 8, true, 2>(0, "osc_step_rows<n25,k7,D3,G8,packed>"),
        rows_entry<25, 7, 3, 8, false, 2>(0, "osc_step_rows<n25,k7,D3,G8,dense>"),
        tiled_entry<25, 7, 3, 8, true, 2>(1, "osc_step_tiled<n25,k7,D3,G8,packed>"),
        rows_entry<25, 12, 2, 16, true, 2>(1, "osc_step_rows<n25,k12,D2,G16,packed>"),
                tiled_entry<25, 12, 2, 16, true, 2>(0, "osc_step_tiled<n25,k12,D2,G16,packed>"),
        tiled_entry<25, 12, 2, 16, false, 2>(0, "osc_step_tiled<n25,k12,D2,G16,dense>"),
        rows_entry<25, 13, 3, 16, true, 2>(1, "osc_step_rows<n25,k13,D3,G16,packed>"),
                tiled_entry<25, 13, 3, 16, true, 2>(0, "osc_step_tiled<n25,k13,D3,G16,packed>"),
        tiled_entry<25, 13, 3, 16, false, 2>(0, "osc_step_tiled<n25,k13,D3,G16,dense>"),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

inline int32_t tiled_prepare() {
    int cnt = 0;
    const TiledEntry *t = tiled_table(&cnt);
    for (int i = 0; i < cnt; ++i) {
        if (cudaFuncSetAttribute(t[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t[i].smem_per_cta) != cudaSuccess)
            return IRLOSC_ERR_CUDA;
    }
    return IRLOSC_OK;
}

inline const TiledEntry *tiled_find(const KParams &P, const KIo &io, int variant = 0) {
    auto al16 = [](const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (io.j_layout != IRLOSC_J_ROWS || io.ldj != P.n || io.j_stride != (int64_t)P.n * P.k) return nullptr;
    const bool packed = io.m_layout == IRLOSC_M_PACKED;
    if (packed && io.m_stride != (int64_t)P.n * (P.n + 1) / 2) return nullptr;
    if (!packed && (io.ldm != P.n || io.m_stride != (int64_t)P.n * P.n)) return nullptr;
    if (!(al16(io.M) && al16(io.J) && al16(io.dq) && al16(io.bias) && al16(io.ee_xyz) && al16(io.ee_quat) &&
          al16(io.target_xyz) && al16(io.target_quat) && al16(io.target_vel) && al16(io.max_vel) &&
          al16(io.ft_xmat) && al16(io.ft_raw)))
        return nullptr;
    int cnt = 0;
    const TiledEntry *t = tiled_table(&cnt);
    for (int i = 0; i < cnt; ++i)
        if (t[i].n == P.n && t[i].k == P.k && t[i].d == P.D && t[i].packed == packed && t[i].variant == variant) return &t[i];
    return nullptr;
}

inline bool tiled_supported(const KParams &P, const KIo &io, int variant = 0) { return tiled_find(P, io, variant) != nullptr; }

inline cudaError_t tiled_launch(const KParams &P, const KIo &io, int64_t B, int sm_count, cudaStream_t st,
                                const char **name, int variant = 0) {
    const TiledEntry *e = tiled_find(P, io, variant);
    if (!e) return cudaErrorNotSupported;
    const int grid = sm_count * e->ctas_per_sm;
    void *args[] = {(void *)&P, (void *)&io, (void *)&B};
    *name = e->name;
    return cudaLaunchKernel(e->fn, dim3(grid), dim3(tiled::kWarpsPerCta * 32), args, e->smem_per_cta, st);
}

}  // namespace irlosc

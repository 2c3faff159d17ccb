// Kernel table and dispatch of the record-staging DualUR5 kernel (osc_tree.cuh: kinematic-tree-sparse, 4 lanes per
// instance, TMA-staged records) - needs irlosc_params.has_topology and the DualUR5 tree.  Without a topology the
// generic kernel (osc_generic.cuh) is the dense fallback.  Variant 0 of a shape is what irlosc_step dispatches to;
// the others are kept for A/B measurements (irlosc_set_kernel(h, 2 + variant)).
#pragma once
#include <cstdlib>
#include "osc_tree.cuh"

namespace irlosc {

struct TiledEntry {
    int kind;             // 0 dense, 1 tree
    int n, k, d;          // shape served (tree: k and d follow from kd / has_base)
    bool packed;
    int variant;
    int ctas_per_sm;
    const void *fn;
    size_t smem_per_cta;
    const char *name;
    int kd;               // tree only
    bool has_base;        // tree only
    int warps, slots;     // tree only: warps per CTA, input slots per CTA
    int slot_fixed, priv; // tree only: bytes of the fixed slot part / per-warp scratch
    bool qm = false;      // tree only: M arrives as MuJoCo's sparse qM (IRLOSC_M_QM, tight stride 155)
};

constexpr size_t kTreeSmemLimit = 227 * 1024;
constexpr int kTreeHeader = 64;

template <int KD, bool HAS_BASE, bool PACKED, int W, int NS, int GW = 1, bool QM = false>
inline TiledEntry tree_entry(int variant, const char *name) {
    using SLT = tree::TreeSlot<KD, HAS_BASE, PACKED, QM>;
    using PVT = tree::TreePriv<KD, HAS_BASE>;
    return TiledEntry{1, tree::kN, SLT::K, SLT::D, PACKED, variant, 1,
                      (const void *)tree::osc_step_tree<KD, HAS_BASE, PACKED, W, NS, GW, QM>,
                      kTreeSmemLimit, name, KD, HAS_BASE, W, NS,
                      (int)(sizeof(SLT) - 16), (int)((sizeof(PVT) + 15) & ~size_t(15)), QM};
}

inline const TiledEntry *tiled_table(int *count) {
    static const TiledEntry table[] = {
        // ---- kinematic-tree-sparse (default when the topology is declared); for one variant number
        //      the first entry whose shared-memory plan fits is taken
        tree_entry<3, true, true, 7, 3, 1>(0, "osc_step_tree<kd3,base,packed,w7,s3,g1>"),
        tree_entry<3, true, true, 6, 2, 1>(0, "osc_step_tree<kd3,base,packed,w6,s2,g1>"),
        tree_entry<3, true, false, 6, 2, 1>(0, "osc_step_tree<kd3,base,dense,w6,s2,g1>"),
        tree_entry<6, false, true, 4, 2, 1>(0, "osc_step_tree<kd6,packed,w4,s2,g1>"),
        tree_entry<6, false, false, 3, 2, 1>(0, "osc_step_tree<kd6,dense,w3,s2,g1>"),
        tree_entry<6, true, true, 4, 2, 1>(0, "osc_step_tree<kd6,base,packed,w4,s2,g1>"),
        tree_entry<6, true, false, 3, 2, 1>(0, "osc_step_tree<kd6,base,dense,w3,s2,g1>"),
        // ---- tree-sparse on MuJoCo's sparse qM (IRLOSC_M_QM): a slot is 27 KB instead of 38 KB, so a fourth slot and
        //      an eighth warp fit (round-1 driver run: 0.112 ms vs 0.143 ms packed at B = 65 536)
        tree_entry<3, true, true, 8, 4, 1, true>(0, "osc_step_tree<kd3,base,qM,w8,s4,g1>"),
        tree_entry<3, true, true, 7, 3, 1, true>(0, "osc_step_tree<kd3,base,qM,w7,s3,g1>"),
        tree_entry<6, false, true, 4, 2, 1, true>(0, "osc_step_tree<kd6,qM,w4,s2,g1>"),
        tree_entry<6, true, true, 4, 2, 1, true>(0, "osc_step_tree<kd6,base,qM,w4,s2,g1>"),
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

// Does the controller description match what osc_tree.cuh is specialised for?
inline bool tree_roles(const KParams &P, tree::Roles &R, int &kd, bool &has_base) {
    R.dev_arm[0] = R.dev_arm[1] = R.dev_base = -1;
    R.row_arm[0] = R.row_arm[1] = R.row_base = 0;
    if (!P.has_topology || P.n != tree::kN || P.D < 2 || P.D > 3) return false;
    for (int j = 0; j < tree::kN; ++j)
        if (P.joint_parent[j] != tree::kDualUr5Parent[j]) return false;
    for (int d = 0; d < P.D; ++d) {
        const KDevice &dv = P.dev[d];
        if (dv.ee_joint == 6 && R.dev_arm[0] < 0) { R.dev_arm[0] = d; R.row_arm[0] = dv.row0; }
        else if (dv.ee_joint == 18 && R.dev_arm[1] < 0) { R.dev_arm[1] = d; R.row_arm[1] = dv.row0; }
        else if (dv.ee_joint == 0 && R.dev_base < 0) { R.dev_base = d; R.row_base = dv.row0; }
        else return false;
    }
    R.opt_tvel = R.opt_mvel = R.opt_ftx = R.opt_ftr = -1;
    R.opt_doubles = R.slot_bytes = R.priv_bytes = R.ctrl_vec = 0;
    if (R.dev_arm[0] < 0 || R.dev_arm[1] < 0) return false;
    kd = P.dev[R.dev_arm[0]].kdev;
    if (P.dev[R.dev_arm[1]].kdev != kd || (kd != 3 && kd != 6)) return false;
    has_base = R.dev_base >= 0;
    if (has_base && P.dev[R.dev_base].kdev != 1) return false;
    if ((P.D == 3) != has_base) return false;
    return true;
}

// Dispatch order for variant v: with a matching topology the tree kernels are variant 0 and the
// dense ones follow (v >= 1); without topology the dense variants shift down by one.
inline const TiledEntry *tiled_find(const KParams &P, const KIo &io, int variant, tree::Roles *roles_out) {
    auto al16 = [](const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (io.j_layout != IRLOSC_J_ROWS || io.ldj != P.n || io.j_stride != (int64_t)P.n * P.k) return nullptr;
    const bool qm = io.m_layout == IRLOSC_M_QM;
    const bool packed = io.m_layout == IRLOSC_M_PACKED || qm;       // qM entries are registered with packed = true
    if (qm && io.m_stride != tree::kQmSize) return nullptr;        // a tile's qM block must be one contiguous bulk copy
    if (!qm && packed && io.m_stride != (int64_t)P.n * (P.n + 1) / 2) return nullptr;
    if (!packed && (io.ldm != P.n || io.m_stride != (int64_t)P.n * P.n)) return nullptr;
    if (!(al16(io.M) && al16(io.J) && al16(io.dq) && al16(io.bias) && al16(io.ee_xyz) && al16(io.ee_quat) &&
          al16(io.target_xyz) && al16(io.target_quat) && al16(io.target_vel) && al16(io.max_vel) &&
          al16(io.ft_xmat) && al16(io.ft_raw)))
        return nullptr;
    tree::Roles R;
    int kd = 0;
    bool has_base = false;
    const bool tree_ok = tree_roles(P, R, kd, has_base);
    if (roles_out) *roles_out = R;
    if (!tree_ok) return nullptr;
    const int want = variant;
    int cnt = 0;
    const TiledEntry *t = tiled_table(&cnt);
    for (int i = 0; i < cnt; ++i) {
        if (t[i].packed != packed || t[i].qm != qm || t[i].variant != want) continue;
        if (t[i].kind == 1) {
            if (!(tree_ok && t[i].kd == kd && t[i].has_base == has_base)) continue;
            // shared-memory plan: optional per-instance arrays go to the slot tail
            int at = 0;
            const int per = tree::kWI * P.D;
            if (io.target_vel) { R.opt_tvel = at; at += 6 * per; }
            if (io.max_vel) { R.opt_mvel = at; at += 2 * per; }
            if (P.admittance) { R.opt_ftx = at; at += 9 * per; R.opt_ftr = at; at += 6 * per; }
            R.opt_doubles = at;
            R.ctrl_vec = al16(io.ctrl) && ((io.gather_offset * (int64_t)P.n_ctrl) % 2 == 0);
            for (int g = 0; g < io.n_gather; ++g) R.ctrl_vec = R.ctrl_vec && al16(io.ctrl_gather[g]);
            R.ctrl_vec = R.ctrl_vec && al16(io.ctrl_mc);
            R.slot_bytes = t[i].slot_fixed + ((at * 8 + 15) & ~15);
            R.priv_bytes = t[i].priv - 16 + ((at * 8 + 15) & ~15);
            const size_t total = kTreeHeader + (size_t)t[i].slots * R.slot_bytes + (size_t)t[i].warps * R.priv_bytes;
            if (total > kTreeSmemLimit) continue;
            if (roles_out) *roles_out = R;
            return &t[i];
        } else if (t[i].n == P.n && t[i].k == P.k && t[i].d == P.D) {
            return &t[i];
        }
    }
    return nullptr;
}

inline bool tiled_supported(const KParams &P, const KIo &io, int variant = 0) {
    return tiled_find(P, io, variant, nullptr) != nullptr;
}

inline cudaError_t tiled_launch(const KParams &P, const KIo &io, int64_t B, int sm_count, cudaStream_t st,
                                const char **name, int variant = 0) {
    tree::Roles R;
    const TiledEntry *e = tiled_find(P, io, variant, &R);
    if (!e) return cudaErrorNotSupported;
    int grid = sm_count * e->ctas_per_sm;
    if (const char *g = getenv("IRLOSC_GRID")) grid = atoi(g) > 0 ? atoi(g) : grid;   // experiments only
    *name = e->name;
    if (e->kind == 1) {
        const size_t smem = kTreeHeader + (size_t)e->slots * R.slot_bytes + (size_t)e->warps * R.priv_bytes;
        void *args[] = {(void *)&P, (void *)&io, (void *)&B, (void *)&R};
        return cudaLaunchKernel(e->fn, dim3(grid), dim3(e->warps * 32), args, smem, st);
    }
    return cudaErrorNotSupported;
}

}  // namespace irlosc

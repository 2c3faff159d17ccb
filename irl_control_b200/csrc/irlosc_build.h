// Host-side flattening of the C-ABI parameter blocks into what the kernels read:
// irlosc_params -> KParams (index maps, gains; device.py:41-74, robot.py:26-32, osc.py:26-39) and
// irlosc_model -> fused::KModel.  Shared by the translation units of libirlosc.so and by the
// host-compiled test harness (tests/host_fused).
#pragma once
#include <cmath>
#include <cstring>
#include "irlosc_internal.h"
#include "osc_stream.cuh"
#include <vector>
#include <algorithm>

namespace irlosc {

inline int32_t build_kparams(const irlosc_params &u, KParams &kp) {
    memset(&kp, 0, sizeof kp);
    if (u.abi_version != IRLOSC_ABI_VERSION)
        return fail(IRLOSC_ERR_INVALID, "abi_version %d, library is %d", u.abi_version, IRLOSC_ABI_VERSION);
    if (u.n < 1 || u.n > IRLOSC_MAX_N) return fail(IRLOSC_ERR_INVALID, "n=%d outside 1..%d", u.n, IRLOSC_MAX_N);
    if (u.n_devices < 1 || u.n_devices > IRLOSC_MAX_DEVICES)
        return fail(IRLOSC_ERR_INVALID, "n_devices=%d outside 1..%d", u.n_devices, IRLOSC_MAX_DEVICES);
    kp.n = u.n;
    kp.D = u.n_devices;
    kp.use_g = u.use_g != 0;
    kp.admittance = u.admittance != 0;
    kp.has_nullspace = u.has_nullspace != 0;
    kp.nullspace_kv = u.nullspace_kv;
    kp.has_topology = u.has_topology != 0;
    kp.check_topology = u.check_topology != 0;
    for (int j = 0; j < IRLOSC_MAX_N; ++j) {
        const int pj = (kp.has_topology && j < u.n) ? u.joint_parent[j] : -1;
        if (pj < -1 || pj >= u.n) return fail(IRLOSC_ERR_INVALID, "joint_parent[%d]=%d outside -1..%d", j, pj, u.n - 1);
        kp.joint_parent[j] = (int8_t)pj;
    }
    int row = 0, ctrl = 0;
    for (int d = 0; d < u.n_devices; ++d) {
        const irlosc_device_params &s = u.dev[d];
        KDevice &t = kp.dev[d];
        t.row0 = row;
        t.ctrl0 = ctrl;
        int kd = 0;
        for (int i = 0; i < 6; ++i) {
            t.dof[i] = s.ctrlr_dof[i] != 0;
            if (t.dof[i]) {
                if (row >= IRLOSC_MAX_K) return fail(IRLOSC_ERR_INVALID, "more than %d task rows", IRLOSC_MAX_K);
                kp.row_dev[row] = (int8_t)d;
                kp.row_comp[row] = (int8_t)i;
                t.dx_idx[kd] = s.dx_idx[kd];
                if (s.dx_idx[kd] < 0) return fail(IRLOSC_ERR_INVALID, "device %d: negative dx_idx", d);
                ++row;
                ++kd;
            }
        }
        t.kdev = kd;
        t.any_xyz = (t.dof[0] + t.dof[1] + t.dof[2]) > 0;
        t.any_abg = (t.dof[3] + t.dof[4] + t.dof[5]) > 0;
        t.has_max_vel = s.has_max_vel != 0;
        t.max_vel[0] = s.max_vel[0];
        t.max_vel[1] = s.max_vel[1];
        t.kp = s.kp; t.kv = s.kv; t.ko = s.ko;
        t.kv_over_kp = s.kp != 0.0 ? s.kv / s.kp : 0.0;
        t.kv_over_ko = s.ko != 0.0 ? s.kv / s.ko : 0.0;
        if (!(s.kv != 0.0)) return fail(IRLOSC_ERR_INVALID, "device %d: kv must be non-zero", d);
        for (int i = 0; i < 6; ++i) {
            t.gain[i] = s.has_gain_vectors ? s.task_space_gains[i] : ((i < 3) ? s.kp : s.ko);
            t.lamb[i] = s.has_gain_vectors ? s.lamb[i] : t.gain[i] / s.kv;
            t.stiff[i] = (i < 3) ? s.k[i] : 1.0;
            t.damp[i] = (i < 3) ? s.d[i] : 1.0;
        }
        if (s.n_joints_all < 0 || s.n_joints_all > u.n)
            return fail(IRLOSC_ERR_INVALID, "device %d: n_joints_all=%d", d, s.n_joints_all);
        t.n_joints_all = s.n_joints_all;
        for (int i = 0; i < s.n_joints_all; ++i) {
            const int j = s.joint_ids_all[i];
            if (j < 0 || j >= u.n) return fail(IRLOSC_ERR_INVALID, "device %d: joint id %d outside 0..%d", d, j, u.n - 1);
            t.joint_mask |= (1u << j);
        }
        if (s.n_ctrl < 0 || s.n_ctrl > u.n) return fail(IRLOSC_ERR_INVALID, "device %d: n_ctrl=%d", d, s.n_ctrl);
        t.n_ctrl = s.n_ctrl;
        for (int i = 0; i < s.n_ctrl; ++i) {
            const int j = s.actuator_trnids[i];
            if (j < 0 || j >= u.n) return fail(IRLOSC_ERR_INVALID, "device %d: actuator joint %d outside 0..%d", d, j, u.n - 1);
            t.actuator[i] = (int8_t)j;
        }
        t.ee_joint = (s.ee_joint >= 0 && s.ee_joint < u.n) ? s.ee_joint : -1;
        ctrl += s.n_ctrl;
    }
    if (row < 1) return fail(IRLOSC_ERR_INVALID, "no controlled task rows");
    if (ctrl < 1 || ctrl > 32) return fail(IRLOSC_ERR_INVALID, "n_ctrl=%d outside 1..32", ctrl);
    kp.k = row;
    kp.n_ctrl = ctrl;
    return IRLOSC_OK;
}


namespace fused_build {
using fused::KJoint;
using fused::KFrame;
using fused::FRoles;
using fused::kN;
constexpr int kDualUr5Parent[kN] = {-1, 0, 1, 2, 3, 4, 5, 6, 7, 6, 6, 10, 6, 0, 13, 14, 15, 16, 17, 18, 19, 18, 18, 22, 18};

inline void quat_to_mat(const double *q, double *R) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double s = 2.0 / (w * w + x * x + y * y + z * z);
    const double xs = x * s, ys = y * s, zs = z * s;
    R[0] = 1.0 - (y * ys + z * zs); R[1] = x * ys - w * zs;         R[2] = x * zs + w * ys;
    R[3] = x * ys + w * zs;         R[4] = 1.0 - (x * xs + z * zs); R[5] = y * zs - w * xs;
    R[6] = x * zs - w * ys;         R[7] = y * zs + w * xs;         R[8] = 1.0 - (x * xs + y * ys);
}

inline int32_t build_joint(const irlosc_joint_model &m, int j, KJoint &k) {
    const double an = std::sqrt(m.axis[0] * m.axis[0] + m.axis[1] * m.axis[1] + m.axis[2] * m.axis[2]);
    if (!(an > 0.0)) return fail(IRLOSC_ERR_INVALID, "joint %d: zero hinge axis", j);
    if (!(m.mass >= 0.0)) return fail(IRLOSC_ERR_INVALID, "joint %d: negative mass", j);
    const double a[3] = {m.axis[0] / an, m.axis[1] / an, m.axis[2] / an};
    double Rf[9];
    quat_to_mat(m.quat, Rf);
    const double ax[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    double Ra[3];
    for (int i = 0; i < 3; ++i) Ra[i] = Rf[3 * i] * a[0] + Rf[3 * i + 1] * a[1] + Rf[3 * i + 2] * a[2];
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) {
            double p1 = 0.0;
            for (int t = 0; t < 3; ++t) p1 += Rf[3 * i + t] * ax[3 * t + c];
            k.P1[3 * i + c] = p1;
            k.P2[3 * i + c] = Ra[i] * a[c];
            k.Q0[3 * i + c] = Rf[3 * i + c] - k.P2[3 * i + c];
        }
    for (int i = 0; i < 3; ++i) { k.pos[i] = m.pos[i]; k.axp[i] = Ra[i]; k.com[i] = m.com[i]; }
    k.mass = m.mass;
    for (int i = 0; i < 6; ++i) k.ic[i] = m.inertia[i];
    return IRLOSC_OK;
}

inline int32_t build_frame(const irlosc_frame_model &f, int want_joint, const char *what, int d, KFrame &k) {
    memset(&k, 0, sizeof k);
    if (f.joint < 0) { k.joint = -1; k.has = 0; k.R[0] = k.R[4] = k.R[8] = 1.0; return IRLOSC_OK; }
    if (f.joint != want_joint)
        return fail(IRLOSC_ERR_INVALID, "%s frame of device %d hangs off joint %d, its Jacobian ends at joint %d", what, d,
                    f.joint, want_joint);
    k.joint = f.joint;
    k.has = 1;
    for (int i = 0; i < 3; ++i) k.pos[i] = f.pos[i];
    quat_to_mat(f.quat, k.R);
    return IRLOSC_OK;
}

// which devices play which role; same contract as tree_roles in osc_dispatch.cuh
inline bool fused_roles(const KParams &P, FRoles &R, int &kd, bool &has_base) {
    R.dev_arm[0] = R.dev_arm[1] = R.dev_base = -1;
    R.row_arm[0] = R.row_arm[1] = R.row_base = 0;
    if (!P.has_topology || P.n != kN || P.D < 2 || P.D > 3) return false;
    for (int j = 0; j < kN; ++j)
        if (P.joint_parent[j] != kDualUr5Parent[j]) return false;
    for (int d = 0; d < P.D; ++d) {
        const KDevice &dv = P.dev[d];
        if (dv.ee_joint == 6 && R.dev_arm[0] < 0) { R.dev_arm[0] = d; R.row_arm[0] = dv.row0; }
        else if (dv.ee_joint == 18 && R.dev_arm[1] < 0) { R.dev_arm[1] = d; R.row_arm[1] = dv.row0; }
        else if (dv.ee_joint == 0 && R.dev_base < 0) { R.dev_base = d; R.row_base = dv.row0; }
        else return false;
    }
    if (R.dev_arm[0] < 0 || R.dev_arm[1] < 0) return false;
    kd = P.dev[R.dev_arm[0]].kdev;
    if (P.dev[R.dev_arm[1]].kdev != kd || (kd != 3 && kd != 6)) return false;
    has_base = R.dev_base >= 0;
    if (has_base && P.dev[R.dev_base].kdev != 1) return false;
    if ((P.D == 3) != has_base) return false;
    // joint -> packed ctrl slot; a joint returned twice cannot be served by the single-store packing
    for (int j = 0; j < IRLOSC_MAX_N; ++j) R.joint_slot[j] = -1;
    for (int d = 0; d < P.D; ++d)
        for (int c = 0; c < P.dev[d].n_ctrl; ++c) {
            const int j = P.dev[d].actuator[c];
            if (R.joint_slot[j] >= 0) return false;
            R.joint_slot[j] = (int8_t)(P.dev[d].ctrl0 + c);
        }
    auto owners = [&](int j) { unsigned m = 0; for (int d = 0; d < P.D; ++d) m |= ((P.dev[d].joint_mask >> j) & 1u) << d; return m; };
    R.uniform_owner = 1;
    for (int arm = 0; arm < 2; ++arm) {
        const int jb = 1 + 12 * arm;
        for (int i = 1; i < 6; ++i) if (owners(jb + i) != owners(jb)) R.uniform_owner = 0;
        for (int half = 0; half < 2; ++half)
            for (int r = 1; r < 3; ++r) if (owners(jb + 6 + 3 * half + r) != owners(jb + 6 + 3 * half)) R.uniform_owner = 0;
    }
    return true;
}

}  // namespace fused_build
using fused_build::fused_roles;
using fused_build::build_joint;
using fused_build::build_frame;

// irlosc_model -> kernel constants (frames as matrices, Rodrigues terms folded with the fixed frames).
inline int32_t build_kmodel(const KParams &P, const irlosc_model &m, fused::KModel &K) {
    using namespace fused;
    memset(&K, 0, sizeof K);
    for (int i = 0; i < 3; ++i) K.gravity[i] = m.gravity[i];
    int32_t rc = build_joint(m.joint[0], 0, K.stand);
    for (int a = 0; a < 2 && rc == IRLOSC_OK; ++a) {
        const int jb = 1 + 12 * a;
        for (int i = 0; i < 6 && rc == IRLOSC_OK; ++i) rc = build_joint(m.joint[jb + i], jb + i, K.arm[a][i]);
        for (int hf = 0; hf < 2 && rc == IRLOSC_OK; ++hf)
            for (int r = 0; r < 3 && rc == IRLOSC_OK; ++r)
                rc = build_joint(m.joint[jb + 6 + 3 * hf + r], jb + 6 + 3 * hf + r, K.grip[a][hf][r]);
    }
    if (rc != IRLOSC_OK) return rc;
    for (int d = 0; d < P.D; ++d) {
        const int want = P.dev[d].ee_joint;
        if (m.ee[d].joint < 0) return fail(IRLOSC_ERR_INVALID, "device %d has no EE frame", d);
        rc = build_frame(m.ee[d], want, "EE", d, K.ee[d]);
        if (rc == IRLOSC_OK) rc = build_frame(m.ft[d], want, "F/T", d, K.ft[d]);
        if (rc != IRLOSC_OK) return rc;
    }
    return IRLOSC_OK;
}

// irlosc_sequence -> kernel constants, validated against the controller description.
inline int32_t build_kseq(const KParams &P, const fused::FRoles &R, const irlosc_sequence &u, fused::KSeq &Q) {
    memset(&Q, 0, sizeof Q);
    if (u.n_actions < 1 || u.n_actions > IRLOSC_MAX_ACTIONS)
        return fail(IRLOSC_ERR_INVALID, "n_actions=%d outside 1..%d", u.n_actions, IRLOSC_MAX_ACTIONS);
    if (u.active_device != R.dev_arm[0] && u.active_device != R.dev_arm[1])
        return fail(IRLOSC_ERR_INVALID, "active_device=%d is not one of the arm devices", u.active_device);
    if (u.gripper_slot < -1 || u.gripper_slot >= P.n_ctrl) return fail(IRLOSC_ERR_INVALID, "gripper_slot=%d outside -1..%d", u.gripper_slot, P.n_ctrl - 1);
    Q.n_actions = u.n_actions;
    Q.active_dev = u.active_device;
    Q.gripper_slot = u.gripper_slot;
    for (int i = 0; i < 4; ++i) Q.passive_quat[i] = u.passive_quat[i];
    for (int a = 0; a < u.n_actions; ++a) {
        const irlosc_action &s = u.action[a];
        if (s.type != IRLOSC_ACT_WP && s.type != IRLOSC_ACT_GRIP) return fail(IRLOSC_ERR_INVALID, "action %d: unknown type %d", a, s.type);
        if (s.type == IRLOSC_ACT_GRIP && s.grip_steps < 0) return fail(IRLOSC_ERR_INVALID, "action %d: negative grip_steps", a);
        Q.act[a].type = s.type;
        Q.act[a].grip_steps = s.grip_steps;
        Q.act[a].kp = s.kp; Q.act[a].max_error = s.max_error;
        Q.act[a].min_speed = s.min_speed_xyz; Q.act[a].max_speed = s.max_speed_xyz;
        Q.act[a].gripper_force = s.gripper_force;
    }
    return IRLOSC_OK;
}

inline int32_t build_kseq_waypoints(const KParams &P, const fused::FRoles &R, const irlosc_waypoints_io &u, fused::KSeq &Q) {
    memset(&Q, 0, sizeof Q);
    if (!u.wps || !u.wp_idx || !u.target_xyz || !u.target_quat) return fail(IRLOSC_ERR_INVALID, "every array of irlosc_waypoints_io is required");
    if (u.W < 1) return fail(IRLOSC_ERR_INVALID, "W=%d", u.W);
    Q.mode = 1;
    Q.active_dev = R.dev_arm[0];
    Q.gripper_slot = -1;
    Q.W = u.W;
    Q.threshold = u.threshold;
    for (int a = 0; a < 2; ++a) {
        const int d = R.dev_arm[a];
        if (u.n_wp[d] < 1 || u.n_wp[d] > u.W) return fail(IRLOSC_ERR_INVALID, "n_wp[%d]=%d outside 1..W=%d", d, u.n_wp[d], u.W);
    }
    for (int d = 0; d < IRLOSC_MAX_DEVICES; ++d) Q.n_wp[d] = u.n_wp[d];
    (void)P;
    return IRLOSC_OK;
}

// MuJoCo's sparse inertia `qM` (mjData.qM, the array robot.py:69 hands to mj_fullM; IRLOSC_M_QM): dof i owns
// the entries M[i][i], M[i][parent(i)], M[i][parent(parent(i))], ... down to its root, stored from
// dof_Madr[i], and dof_Madr[i + 1] = dof_Madr[i] + depth(i) (mj_fullM walks exactly this: `adr = dof_Madr[i];
// for (j = i; j >= 0; j = dof_parentid[j]) dst[i][j] = dst[j][i] = qM[adr++]`).  Callers have checked that
// every parent precedes its child.
inline int qm_depth(const KParams &P, int i) {
    int d = 0;
    for (int j = i; j >= 0; j = P.joint_parent[j]) ++d;
    return d;
}
inline int qm_size(const KParams &P) {
    int s = 0;
    for (int i = 0; i < P.n; ++i) s += qm_depth(P, i);
    return s;
}
inline int qm_offset(const KParams &P, int i, int j) {       // -1 unless j is i or one of its ancestors
    int adr = 0;
    for (int r = 0; r < i; ++r) adr += qm_depth(P, r);
    for (int a = i; a >= 0; a = P.joint_parent[a], ++adr)
        if (a == j) return adr;
    return -1;
}

// M part of irlosc_io -> KIo: layout, leading dimension and instance stride with their defaults.
inline int32_t resolve_m_layout(const KParams &P, const irlosc_io &io, KIo &k) {
    k.M = io.M; k.m_layout = io.m_layout;
    if (io.m_layout == IRLOSC_M_DENSE) {
        k.ldm = io.ldm ? io.ldm : P.n;
        if (k.ldm < P.n) return fail(IRLOSC_ERR_INVALID, "ldm=%d < n=%d", k.ldm, P.n);
        k.m_stride = io.m_stride ? io.m_stride : (int64_t)k.ldm * P.n;
    } else if (io.m_layout == IRLOSC_M_PACKED) {
        k.ldm = 0;
        k.m_stride = io.m_stride ? io.m_stride : (int64_t)P.n * (P.n + 1) / 2;
    } else if (io.m_layout == IRLOSC_M_QM) {
        if (!P.has_topology) return fail(IRLOSC_ERR_INVALID, "m_layout IRLOSC_M_QM needs has_topology (joint_parent = dof_parentid)");
        for (int j = 0; j < P.n; ++j)
            if (P.joint_parent[j] >= j) return fail(IRLOSC_ERR_INVALID, "IRLOSC_M_QM: joint_parent[%d]=%d does not precede it", j, P.joint_parent[j]);
        const int nm = qm_size(P);
        k.ldm = 0;
        k.m_stride = io.m_stride ? io.m_stride : (int64_t)nm;
        if (k.m_stride < nm) return fail(IRLOSC_ERR_INVALID, "IRLOSC_M_QM: m_stride=%lld < nM=%d", (long long)k.m_stride, nm);
    } else {
        return fail(IRLOSC_ERR_INVALID, "unknown m_layout %d", io.m_layout);
    }
    return IRLOSC_OK;
}

// Copy plan of the streaming step (osc_stream.cuh): which doubles of an instance's arrays go to which
// stage entry of which group, as 8-entry chunks with per-lane byte offsets.  Works for every M / J
// layout and stride of irlosc_io because a chunk carries its array's base pointer and stride.
inline int32_t build_stream_plan(const KParams &P, const KIo &io, const fused::FRoles &R, int kd, bool has_base,
                                 stream::Plan &plan) {
    using namespace stream;
    struct Item { const void *base; int64_t stride; int64_t off; int dst; };
    std::vector<Item> groups[kGroups];
    const int n = P.n, D = P.D;
    bool qm_miss = false;
    auto M_at = [&](int i, int j) -> Item {        // i >= j, and j is i or an ancestor of i (the tree's non-zeros)
        int64_t e = io.m_layout == IRLOSC_M_PACKED ? (int64_t)i * (i + 1) / 2 + j : (int64_t)i * io.ldm + j;
        if (io.m_layout == IRLOSC_M_QM) {
            e = qm_offset(P, i, j);
            if (e < 0) { qm_miss = true; e = 0; }
        }
        return Item{io.M, io.m_stride * 8, e * 8, 0};
    };
    auto J_at = [&](int row, int j) -> Item {
        const int64_t r = io.j_layout == IRLOSC_J_ROWS ? row : (int64_t)P.row_dev[row] * 6 + P.row_comp[row];
        return Item{io.J, io.j_stride * 8, (r * io.ldj + j) * 8, 0};
    };
    auto vec = [&](const double *base, int per, int idx) -> Item { return Item{base, (int64_t)per * 8, (int64_t)idx * 8, 0}; };
    auto put = [&](int g, Item it, int dst) { it.dst = dst; groups[g].push_back(it); };
    const double *bias = io.bias ? io.bias : io.dq;      // use_g == 0: any finite value (multiplied by 0)
    auto device_block = [&](int g, int d, int e0) {
        for (int i = 0; i < 3; ++i) { put(g, vec(io.ee_xyz, 3 * D, d * 3 + i), e0 + kEeXyz + i); put(g, vec(io.target_xyz, 3 * D, d * 3 + i), e0 + kTXyz + i); }
        for (int i = 0; i < 4; ++i) { put(g, vec(io.ee_quat, 4 * D, d * 4 + i), e0 + kEeQuat + i); put(g, vec(io.target_quat, 4 * D, d * 4 + i), e0 + kTQuat + i); }
        if (io.max_vel)
            for (int i = 0; i < 2; ++i) put(g, vec(io.max_vel, 2 * D, d * 2 + i), e0 + kMaxVel + i);
        if (P.admittance) {
            for (int i = 0; i < 9; ++i) put(g, vec(io.ft_xmat, 9 * D, d * 9 + i), e0 + kFtX + i);
            for (int i = 0; i < 6; ++i) put(g, vec(io.ft_raw, 6 * D, d * 6 + i), e0 + kFtRaw + i);
        }
    };
    int max_entries = kG1Entries;
    // G0
    put(0, vec(bias, n, 0), kG0Bias0);
    if (has_base) {
        put(0, J_at(R.row_base, 0), kG0Jbase);
        device_block(0, R.dev_base, kG0Dev);
        max_entries = std::max(max_entries, kG0Dev + kDevEntries);
    }
    for (int arm = 0; arm < 2; ++arm) {
        const int jb = 1 + 12 * arm, g0 = 1 + 5 * arm;
        auto C = [&](int i) { return i == 0 ? 0 : jb + i - 1; };
        for (int i = 0; i < 7; ++i) {
            for (int j = 0; j <= i; ++j) put(g0, M_at(C(i), C(j)), kCC + i * (i + 1) / 2 + j);
            put(g0, vec(io.dq, n, C(i)), kDqC + i);
        }
        for (int half = 0; half < 2; ++half) {
            const int g = g0 + 1 + half, gj = jb + 6 + 3 * half;
            for (int i = 0; i < 7; ++i) {
                put(g, M_at(gj + 1, C(i)), kRg1 + i);
                put(g, M_at(gj, C(i)), kRg0 + i);
                put(g, M_at(gj + 2, C(i)), kRg2 + i);
            }
            put(g, M_at(gj + 1, gj), kE10);
            put(g, M_at(gj + 1, gj + 1), kD1);
            put(g, M_at(gj, gj), kD0);
            put(g, M_at(gj + 2, gj + 2), kD2);
            for (int r = 0; r < 3; ++r) { put(g, vec(io.dq, n, gj + r), kDqG + r); put(g, vec(bias, n, gj + r), kBiasG + r); }
        }
        for (int cr = 0; cr < kd; ++cr)
            for (int i = 0; i < 7; ++i) put(g0 + 3, J_at(R.row_arm[arm] + cr, C(i)), cr * 7 + i);
        for (int i = 0; i < 6; ++i) put(g0 + 3, vec(bias, n, jb + i), kd * 7 + i);
        max_entries = std::max(max_entries, kd * 7 + 6);
        device_block(g0 + 4, R.dev_arm[arm], 0);
    }
    if (qm_miss) return fail(IRLOSC_ERR_INVALID, "IRLOSC_M_QM: the plan asked for an entry outside the kinematic tree");
    memset(&plan, 0, sizeof plan);
    plan.has_mvel = io.max_vel != nullptr;
    const int trash = max_entries;
    plan.stage_entries = max_entries + 1;
    int nc = 0;
    for (int g = 0; g < kGroups; ++g) {
        plan.first[g] = nc;
        std::vector<Item> &v = groups[g];
        std::stable_sort(v.begin(), v.end(), [](const Item &a, const Item &b) {
            return a.base != b.base ? a.base < b.base : a.off < b.off;
        });
        size_t at = 0;
        while (at < v.size()) {
            if (nc >= kMaxChunks) return fail(kErrPlanTooLarge, "stream plan needs more than %d chunks", kMaxChunks);
            Chunk &c = plan.ch[nc++];
            c.base = (uint64_t)(uintptr_t)v[at].base;
            c.stride = (int32_t)v[at].stride;
            if ((int64_t)c.stride != v[at].stride) return fail(IRLOSC_ERR_INVALID, "instance stride too large for the streaming kernel");
            int l = 0;
            for (; l < 8 && at < v.size() && v[at].base == (const void *)(uintptr_t)c.base && v[at].stride == c.stride; ++l, ++at) {
                if (v[at].off > INT32_MAX) return fail(IRLOSC_ERR_INVALID, "record too large for the streaming kernel");
                c.off[l] = (int32_t)v[at].off;
                c.dst[l] = (int16_t)v[at].dst;
            }
            for (; l < 8; ++l) { c.off[l] = c.off[0]; c.dst[l] = (int16_t)trash; }
        }
    }
    plan.first[kGroups] = nc;
    plan.n_chunks = nc;
    for (int c = 0; c < nc; ++c) {               // distinct arrays, for the whole-tile L2 prefetch
        bool seen = false;
        for (int a = 0; a < plan.n_arrays; ++a) seen = seen || plan.arr_base[a] == plan.ch[c].base;
        if (!seen && plan.n_arrays < kMaxArrays) {
            plan.arr_base[plan.n_arrays] = plan.ch[c].base;
            plan.arr_stride[plan.n_arrays++] = plan.ch[c].stride;
        }
    }
    return IRLOSC_OK;
}

}  // namespace irlosc

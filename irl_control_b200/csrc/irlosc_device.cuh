// Device-side building blocks shared by every OSC kernel: the flattened controller
// parameters, quaternion / Euler helpers and the per-device task-space law.
//
// Reference being restated (file:line of ir-lab/irl_control):
//   osc.py:101-118  calc_error        -> device_pose_error
//   osc.py:70-99    __limit_vel       -> device_task_signal (saturation + gains)
//   osc.py:156-181  per-device loop   -> device_task_signal
//   device.py:135-170 F/T rotation    -> rotate_wrench
// transforms3d ('sxyz', wxyz) helpers follow the published 0.4.x algorithm (see oracle/t3d.py).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/irlosc.h"

// Functions shared with the host-compiled test harness (tests/host_fused) are host + device;
// nothing in the product calls them on the host.
#define IRLOSC_HD __host__ __device__ __forceinline__

namespace irlosc {

constexpr double kDetThreshold = 1e-4;   // osc.py:51
constexpr double kPinvRcond = 1e-5;      // osc.py:55 (threshold * 0.1)
constexpr double kEps = 2.220446049250313e-16;

// Controller description as the kernels see it (built once in irlosc_create).
struct KDevice {
    int32_t dof[6];
    int32_t row0;             // first stacked task row of this device
    int32_t kdev;             // number of controlled rows
    int32_t any_xyz, any_abg; // np.sum(ctrlr_dof_xyz) > 0 / abg (osc.py:108,113)
    int32_t has_max_vel;
    int32_t ctrl0;            // first packed ctrl slot of this device
    int32_t n_ctrl;
    int32_t n_joints_all;
    int32_t dx_idx[6];
    double max_vel[2];
    double kp, kv, ko;
    double kv_over_kp, kv_over_ko;   // saturation gains of osc.py:83,90 (max_vel / kp * kv)
    double gain[6];           // task_space_gains (osc.py:37)
    double lamb[6];           // gains / kv      (osc.py:39)
    double stiff[6];          // k + [1,1,1]     (osc.py:160)
    double damp[6];           // d + [1,1,1]     (osc.py:161)
    uint32_t joint_mask;      // bit j set <=> j in joint_ids_all
    int32_t ee_joint;         // deepest joint moving the EE body (-1 unknown)
    int8_t actuator[IRLOSC_MAX_N];
};

struct KParams {
    int32_t n, D, k, n_ctrl;
    int32_t use_g, admittance, has_nullspace;
    double nullspace_kv;
    int8_t row_dev[IRLOSC_MAX_K];   // stacked row -> device
    int8_t row_comp[IRLOSC_MAX_K];  // stacked row -> component 0..5 of [xyz, abg]
    int32_t has_topology, check_topology;
    int8_t joint_parent[IRLOSC_MAX_N];
    KDevice dev[IRLOSC_MAX_DEVICES];
};

// Per-step pointers (device memory) - mirror of irlosc_io with resolved strides.
struct KIo {
    const double *M; int32_t m_layout; int32_t ldm; int64_t m_stride;
    const double *J; int32_t j_layout; int32_t ldj; int64_t j_stride;
    const double *dq, *bias, *ee_xyz, *ee_quat, *target_xyz, *target_quat;
    const double *target_vel, *max_vel, *ft_xmat, *ft_raw;
    double *u_all, *ctrl; uint8_t *status;
    int32_t n_gather; int64_t gather_offset;
    double *ctrl_gather[IRLOSC_MAX_PEERS];
    double *ctrl_mc;
};

// Packed control output (osc.py:203-208): local array plus, when the gather is fused, the same
// row in every peer's gathered array: one multimem store through the NVSwitch multicast mapping,
// or plain stores to each peer-mapped array (both travel over NVLink).
__device__ __forceinline__ void multimem_st(double *p, double v) {
    asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void multimem_st(double2 *p, double2 v) {      // 16-byte form (v2.f64 does not exist)
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(__double2loint(v.x)),
                 "r"(__double2hiint(v.x)), "r"(__double2loint(v.y)), "r"(__double2hiint(v.y)) : "memory");
}
__device__ __forceinline__ void store_ctrl(const KIo &io, int n_ctrl, int64_t inst, int c, double v) {
    io.ctrl[inst * n_ctrl + c] = v;
    if (io.ctrl_mc) { multimem_st(io.ctrl_mc + (io.gather_offset + inst) * n_ctrl + c, v); return; }
    for (int g = 0; g < io.n_gather; ++g) io.ctrl_gather[g][(io.gather_offset + inst) * n_ctrl + c] = v;
}

// ---------------------------------------------------------------- fast fp64 reciprocal / sqrt
// IEEE division and sqrt expand to ~25-30 instructions with a slow path each; the task law has a
// dozen of them on one or two lanes per instance.  Hardware seed + Newton steps give <= 1 ulp-level
// results in 5-7 instructions (differences vs. the reference's correctly rounded numpy values are
// ~1e-16 relative, far inside the parity tolerance).
IRLOSC_HD double fast_rcp(double d) {
#ifndef __CUDA_ARCH__
    return 1.0 / d;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
#endif
}
IRLOSC_HD double fast_sqrt(double x) {     // x >= 0 and not denormal-small; sqrt(0) = 0
#ifndef __CUDA_ARCH__
    return sqrt(x);
#else
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    // two Newton steps on r ~ 1/sqrt(x), then one correction of s = x r
    double h = 0.5 * r;
    double e = fma(-x * r, h, 0.5);
    r = fma(r, e, r);
    h = 0.5 * r;
    e = fma(-x * r, h, 0.5);
    r = fma(r, e, r);
    double s = x * r;
    s = fma(fma(-s, s, x), 0.5 * r, s);
    return x > 0.0 ? s : 0.0;
#endif
}

// ---------------------------------------------------------------- rotations
IRLOSC_HD void quat_mul(const double *a, const double *b, double *o) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    o[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}

// quat2euler(q) = mat2euler(quat2mat(q)), static x-y-z angles.
IRLOSC_HD void quat_to_euler_sxyz(const double *q, double *e) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double nq = w * w + x * x + y * y + z * z;
    double m00 = 1.0, m10 = 0.0, m20 = 0.0, m21 = 0.0, m22 = 1.0, m11 = 1.0, m12 = 0.0;
    if (nq >= kEps) {
        const double s = 2.0 * fast_rcp(nq);
        const double X = x * s, Y = y * s, Z = z * s;
        const double wX = w * X, wY = w * Y, wZ = w * Z;
        const double xX = x * X, xY = x * Y, xZ = x * Z;
        const double yY = y * Y, yZ = y * Z, zZ = z * Z;
        m00 = 1.0 - (yY + zZ);
        m10 = xY + wZ;
        m20 = xZ - wY;
        m21 = yZ + wX;
        m22 = 1.0 - (xX + yY);
        m11 = 1.0 - (xX + zZ);
        m12 = yZ - wX;
    }
    const double cy = fast_sqrt(m00 * m00 + m10 * m10);
    if (cy > 4.0 * kEps) {
        e[0] = atan2(m21, m22);
        e[1] = atan2(-m20, cy);
        e[2] = atan2(m10, m00);
    } else {
        e[0] = atan2(-m12, m11);
        e[1] = atan2(-m20, cy);
        e[2] = 0.0;
    }
}

// osc.py:101-118 - unmasked 6-vector [ee - target ; euler(conj(q_d * conj(q_ee)))].
IRLOSC_HD void device_pose_error(const KDevice &dv, const double *ee_xyz,
                                                  const double *ee_quat, const double *t_xyz,
                                                  const double *t_quat, double *u) {
#pragma unroll
    for (int i = 0; i < 6; ++i) u[i] = 0.0;
    if (dv.any_xyz) {
#pragma unroll
        for (int i = 0; i < 3; ++i) u[i] = ee_xyz[i] - t_xyz[i];
    }
    if (dv.any_abg) {
        const double n2 = t_quat[0] * t_quat[0] + t_quat[1] * t_quat[1] + t_quat[2] * t_quat[2] + t_quat[3] * t_quat[3];
        const double inn = fast_rcp(fast_sqrt(n2));
        const double qd[4] = {t_quat[0] * inn, t_quat[1] * inn, t_quat[2] * inn, t_quat[3] * inn};
        const double qc[4] = {ee_quat[0], -ee_quat[1], -ee_quat[2], -ee_quat[3]};
        double qr[4];
        quat_mul(qd, qc, qr);
        const double qrc[4] = {qr[0], -qr[1], -qr[2], -qr[3]};
        quat_to_euler_sxyz(qrc, u + 3);
    }
}

// device.py:135-170 - world-frame [force ; torque] of one device.
IRLOSC_HD void rotate_wrench(const double *R, const double *raw, double *out) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 3; ++i)
            out[3 * h + i] = R[3 * i + 0] * raw[3 * h + 0] + R[3 * i + 1] * raw[3 * h + 1] +
                             R[3 * i + 2] * raw[3 * h + 2];
}

// osc.py:159-177 for one device: pose error -> saturated / gained task signal.
//   u      out: 6-vector AFTER gains, stiffness and (if taken) the velocity-tracking term
//   returns true when the non-zero-target-velocity branch was taken (osc.py:175-177)
//   dx / k are only read on that branch; *dx_oob is set when dx_idx runs past k (IndexError in numpy)
IRLOSC_HD bool device_task_signal(const KDevice &dv, const double *ee_xyz,
                                                   const double *ee_quat, const double *t_xyz,
                                                   const double *t_quat, const double *t_vel,
                                                   const double *max_vel, const double *dx, int k,
                                                   double *u, bool *dx_oob) {
    device_pose_error(dv, ee_xyz, ee_quat, t_xyz, t_quat, u);
    if (dv.has_max_vel) {
        // osc.py:79-94
        double sc_xyz = 1.0, sc_abg = 1.0;
        const double n_xyz = fast_sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        const double sat_xyz = max_vel[0] * dv.kv_over_kp;
        if (n_xyz > sat_xyz) sc_xyz = sat_xyz * fast_rcp(n_xyz);
        const double n_abg = fast_sqrt(u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);
        const double sat_abg = max_vel[1] * dv.kv_over_ko;
        if (n_abg > sat_abg) sc_abg = sat_abg * fast_rcp(n_abg);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double sc = (i < 3) ? sc_xyz : sc_abg;
            u[i] = ((dv.kv * sc) * dv.lamb[i]) * u[i];   // kv * scale * lamb * u_task, left to right
            u[i] *= dv.stiff[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) u[i] *= dv.gain[i] * dv.stiff[i];   // osc.py:167-168
    }
    // osc.py:172-177: "np.all(target_vel) == 0" is False only if ALL six entries are non-zero
    bool tracking = false;
    if (t_vel != nullptr) {
        tracking = true;
#pragma unroll
        for (int i = 0; i < 6; ++i) tracking = tracking && (t_vel[i] != 0.0);
    }
    if (tracking) {
        int r = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (dv.dof[i]) {
                const int src = dv.dx_idx[r++];
                if (src >= k) {
                    *dx_oob = true;
                } else {
                    u[i] += dv.kv * (dx[src] - t_vel[i]) * dv.damp[i];
                }
            }
        }
    }
    return tracking;
}

}  // namespace irlosc

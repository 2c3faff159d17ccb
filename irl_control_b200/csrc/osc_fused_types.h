// Plain data types of the fused state provider + OSC step (kernels: osc_fused.cuh).
#pragma once
#include <cstdint>
#include "../../include/irlosc.h"

namespace irlosc {
namespace fused {

constexpr int kN = 25;

struct KJoint {
    double Q0[9], P1[9], P2[9];   // R_local(q) = cos q Q0 + sin q P1 + P2 (row-major; fixed frame folded in)
    double pos[3];                // body origin in the parent joint body's frame
    double axp[3];                // hinge axis in the parent joint body's frame
    double mass, com[3], ic[6];   // lumped body: COM (body frame), inertia about it (xx yy zz xy xz yz)
};
struct KFrame {
    int32_t joint, has;
    double pos[3], R[9];
};
struct KModel {
    double gravity[3], pad_;
    KJoint stand;                 // joint 0
    KJoint arm[2][6];             // joints 1..6 / 13..18
    KJoint grip[2][2][3];         // [arm][half]: g0 (child of arm joint 6), g1 (child of g0), g2 (child of arm joint 6)
    KFrame ee[IRLOSC_MAX_DEVICES], ft[IRLOSC_MAX_DEVICES];
};
struct FRoles {
    int dev_arm[2], dev_base, row_arm[2], row_base;
    int8_t joint_slot[IRLOSC_MAX_N];   // packed ctrl slot of joint j (osc.py:203-208), -1 = not returned
    // 1 when, within each arm's six arm joints and within each gripper half's three joints, every joint belongs to the
    // same set of target devices (true for every shipped configuration): the velocity-term coefficient of osc.py:174
    // is then evaluated once per group instead of once per joint
    int32_t uniform_owner;
};
struct FIo {
    const double *q, *dq, *target_xyz, *target_quat, *target_vel, *max_vel, *ft_raw;
    double *ctrl, *u_all;
    uint8_t *status;
    double *ee_xyz, *ee_quat;
    // action-sequence mode (irlosc_step_sequence): episode state, see irlosc_sequence_io
    const double *wp_xyz, *wp_quat;
    int32_t *seq_action, *seq_entered, *seq_timer;
    double *seq_err, *seq_mv0, *seq_tgt_xyz, *seq_tgt_quat;
    // waypoint-cycling mode (irlosc_step_waypoints), see irlosc_waypoints_io
    const double *wps;
    int32_t *wp_idx;
};
struct KAction {
    int32_t type, grip_steps;
    double kp, max_error, min_speed, max_speed, gripper_force;
};
struct KSeq {
    int32_t n_actions, active_dev, gripper_slot;
    int32_t mode;                      // 0 action sequence (insertion_task.py), 1 waypoint cycling (gain_test.py)
    int32_t W, n_wp[IRLOSC_MAX_DEVICES], pad_;
    double threshold;
    double passive_quat[4];
    KAction act[IRLOSC_MAX_ACTIONS];
};
// Record a thread hands to its warp when it cannot decide the task-space solve by itself (osc_tail.cuh).
template <int KD, bool HAS_BASE>
struct Rec {                       // record layout in doubles
    static constexpr int K = 2 * KD + (HAS_BASE ? 1 : 0);
    static constexpr int A = 0;                    // K x K row-major
    static constexpr int G = A + K * K;            // rhs g
    static constexpr int BASE = G + K;             // base[13]: stand, arm0 joints 1..6, arm1 joints 1..6
    static constexpr int JST = BASE + 13;          // J[r][stand], r < K
    static constexpr int JARM = JST + K;           // J[row_arm[a] + cr][arm joint i]: [a][i][cr]
    static constexpr int ABAD = JARM + 2 * 6 * KD; // 0.0: pinv branch known to be taken; 1.0: the eigen-solver decides from its determinant
    static constexpr int SIZE = ABAD + 1;
};

}  // namespace fused
}  // namespace irlosc

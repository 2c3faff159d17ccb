// Per-instance action-sequence state machine of the insertion demo, run inside the fused step
// (examples/insertion_task.py:146-179 send_forces, 190-204 grip, 279-297 go_to_waypoint,
// 312-317 run_sequence).  One call = the decisions the reference takes between two `generate` calls.
#pragma once
#include <cmath>
#include "irlosc_device.cuh"
#include "osc_fused_types.h"

namespace irlosc {
namespace fused {

struct SeqResult {
    bool entered_wp;        // a WP action started this step: the passive arm latches its current xyz
    double gripper_force;   // 0 = leave the gripper slot to the controller
};

// ee_p / ee_q: pose of the active arm's EE for the state this step is computed from (the reference
// evaluates calc_error right after sim.step(), i.e. on the same state the next generate() sees).
IRLOSC_HD SeqResult seq_advance(const KSeq &Q, const KDevice &dv, const FIo &io, int64_t inst, int D, const double *ee_p,
                                const double *ee_q) {
    SeqResult res{false, 0.0};
    int action = io.seq_action[inst], entered = io.seq_entered[inst], timer = io.seq_timer[inst];
    double err = io.seq_err[inst], mv0 = io.seq_mv0[inst];
    double *txyz = io.seq_tgt_xyz + (inst * D + Q.active_dev) * 3;
    double *tquat = io.seq_tgt_quat + (inst * D + Q.active_dev) * 4;
    auto pose_err = [&]() {                                   // np.linalg.norm(calc_error(target, device))
        double u[6];
        device_pose_error(dv, ee_p, ee_q, txyz, tquat, u);
        return sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);
    };
    for (int guard = 0; guard <= Q.n_actions; ++guard) {
        if (action >= Q.n_actions) break;                     // sequence finished: hold the last targets
        const KAction &A = Q.act[action];
        if (A.type == IRLOSC_ACT_WP) {
            if (!entered) {                                   // set_waypoint_targets + errors = inf
                entered = 1;
                res.entered_wp = true;
                for (int i = 0; i < 3; ++i) txyz[i] = io.wp_xyz[(inst * Q.n_actions + action) * 3 + i];
                for (int i = 0; i < 4; ++i) tquat[i] = io.wp_quat[(inst * Q.n_actions + action) * 4 + i];
                err = HUGE_VAL;
            } else {
                err = pose_err();
            }
            if (err > A.max_error) {                          // one more pass of the while loop
                mv0 = fmax(A.min_speed, fmin(A.max_speed, A.kp * err));
                res.gripper_force = A.gripper_force;
                break;
            }
        } else {
            if (!entered) { entered = 1; timer = A.grip_steps; }
            else err = pose_err();                            // send_forces keeps the error current
            if (timer > 0) {
                --timer;
                res.gripper_force = A.gripper_force;
                break;
            }
        }
        ++action;                                             // action done: the next one starts this step
        entered = 0;
    }
    io.seq_action[inst] = action;
    io.seq_entered[inst] = entered;
    io.seq_timer[inst] = timer;
    io.seq_err[inst] = err;
    io.seq_mv0[inst] = mv0;
    return res;
}

}  // namespace fused
}  // namespace irlosc

/*
 * irlosc.h - C ABI of the B200-native batched operational-space controller.
 *
 * The reference (ir-lab/irl_control) has no FFI layer: its per-timestep hot
 * path is the Python method `OSC.generate` (irl_control/osc.py:120-210) fed by
 * the state pulls of `Robot.get_all_states` (irl_control/robot.py:125-136) and
 * `Device.get_all_states` (irl_control/device.py:183-197).  This header is the
 * boundary a maintainer would bind from `OSC.generate` (ctypes stub shown in
 * INTEGRATION.md): plain pointers and sizes, no torch / numpy types.
 *
 * One handle == one controller configuration (what the `Device`, `Robot` and
 * `OSC` constructors resolve: index maps, DoF masks, gains) on the CUDA
 * device that is current when `irlosc_create` is called.  A step evaluates the
 * whole control law for B independent robot instances.  All arithmetic is
 * IEEE float64, like the reference's numpy / MuJoCo mjtNum path.
 *
 * Threading: a handle is thread-compatible, not thread-safe (one caller at a
 * time per handle), matching the reference ("one caller thread runs
 * generate").  The device-pointer steps (`irlosc_step`,
 * `irlosc_step_tiles`, `irlosc_step_fused`, `irlosc_step_sequence`,
 * `irlosc_step_waypoints`) are ONE kernel launch each, never synchronise with
 * the host and never allocate; steps issued on different streams may overlap
 * freely.  (The only per-handle device state is 16 ticket counters of the pair
 * kernel, allocated by the first tile call of a handle and keyed by stream; a
 * launch rewinds its own counter.)  The `*_host` forms own staging buffers and
 * internal streams per handle and return when the outputs are valid.
 */
#ifndef IRLOSC_H_
#define IRLOSC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IRLOSC_ABI_VERSION 8
#define IRLOSC_MAX_DEVICES 4   /* target devices per controller (DualUR5: base + 2 arms) */
#define IRLOSC_MAX_N 32        /* robot DoF, Robot.num_joints_total (robot.py:32); DualUR5: 25 */
#define IRLOSC_MAX_K 24        /* stacked task rows, sum of ctrlr_dof over targets; DualUR5: <= 13 */
#define IRLOSC_MAX_PEERS 8     /* GPUs of one NVSwitch domain the result gather can be fused over */

/* return codes */
#define IRLOSC_OK 0
#define IRLOSC_ERR_INVALID 1   /* bad argument / parameter block          */
#define IRLOSC_ERR_CUDA 2      /* CUDA runtime error, see irlosc_last_error */
#define IRLOSC_ERR_NOMEM 3

/* per-instance status byte written by a step (bit field) */
#define IRLOSC_ST_PINV 0x01         /* |det(J M^-1 J^T)| < 1e-4: pinv(rcond=1e-5) branch (osc.py:52-55) */
#define IRLOSC_ST_M_NOT_PD 0x02     /* M had a non-positive pivot: outputs are NaN                      */
#define IRLOSC_ST_EIGEN 0x04        /* task-space solve finished by the warp-cooperative Jacobi eigen-solver
                                       (the thread's own inertia / deflation resolution could not decide it) */
#define IRLOSC_ST_VEL_BRANCH 0x08   /* >=1 device took the non-zero target-velocity branch (osc.py:175-177) */
#define IRLOSC_ST_DX_RANGE 0x10     /* that branch indexed dx out of range: the reference raises IndexError
                                       (robot.py:52-55 vs osc.py:150,176); outputs are NaN               */
#define IRLOSC_ST_SPARSITY 0x20     /* check_topology was set and M / J had a non-zero where the declared
                                       kinematic tree says zero; outputs are NaN                         */

/* layouts of the inertia / Jacobian inputs */
#define IRLOSC_M_DENSE 0        /* [B][ldm rows used: n][ldm]  row-major n x n block, row stride ldm   */
#define IRLOSC_M_PACKED 1       /* [B][n(n+1)/2] lower triangle, row-major: (i,j<=i) at i(i+1)/2+j     */
#define IRLOSC_M_QM 2           /* [B][m_stride >= nM] MuJoCo's own sparse inertia, mjData.qM - what robot.py:69
                                   hands to mj_fullM: dof i owns M[i][i], M[i][parent(i)], M[i][parent(parent(i))],
                                   ... down to its root, stored from dof_Madr[i] = sum of depth(r) for r < i
                                   (depth counts i itself).  Needs has_topology (joint_parent = dof_parentid of the
                                   robot's dofs, which must be the scene's first n dofs); nM = sum of depths (155
                                   for the DualUR5), m_stride = the scene's nM */
#define IRLOSC_J_ROWS 0         /* [B][k][ldj]   only the controlled rows, target order (osc.py:136-138) */
#define IRLOSC_J_FULL6 1        /* [B][D][6][ldj] full [jacp;jacr] per target device (device.py:125-130);
                                   rows with ctrlr_dof == 0 are never read                              */

/* One target device, in TARGET order (the order of the dict given to generate). */
typedef struct irlosc_device_params {
    int32_t ctrlr_dof[6];                 /* device.py:33-36: xyz then abg mask                        */
    int32_t n_joints_all;                 /* len(Device.joint_ids_all)            (device.py:69)       */
    int32_t joint_ids_all[IRLOSC_MAX_N];  /* robot-local joint positions          (robot.py:64)        */
    int32_t n_ctrl;                       /* len(Device.actuator_trnids)          (device.py:73-74)    */
    int32_t actuator_trnids[IRLOSC_MAX_N];/* robot-local joints whose force is returned (osc.py:207)   */
    int32_t dx_idx[6];                    /* Robot J_idxs[name] (robot.py:54), first sum(ctrlr_dof) used;
                                             may exceed k-1, see IRLOSC_ST_DX_RANGE                     */
    int32_t has_max_vel;                  /* Device.max_vel is not None           (osc.py:163)         */
    double max_vel[2];                    /* default when io.max_vel == NULL      (device.py:31)       */
    double kp, kv, ko;                    /* controller_configs entry             (osc.py:36)          */
    double k[3], d[3];                    /* stiffness / damping, xyz part        (osc.py:160-161)     */
    /* What OSC.__init__ stored in the controller config (osc.py:35-39) - the reference computes these ONCE and reads
     * them back every step, while kp / kv / ko are read fresh (osc.py:76,170): a caller that edits a gain after
     * construction gets the stale vectors.  has_gain_vectors == 0: derived from kp / kv / ko above. */
    int32_t has_gain_vectors;
    int32_t reserved2_;
    double task_space_gains[6];           /* [kp] * 3 + [ko] * 3 at construction  (osc.py:37)          */
    double lamb[6];                       /* task_space_gains / kv at construction (osc.py:39)         */
    int32_t ee_joint;                     /* robot-local id of the deepest joint that moves the EE body,
                                             i.e. the last entry of Device.joint_ids (device.py:62-64);
                                             -1 = unknown (only read when has_topology)                 */
    int32_t reserved_;
} irlosc_device_params;

typedef struct irlosc_params {
    int32_t abi_version;                  /* IRLOSC_ABI_VERSION                                        */
    int32_t n;                            /* Robot.num_joints_total                                    */
    int32_t n_devices;                    /* number of targets                                         */
    int32_t use_g;                        /* OSC(use_g=...)                       (osc.py:190)         */
    int32_t admittance;                   /* OSC(admittance=...)                  (osc.py:184)         */
    int32_t has_nullspace;                /* nullspace_config is not None         (osc.py:195)         */
    double nullspace_kv;                  /* nullspace_config['kv']               (osc.py:196)         */
    /* Optional kinematic-tree description (what Device.__init__ walks, device.py:41-64).  When
     * has_topology != 0 the caller guarantees the sparsity every MuJoCo state has:
     *   M[i][j] == 0 unless joint i is an ancestor of joint j or vice versa (mj_fullM),
     *   row r of a device's Jacobian == 0 outside the ancestors-or-self of its ee_joint (mj_jacBody).
     * Kernels specialised for a topology (DualUR5) then skip the structural zeros - the result is
     * the same as the dense elimination.  check_topology makes the step verify the zeros
     * (IRLOSC_ST_SPARSITY).  Without topology the dense kernels are used. */
    int32_t has_topology;
    int32_t check_topology;
    int32_t joint_parent[IRLOSC_MAX_N];   /* robot-local parent joint, -1 for a root                   */
    irlosc_device_params dev[IRLOSC_MAX_DEVICES];
} irlosc_params;

/*
 * Per-step arrays; every array has the instance index as its leading axis and
 * is densely packed unless a stride is given.  D = n_devices, n = params.n,
 * k = sum of ctrlr_dof.  For `irlosc_step` these are DEVICE pointers, for
 * `irlosc_step_host` HOST pointers.  Optional pointers may be NULL.
 */
typedef struct irlosc_io {
    const double *M;          /* RobotState.M (robot.py:68-72)                                        */
    int32_t m_layout;         /* IRLOSC_M_DENSE | IRLOSC_M_PACKED | IRLOSC_M_QM                        */
    int32_t ldm;              /* dense: row stride in doubles (>= n; scene nv if the block is a view)  */
    int64_t m_stride;         /* doubles between instances; 0 = tight (ldm*n dense, n(n+1)/2 packed, nM qM) */
    const double *J;          /* RobotState.J stacked for the targets                                  */
    int32_t j_layout;         /* IRLOSC_J_ROWS | IRLOSC_J_FULL6                                        */
    int32_t ldj;              /* row stride in doubles (>= n)                                          */
    int64_t j_stride;         /* doubles between instances; 0 = tight                                  */
    const double *dq;         /* [B][n]    RobotState.DQ (robot.py:60-65)                              */
    const double *bias;       /* [B][n]    sim.data.qfrc_bias[joint_ids_all] (osc.py:191); NULL iff !use_g */
    const double *ee_xyz;     /* [B][D][3] DeviceState.EE_XYZ  (device.py:93)                          */
    const double *ee_quat;    /* [B][D][4] DeviceState.EE_QUAT (device.py:95), w x y z                 */
    const double *target_xyz; /* [B][D][3] Target.get_xyz()    (utils.py:17)                           */
    const double *target_quat;/* [B][D][4] Target.get_quat()   (utils.py:23)                           */
    const double *target_vel; /* [B][D][6] [xyz_vel, abg_vel] (osc.py:172); NULL = all zero            */
    const double *max_vel;    /* [B][D][2] Device.max_vel per instance (insertion_task.py:294); NULL = params */
    const double *ft_xmat;    /* [B][D][9] site_xmat of the F/T frame (device.py:135-143); NULL iff !admittance */
    const double *ft_raw;     /* [B][D][6] sensor-frame force|torque (device.py:150-167); NULL iff !admittance  */
    double *u_all;            /* [B][n]      out, optional: joint-space signal before packing (osc.py:152-200) */
    double *ctrl;             /* [B][n_ctrl] out: u_all[actuator_trnids] per target, concatenated (osc.py:203-208) */
    uint8_t *status;          /* [B]         out, optional: IRLOSC_ST_* bits                           */
    /* Fused result gather (irlosc_step only).  When the batch is sharded over the GPUs of one box,
     * the step kernel can write its packed ctrl rows straight into the gathered [B_total][n_ctrl]
     * array of every rank through peer-mapped pointers (NVLink 5 / NVSwitch stores) instead of a
     * separate all-gather: row (gather_offset + i) of each ctrl_gather[g].  The caller provides the
     * cross-GPU barrier before anyone reads the gathered arrays. */
    int32_t n_gather;         /* number of entries of ctrl_gather, 0 = no fused gather               */
    int32_t reserved_;
    int64_t gather_offset;    /* first gathered row of this rank's shard                             */
    double *ctrl_gather[IRLOSC_MAX_PEERS];
    /* Same gather through an NVSwitch multicast (NVLS) mapping of the gathered array: ONE
     * multimem store per row reaches every GPU bound to the mapping (this one included), so the
     * NVLink egress of a rank is its own shard once instead of once per peer.  When non-NULL it
     * is used instead of ctrl_gather[] (n_gather may be 0). */
    double *ctrl_multicast;
} irlosc_io;

/* ------------------------------------------------------------------------------------------
 * Fused state provider: the step BEFORE the path.  The reference pulls M, J,
 * qfrc_bias and the EE poses out of MuJoCo every timestep (robot.py:68-72 mj_fullM,
 * device.py:115-133 mj_jacBody via get_body_jacp/jacr, osc.py:191 qfrc_bias, device.py:93-95
 * xpos/xquat, device.py:135-143 site_xmat).  With a rigid-body description of the robot the
 * library computes all of that on the GPU from (q, dq) inside the same kernel as the control
 * law, so only ~0.6 KB per instance crosses HBM / PCIe instead of ~4.9 KB.
 *
 * The description is the MuJoCo model reduced to its joints: bodies without joints are folded
 * into the nearest ancestor body that has one (frames composed, inertias lumped).  All hinge,
 * anchored at the body origin (true for every joint of scenes/dual_ur5.xml:55-251). */
typedef struct irlosc_joint_model {
    int32_t parent;           /* robot-local parent joint, -1 = attached to the world               */
    int32_t reserved_;
    double pos[3];            /* body frame origin in the parent joint's body frame (body_pos chain) */
    double quat[4];           /* body frame orientation likewise, w x y z          (body_quat chain) */
    double axis[3];           /* hinge axis in the body frame, unit                 (jnt_axis)       */
    double mass;              /* lumped mass of the body and everything welded to it (body_mass)     */
    double com[3];            /* centre of mass, body frame                          (body_ipos)     */
    double inertia[6];        /* about the COM, body-frame axes: xx yy zz xy xz yz   (body_inertia, body_iquat) */
} irlosc_joint_model;

typedef struct irlosc_frame_model {
    int32_t joint;            /* robot-local joint whose body carries the frame; -1 = frame absent   */
    int32_t reserved_;
    double pos[3];            /* in that body's frame                                                */
    double quat[4];           /* w x y z                                                             */
} irlosc_frame_model;

typedef struct irlosc_model {
    int32_t n_joints;         /* == params.n                                                         */
    int32_t reserved_;
    double gravity[3];        /* mjOption.gravity, world frame                                       */
    irlosc_joint_model joint[IRLOSC_MAX_N];
    irlosc_frame_model ee[IRLOSC_MAX_DEVICES]; /* per TARGET device: its EE body (device.py:93-95,125-128) */
    irlosc_frame_model ft[IRLOSC_MAX_DEVICES]; /* per TARGET device: its F/T site (device.py:135-143);
                                                  joint = -1: no sensor, force = torque = 0 (device.py:162-170) */
} irlosc_model;

/* Per-step arrays of the fused step; DEVICE pointers for irlosc_step_fused, HOST pointers for
 * irlosc_step_fused_host.  Same conventions as irlosc_io. */
typedef struct irlosc_fused_io {
    const double *q;          /* [B][n]    sim.data.qpos[joint_ids_all]                              */
    const double *dq;         /* [B][n]    sim.data.qvel[joint_ids_all]   (robot.py:60-65)           */
    const double *target_xyz; /* [B][D][3]                                                           */
    const double *target_quat;/* [B][D][4]                                                           */
    const double *target_vel; /* [B][D][6] optional                                                  */
    const double *max_vel;    /* [B][D][2] optional                                                  */
    const double *ft_raw;     /* [B][D][6] sensor-frame force|torque; NULL iff !admittance           */
    double *ctrl;             /* [B][n_ctrl] out                                                     */
    double *u_all;            /* [B][n]      out, optional                                           */
    uint8_t *status;          /* [B]         out, optional                                           */
    double *ee_xyz;           /* [B][D][3]   out, optional: DeviceState.EE_XYZ the step computed     */
    double *ee_quat;          /* [B][D][4]   out, optional: DeviceState.EE_QUAT (w x y z, w >= 0 branch) */
} irlosc_fused_io;

/* ------------------------------------------------------------------------------------------
 * Action sequences: the step AFTER the path.  The reference's insertion demo drives
 * `OSC.generate` from a per-robot state machine - `run_sequence` over WP / GRIP actions
 * (examples/insertion_task.py:312-317), `go_to_waypoint` (279-297: loop until the pose error of the
 * active arm is <= max_error, max_vel[0] = clip(kp * error, min_speed, max_speed) every step),
 * `grip` (190-204: hold for a duration), `send_forces` (146-179: gripper force override, error
 * update) and `set_waypoint_targets` (206-268: the passive arm holds the xyz it has when a WP starts).
 * `irlosc_step_sequence` runs that state machine per instance inside the fused step, so a batch of
 * episodes advances without host round trips. */
#define IRLOSC_MAX_ACTIONS 16
#define IRLOSC_ACT_WP 0
#define IRLOSC_ACT_GRIP 1

typedef struct irlosc_action {
    int32_t type;             /* IRLOSC_ACT_WP | IRLOSC_ACT_GRIP                   (insertion_task.py:22-29) */
    int32_t grip_steps;       /* GRIP: control steps the action lasts (gripper_duration / step period)        */
    double kp, max_error, min_speed_xyz, max_speed_xyz;     /* WP parameters      (insertion_task.py:88-96) */
    double gripper_force;     /* written to the gripper's ctrl slot when non-zero  (insertion_task.py:160-161) */
} irlosc_action;

typedef struct irlosc_sequence {
    int32_t n_actions;
    int32_t active_device;    /* target-order index of the active arm              (insertion_task.py:115-128) */
    int32_t gripper_slot;     /* packed ctrl slot that receives gripper_force (sim.data.ctrl[7] / [14]), -1 none */
    int32_t reserved_;
    double passive_quat[4];   /* DEFAULT_EE_QUAT, target orientation of the passive arm (insertion_task.py:213) */
    irlosc_action action[IRLOSC_MAX_ACTIONS];
} irlosc_sequence;

/* Per-instance episode state and waypoints (device pointers). */
typedef struct irlosc_sequence_io {
    const double *wp_xyz;     /* [B][n_actions][3] active-arm target of every WP action (set_waypoint_targets) */
    const double *wp_quat;    /* [B][n_actions][4]                                                             */
    int32_t *action;          /* [B] in/out: current action, n_actions = sequence finished (holds last targets) */
    int32_t *entered;         /* [B] in/out: 1 once the current action has run its first step                  */
    int32_t *timer;           /* [B] in/out: remaining steps of a GRIP action                                  */
    double *err;              /* [B] in/out: self.errors[active arm] (inf when a WP starts)                    */
    double *max_vel0;         /* [B] in/out: active_arm.max_vel[0], persists across actions (insertion_task.py:294) */
    double *target_xyz;       /* [B][D][3] in/out: self.targets, used instead of irlosc_fused_io.target_xyz    */
    double *target_quat;      /* [B][D][4] in/out                                                              */
} irlosc_sequence_io;

/* Waypoint cycling of the gain_test demo (examples/gain_test.py:134-162): every arm walks its own
 * list of xyz waypoints; after each generate the distance |EE_XYZ - target| is compared with a
 * threshold and the arm's waypoint index advances (wrapping) when it is below. */
typedef struct irlosc_waypoints_io {
    const double *wps;        /* [B][D][W][3] waypoint lists per target device (rows of devices that are not arms are ignored) */
    int32_t W;                /* allocated waypoints per device                                                */
    int32_t n_wp[IRLOSC_MAX_DEVICES]; /* used waypoints per device (right_wps.shape[0] / left_wps.shape[0])      */
    int32_t reserved_;
    double threshold;         /* threshold_ee = 0.1                                  (gain_test.py:121)        */
    int32_t *wp_idx;          /* [B][D] in/out: right_wp_idx / left_wp_idx                                     */
    double *target_xyz;       /* [B][D][3] in/out: targets[...].xyz, used instead of irlosc_fused_io.target_xyz */
    double *target_quat;      /* [B][D][4] in: Target() default [1,0,0,0] unless the caller changed it         */
} irlosc_waypoints_io;

/* ------------------------------------------------------------------------------------------
 * Batch-interleaved tiles: the native HBM layout of the step.  Instances are grouped in
 * tiles of 32; inside a tile every scalar of the state is stored for its 32 instances side by side,
 *     tiles[t][e][l] = entry e of instance 32 t + l,   e < E = irlosc_tile_entries(h),  l < 32,
 * in the order the elimination consumes them and restricted to the entries the kinematic tree makes non-zero
 * (the tree contract of has_topology).  What the reference pulls per robot (robot.py:44-72 M, J, dq;
 * osc.py:191 bias; device.py:93-95,135-170 poses, F/T) is the same information; irlosc_tile_spec lists, for
 * every entry, where it comes from, and irlosc_pack_tiles converts from the per-variable arrays of irlosc_io.
 * Instances past B in the last tile are padding (irlosc_pack_tiles repeats the last instance there). */
#define IRLOSC_TILE 32
#define IRLOSC_ARR_PAD 0        /* unused slot                                                          */
#define IRLOSC_ARR_M 1          /* M[i][j], i >= j, j an ancestor of i or i itself     (robot.py:68-72)  */
#define IRLOSC_ARR_J 2          /* J[task row i][joint j]                             (osc.py:136-138)  */
#define IRLOSC_ARR_DQ 3         /* dq[i]                                              (robot.py:60-65)  */
#define IRLOSC_ARR_BIAS 4       /* qfrc_bias[i]; zero when !use_g                     (osc.py:191)      */
#define IRLOSC_ARR_EE_XYZ 5     /* ee_xyz[target device i][j]                         (device.py:93)    */
#define IRLOSC_ARR_EE_QUAT 6    /* ee_quat[i][j], w x y z                             (device.py:95)    */
#define IRLOSC_ARR_T_XYZ 7      /* target_xyz[i][j]                                   (utils.py:17)     */
#define IRLOSC_ARR_T_QUAT 8     /* target_quat[i][j]                                  (utils.py:23)     */
#define IRLOSC_ARR_MAX_VEL 9    /* max_vel[i][j]; always present in a tile            (device.py:31)    */
#define IRLOSC_ARR_FT_XMAT 10   /* ft_xmat[i][j], row-major 3 x 3, admittance only    (device.py:135-143) */
#define IRLOSC_ARR_FT_RAW 11    /* ft_raw[i][j], admittance only                      (device.py:150-167) */

typedef struct irlosc_tile_entry {
    int32_t array;            /* IRLOSC_ARR_*                                                         */
    int32_t i, j;             /* indices as listed above (j = 0 for dq / bias)                        */
} irlosc_tile_entry;

/* Per-step arrays of the tiled step; DEVICE pointers for irlosc_step_tiles, HOST pointers for
 * irlosc_step_tiles_host.  Outputs and the fused gather as in irlosc_io. */
typedef struct irlosc_tiles_io {
    const double *tiles;      /* [ceil(B / 32)][E][32]                                                 */
    const double *target_vel; /* [B][D][6] plain array, optional (NULL = all zero, osc.py:172)         */
    double *u_all;            /* [B][n]      out, optional                                             */
    double *ctrl;             /* [B][n_ctrl] out                                                       */
    uint8_t *status;          /* [B]         out, optional                                             */
    int32_t n_gather;
    int32_t reserved_;
    int64_t gather_offset;
    double *ctrl_gather[IRLOSC_MAX_PEERS];
    double *ctrl_multicast;
} irlosc_tiles_io;

typedef struct irlosc_handle irlosc_handle;

/* Thread-local, human-readable description of the last failure on this thread. */
const char *irlosc_last_error(void);
int32_t irlosc_abi_version(void);

/* Replaces: Device.__init__ index maps (device.py:41-74), Robot.__init__ (robot.py:26-32),
 * OSC.__init__ gain precompute (osc.py:26-39).  Validates and copies `params`. */
int32_t irlosc_create(const irlosc_params *params, irlosc_handle **out);
int32_t irlosc_destroy(irlosc_handle *h);

/* Sizes derived from the parameter block. */
int32_t irlosc_num_task_rows(const irlosc_handle *h);   /* k      */
int32_t irlosc_num_ctrl(const irlosc_handle *h);        /* n_ctrl */

/* Replaces: OSC.generate (osc.py:120-210) for B instances whose state already lives in
 * device memory.  Asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream). */
int32_t irlosc_step(irlosc_handle *h, int64_t B, const irlosc_io *io_device, void *cuda_stream);

/* Same with HOST buffers: copies inputs host->device, runs the step, copies ctrl / u_all /
 * status back and returns when they are valid.  Work is pipelined in chunks over internal
 * streams; buffers from irlosc_host_alloc (pinned) make the copies asynchronous. */
int32_t irlosc_step_host(irlosc_handle *h, int64_t B, const irlosc_io *io_host);

/* Attach the rigid-body description used by the fused step.  Validates that the joint tree is the
 * DualUR5 one the kernels are specialised for and that every EE / F-T frame hangs off the joint
 * its device's Jacobian ends at. */
int32_t irlosc_set_model(irlosc_handle *h, const irlosc_model *model);

/* Replaces: Robot.get_all_states + Device.get_all_states + OSC.generate (robot.py:125-136,
 * device.py:183-197, osc.py:120-210) for B instances given only joint positions / velocities and
 * targets.  Asynchronous on `cuda_stream`, one kernel. */
int32_t irlosc_step_fused(irlosc_handle *h, int64_t B, const irlosc_fused_io *io_device, void *cuda_stream);
/* Same with HOST buffers, pipelined in chunks like irlosc_step_host. */
int32_t irlosc_step_fused_host(irlosc_handle *h, int64_t B, const irlosc_fused_io *io_host);
/* One control step of B episodes of an action sequence: state machine + fused step in one kernel.
 * io->target_xyz / target_quat are ignored (the targets live in sio). */
int32_t irlosc_step_sequence(irlosc_handle *h, int64_t B, const irlosc_fused_io *io_device, const irlosc_sequence *seq,
                             const irlosc_sequence_io *sio_device, void *cuda_stream);
/* One control step of B gain_test-style episodes: waypoint cycling + fused step in one kernel. */
int32_t irlosc_step_waypoints(irlosc_handle *h, int64_t B, const irlosc_fused_io *io_device,
                              const irlosc_waypoints_io *wio_device, void *cuda_stream);

/* Tile layout of this controller: E = entries per instance (0 when the configuration has no tile layout, i.e. it
 * is not the DualUR5 tree with two equally masked arm devices [+ the base]); the entry table (returns E, writes at
 * most `capacity` entries); doubles needed for B instances. */
int32_t irlosc_tile_entries(const irlosc_handle *h);
int32_t irlosc_tile_spec(const irlosc_handle *h, irlosc_tile_entry *out, int32_t capacity);
int64_t irlosc_tiles_doubles(const irlosc_handle *h, int64_t B);
/* Per-variable arrays (the input part of irlosc_io, any M / J layout) -> tiles.  Device pointers, asynchronous on
 * `cuda_stream`; the _host form does the same with host pointers on the calling thread (a data-layout helper for
 * callers that assemble their batch in host memory - no control arithmetic runs on the host). */
int32_t irlosc_pack_tiles(irlosc_handle *h, int64_t B, const irlosc_io *io_device, double *tiles_device, void *cuda_stream);
int32_t irlosc_pack_tiles_host(irlosc_handle *h, int64_t B, const irlosc_io *io_host, double *tiles_host);
/* Replaces: OSC.generate (osc.py:120-210) for B instances stored as tiles in device memory: ONE kernel, the
 * tile kernels (csrc/osc_lane.cuh, csrc/osc_pair.cuh; irlosc_set_tile_kernel).  Asynchronous on `cuda_stream`, never
 * synchronises with the host. */
int32_t irlosc_step_tiles(irlosc_handle *h, int64_t B, const irlosc_tiles_io *io_device, void *cuda_stream);
/* Same with HOST buffers (pinned via irlosc_host_alloc): tiles go host -> device in chunks, ctrl / u_all / status
 * come back, pipelined over internal streams; returns when the outputs are valid. */
int32_t irlosc_step_tiles_host(irlosc_handle *h, int64_t B, const irlosc_tiles_io *io_host);

/* Replaces: OSC.calc_error (osc.py:101-118), also called by insertion_task.py:173-179.
 * err[B][D][6] (unmasked), device pointers, asynchronous. */
int32_t irlosc_calc_error(irlosc_handle *h, int64_t B, const double *ee_xyz, const double *ee_quat,
                          const double *target_xyz, const double *target_quat, double *err,
                          void *cuda_stream);

/* Pinned host memory for irlosc_step_host callers. */
int32_t irlosc_host_alloc(void **ptr, int64_t bytes);
int32_t irlosc_host_free(void *ptr);

/* Kernel selection of irlosc_step (per-variable arrays): 0 = auto, 1 = generic (any n, k, layout), 2 = the tree-sparse
 * 4-lane record-staging kernel, 9 = streaming thread-per-instance kernel.  Non-default choices exist for A/B
 * measurements, see DESIGN.md.  (irlosc_step_tiles has its own selector below.) */
int32_t irlosc_set_kernel(irlosc_handle *h, int32_t which);
/* Kernel selection of irlosc_step_tiles / irlosc_step_tiles_host: IRLOSC_TILES_AUTO picks by layout and batch size
 * (DESIGN.md), IRLOSC_TILES_LANE = one thread per instance (osc_step_lane), IRLOSC_TILES_PAIR = two lanes per
 * instance, one per arm (osc_step_pair).  Both restate osc.py:120-210 on the same tiles; they differ in summation
 * order only (results agree to rounding, the status flags exactly). */
#define IRLOSC_TILES_AUTO 0
#define IRLOSC_TILES_LANE 1
#define IRLOSC_TILES_PAIR 2
int32_t irlosc_set_tile_kernel(irlosc_handle *h, int32_t which);
/* Leave `sms` streaming multiprocessors free when launching the step kernel (default 0), so that a
 * collective running on another stream (the NCCL gather of ctrl) can overlap instead of queueing
 * behind a grid that fills every SM. */
int32_t irlosc_set_sm_margin(irlosc_handle *h, int32_t sms);
/* Number of kernels this handle has launched since creation (bench.py "gpu_launches"). */
int64_t irlosc_kernel_launches(const irlosc_handle *h);
/* Name of the kernel the last step dispatched to (static string). */
const char *irlosc_last_kernel(const irlosc_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* IRLOSC_H_ */

"""ctypes binding of libirlosc.so (include/irlosc.h).

The CUDA library is the only compute path of this package: there is no CPU
fallback.  `load()` raises `NativeLibraryError` when the shared object is
missing (run `python -c "import __graft_entry__ as g; g.build()"`), and every
entry point raises `OscError` with `irlosc_last_error()` on a non-zero code.
"""
import ctypes as C
import os
from typing import Optional

MAX_DEVICES = 4
MAX_N = 32
MAX_K = 24
MAX_PEERS = 8
ABI_VERSION = 8
MAX_ACTIONS = 16
ACT_WP, ACT_GRIP = 0, 1

ST_PINV = 0x01
ST_M_NOT_PD = 0x02
ST_EIGEN = 0x04
ST_VEL_BRANCH = 0x08
ST_DX_RANGE = 0x10
ST_SPARSITY = 0x20

M_DENSE, M_PACKED, M_QM = 0, 1, 2
J_ROWS, J_FULL6 = 0, 1

KERNEL_AUTO, KERNEL_GENERIC, KERNEL_TILED = 0, 1, 2


class NativeLibraryError(RuntimeError):
    pass


class OscError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("irlosc error %d: %s" % (code, message))
        self.code = code


class DeviceParams(C.Structure):
    _fields_ = [
        ("ctrlr_dof", C.c_int32 * 6),
        ("n_joints_all", C.c_int32),
        ("joint_ids_all", C.c_int32 * MAX_N),
        ("n_ctrl", C.c_int32),
        ("actuator_trnids", C.c_int32 * MAX_N),
        ("dx_idx", C.c_int32 * 6),
        ("has_max_vel", C.c_int32),
        ("max_vel", C.c_double * 2),
        ("kp", C.c_double), ("kv", C.c_double), ("ko", C.c_double),
        ("k", C.c_double * 3), ("d", C.c_double * 3),
        ("has_gain_vectors", C.c_int32), ("reserved2_", C.c_int32),
        ("task_space_gains", C.c_double * 6), ("lamb", C.c_double * 6),
        ("ee_joint", C.c_int32), ("reserved_", C.c_int32),
    ]


class Params(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n", C.c_int32),
        ("n_devices", C.c_int32),
        ("use_g", C.c_int32),
        ("admittance", C.c_int32),
        ("has_nullspace", C.c_int32),
        ("nullspace_kv", C.c_double),
        ("has_topology", C.c_int32),
        ("check_topology", C.c_int32),
        ("joint_parent", C.c_int32 * MAX_N),
        ("dev", DeviceParams * MAX_DEVICES),
    ]


_dp = C.POINTER(C.c_double)


class Io(C.Structure):
    _fields_ = [
        ("M", C.c_void_p), ("m_layout", C.c_int32), ("ldm", C.c_int32), ("m_stride", C.c_int64),
        ("J", C.c_void_p), ("j_layout", C.c_int32), ("ldj", C.c_int32), ("j_stride", C.c_int64),
        ("dq", C.c_void_p), ("bias", C.c_void_p),
        ("ee_xyz", C.c_void_p), ("ee_quat", C.c_void_p),
        ("target_xyz", C.c_void_p), ("target_quat", C.c_void_p),
        ("target_vel", C.c_void_p), ("max_vel", C.c_void_p),
        ("ft_xmat", C.c_void_p), ("ft_raw", C.c_void_p),
        ("u_all", C.c_void_p), ("ctrl", C.c_void_p), ("status", C.c_void_p),
        ("n_gather", C.c_int32), ("reserved_", C.c_int32), ("gather_offset", C.c_int64),
        ("ctrl_gather", C.c_void_p * MAX_PEERS),
        ("ctrl_multicast", C.c_void_p),
    ]


class TileEntry(C.Structure):
    _fields_ = [("array", C.c_int32), ("i", C.c_int32), ("j", C.c_int32)]


class TilesIo(C.Structure):
    _fields_ = [
        ("tiles", C.c_void_p), ("target_vel", C.c_void_p),
        ("u_all", C.c_void_p), ("ctrl", C.c_void_p), ("status", C.c_void_p),
        ("n_gather", C.c_int32), ("reserved_", C.c_int32), ("gather_offset", C.c_int64),
        ("ctrl_gather", C.c_void_p * MAX_PEERS),
        ("ctrl_multicast", C.c_void_p),
    ]


TILE = 32
(ARR_PAD, ARR_M, ARR_J, ARR_DQ, ARR_BIAS, ARR_EE_XYZ, ARR_EE_QUAT, ARR_T_XYZ, ARR_T_QUAT, ARR_MAX_VEL, ARR_FT_XMAT,
 ARR_FT_RAW) = range(12)


class JointModel(C.Structure):
    _fields_ = [
        ("parent", C.c_int32), ("reserved_", C.c_int32),
        ("pos", C.c_double * 3), ("quat", C.c_double * 4), ("axis", C.c_double * 3),
        ("mass", C.c_double), ("com", C.c_double * 3), ("inertia", C.c_double * 6),
    ]


class FrameModel(C.Structure):
    _fields_ = [("joint", C.c_int32), ("reserved_", C.c_int32), ("pos", C.c_double * 3), ("quat", C.c_double * 4)]


class Model(C.Structure):
    _fields_ = [
        ("n_joints", C.c_int32), ("reserved_", C.c_int32),
        ("gravity", C.c_double * 3),
        ("joint", JointModel * MAX_N),
        ("ee", FrameModel * MAX_DEVICES),
        ("ft", FrameModel * MAX_DEVICES),
    ]


class FusedIo(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("dq", C.c_void_p),
        ("target_xyz", C.c_void_p), ("target_quat", C.c_void_p),
        ("target_vel", C.c_void_p), ("max_vel", C.c_void_p), ("ft_raw", C.c_void_p),
        ("ctrl", C.c_void_p), ("u_all", C.c_void_p), ("status", C.c_void_p),
        ("ee_xyz", C.c_void_p), ("ee_quat", C.c_void_p),
    ]


class Action(C.Structure):
    _fields_ = [("type", C.c_int32), ("grip_steps", C.c_int32), ("kp", C.c_double), ("max_error", C.c_double),
                ("min_speed_xyz", C.c_double), ("max_speed_xyz", C.c_double), ("gripper_force", C.c_double)]


class Sequence(C.Structure):
    _fields_ = [("n_actions", C.c_int32), ("active_device", C.c_int32), ("gripper_slot", C.c_int32),
                ("reserved_", C.c_int32), ("passive_quat", C.c_double * 4), ("action", Action * MAX_ACTIONS)]


class SequenceIo(C.Structure):
    _fields_ = [("wp_xyz", C.c_void_p), ("wp_quat", C.c_void_p), ("action", C.c_void_p), ("entered", C.c_void_p),
                ("timer", C.c_void_p), ("err", C.c_void_p), ("max_vel0", C.c_void_p),
                ("target_xyz", C.c_void_p), ("target_quat", C.c_void_p)]


class WaypointsIo(C.Structure):
    _fields_ = [("wps", C.c_void_p), ("W", C.c_int32), ("n_wp", C.c_int32 * MAX_DEVICES), ("reserved_", C.c_int32),
                ("threshold", C.c_double), ("wp_idx", C.c_void_p), ("target_xyz", C.c_void_p), ("target_quat", C.c_void_p)]


EXPORTS = [
    "irlosc_set_model", "irlosc_step_fused", "irlosc_step_fused_host", "irlosc_step_sequence", "irlosc_step_waypoints",
    "irlosc_last_error", "irlosc_abi_version", "irlosc_create", "irlosc_destroy",
    "irlosc_num_task_rows", "irlosc_num_ctrl", "irlosc_step", "irlosc_step_host",
    "irlosc_calc_error", "irlosc_host_alloc", "irlosc_host_free", "irlosc_set_kernel", "irlosc_set_tile_kernel", "irlosc_set_sm_margin",
    "irlosc_kernel_launches", "irlosc_last_kernel",
    "irlosc_tile_entries", "irlosc_tile_spec", "irlosc_tiles_doubles", "irlosc_pack_tiles", "irlosc_pack_tiles_host",
    "irlosc_step_tiles", "irlosc_step_tiles_host",
]

_lib: Optional[C.CDLL] = None


def library_path() -> str:
    # IRLOSC_LIB: an alternative build of the same library (A/B experiments of kernel variants only)
    return os.environ.get("IRLOSC_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libirlosc.so")


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.isfile(path):
        raise NativeLibraryError(
            "%s not found - the CUDA library is the only compute path of irl_control_b200; "
            "build it with `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    try:
        lib = C.CDLL(path)
    except OSError as exc:  # missing libcudart etc.
        raise NativeLibraryError("cannot load %s: %s" % (path, exc)) from exc
    lib.irlosc_last_error.restype = C.c_char_p
    lib.irlosc_last_error.argtypes = []
    lib.irlosc_abi_version.restype = C.c_int32
    lib.irlosc_create.restype = C.c_int32
    lib.irlosc_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
    lib.irlosc_destroy.restype = C.c_int32
    lib.irlosc_destroy.argtypes = [C.c_void_p]
    lib.irlosc_num_task_rows.restype = C.c_int32
    lib.irlosc_num_task_rows.argtypes = [C.c_void_p]
    lib.irlosc_num_ctrl.restype = C.c_int32
    lib.irlosc_num_ctrl.argtypes = [C.c_void_p]
    lib.irlosc_step.restype = C.c_int32
    lib.irlosc_step.argtypes = [C.c_void_p, C.c_int64, C.POINTER(Io), C.c_void_p]
    lib.irlosc_step_host.restype = C.c_int32
    lib.irlosc_step_host.argtypes = [C.c_void_p, C.c_int64, C.POINTER(Io)]
    lib.irlosc_set_model.restype = C.c_int32
    lib.irlosc_set_model.argtypes = [C.c_void_p, C.POINTER(Model)]
    lib.irlosc_step_fused.restype = C.c_int32
    lib.irlosc_step_fused.argtypes = [C.c_void_p, C.c_int64, C.POINTER(FusedIo), C.c_void_p]
    lib.irlosc_step_fused_host.restype = C.c_int32
    lib.irlosc_step_fused_host.argtypes = [C.c_void_p, C.c_int64, C.POINTER(FusedIo)]
    lib.irlosc_step_sequence.restype = C.c_int32
    lib.irlosc_step_sequence.argtypes = [C.c_void_p, C.c_int64, C.POINTER(FusedIo), C.POINTER(Sequence),
                                         C.POINTER(SequenceIo), C.c_void_p]
    lib.irlosc_step_waypoints.restype = C.c_int32
    lib.irlosc_step_waypoints.argtypes = [C.c_void_p, C.c_int64, C.POINTER(FusedIo), C.POINTER(WaypointsIo), C.c_void_p]
    lib.irlosc_calc_error.restype = C.c_int32
    lib.irlosc_calc_error.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.c_void_p]
    lib.irlosc_host_alloc.restype = C.c_int32
    lib.irlosc_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_int64]
    lib.irlosc_host_free.restype = C.c_int32
    lib.irlosc_host_free.argtypes = [C.c_void_p]
    lib.irlosc_set_kernel.restype = C.c_int32
    lib.irlosc_set_kernel.argtypes = [C.c_void_p, C.c_int32]
    lib.irlosc_set_tile_kernel.restype = C.c_int32
    lib.irlosc_set_tile_kernel.argtypes = [C.c_void_p, C.c_int32]
    lib.irlosc_set_sm_margin.restype = C.c_int32
    lib.irlosc_set_sm_margin.argtypes = [C.c_void_p, C.c_int32]
    lib.irlosc_kernel_launches.restype = C.c_int64
    lib.irlosc_kernel_launches.argtypes = [C.c_void_p]
    lib.irlosc_last_kernel.restype = C.c_char_p
    lib.irlosc_last_kernel.argtypes = [C.c_void_p]
    lib.irlosc_tile_entries.restype = C.c_int32
    lib.irlosc_tile_entries.argtypes = [C.c_void_p]
    lib.irlosc_tile_spec.restype = C.c_int32
    lib.irlosc_tile_spec.argtypes = [C.c_void_p, C.POINTER(TileEntry), C.c_int32]
    lib.irlosc_tiles_doubles.restype = C.c_int64
    lib.irlosc_tiles_doubles.argtypes = [C.c_void_p, C.c_int64]
    lib.irlosc_pack_tiles.restype = C.c_int32
    lib.irlosc_pack_tiles.argtypes = [C.c_void_p, C.c_int64, C.POINTER(Io), C.c_void_p, C.c_void_p]
    lib.irlosc_pack_tiles_host.restype = C.c_int32
    lib.irlosc_pack_tiles_host.argtypes = [C.c_void_p, C.c_int64, C.POINTER(Io), C.c_void_p]
    lib.irlosc_step_tiles.restype = C.c_int32
    lib.irlosc_step_tiles.argtypes = [C.c_void_p, C.c_int64, C.POINTER(TilesIo), C.c_void_p]
    lib.irlosc_step_tiles_host.restype = C.c_int32
    lib.irlosc_step_tiles_host.argtypes = [C.c_void_p, C.c_int64, C.POINTER(TilesIo)]
    if lib.irlosc_abi_version() != ABI_VERSION:
        raise NativeLibraryError("libirlosc ABI %d != binding %d" % (lib.irlosc_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        raise OscError(code, load().irlosc_last_error().decode("utf-8", "replace"))

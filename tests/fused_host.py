"""TEST INFRASTRUCTURE: builds and drives tests/host_fused/fused_host.cu, which runs the
per-instance function of the fused CUDA kernel (osc_fused.cuh, __host__ __device__) on the CPU.
Never imported by the package."""
import ctypes as C
import os
import subprocess

import numpy as np

from irl_control_b200 import _native

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_fused", "fused_host.cu")
LIB = os.path.join(HERE, "_build", "libfused_host.so")
CSRC = os.path.join(os.path.dirname(HERE), "irl_control_b200", "csrc")
_lib = None


def build():
    deps = [SRC, os.path.join(os.path.dirname(HERE), "include", "irlosc.h")] + [
        os.path.join(CSRC, f) for f in ("osc_fused.cuh", "osc_fused_types.h", "irlosc_device.cuh", "osc_tail.cuh",
                                        "osc_stream.cuh", "osc_lane.cuh", "osc_eigen.cuh", "osc_sequence.cuh", "irlosc_build.h",
                                        "irlosc_internal.h")]
    if os.path.isfile(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(p) for p in deps):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [os.environ.get("NVCC", "nvcc"), "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
           "-diag-suppress", "20014,20011", "-shared", "-Xcompiler", "-fPIC", "-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.fused_host_run.restype = C.c_int64
        _lib.fused_host_run.argtypes = [C.POINTER(_native.Params), C.POINTER(_native.Model), C.c_int64,
                                        C.POINTER(_native.FusedIo)] + [C.c_void_p] * 6
        _lib.fused_host_error.restype = C.c_char_p
        _lib.stream_host_run.restype = C.c_int64
        _lib.stream_host_run.argtypes = [C.POINTER(_native.Params), C.c_int64, C.POINTER(_native.Io)] + [C.c_void_p] * 6
        _lib.host_set_how.restype = None
        _lib.host_set_how.argtypes = [C.c_void_p]
    return _lib


def run(layout, model, inp, debug=False):
    """inp: dict of numpy arrays q, dq, target_xyz, target_quat [, target_vel, max_vel, ft_raw]."""
    lib = load()
    B = int(inp["q"].shape[0])
    n, D, k, nc = layout.n, layout.D, layout.k, layout.n_ctrl
    keep = {k_: np.ascontiguousarray(v, dtype=np.float64) for k_, v in inp.items()}
    out = {"ctrl": np.zeros((B, nc)), "u_all": np.zeros((B, n)), "status": np.zeros(B, dtype=np.uint8),
           "ee_xyz": np.zeros((B, D, 3)), "ee_quat": np.zeros((B, D, 4))}
    io = _native.FusedIo()
    for name in ("q", "dq", "target_xyz", "target_quat", "target_vel", "max_vel", "ft_raw"):
        setattr(io, name, keep[name].ctypes.data if name in keep else None)
    for name in out:
        setattr(io, name, out[name].ctypes.data)
    dbg = {}
    ptrs = [None] * 6
    if debug:
        dbg = {"A": np.zeros((B, k, k)), "g": np.zeros((B, k)), "uv": np.zeros((B, n)),
               "bias": np.zeros((B, n)), "dx": np.zeros((B, k)), "J": np.zeros((B, k, n))}
        ptrs = [dbg[x].ctypes.data for x in ("A", "g", "uv", "bias", "dx", "J")]
    params = layout.to_c_params()
    how = np.zeros(B, dtype=np.int32)
    lib.host_set_how(how.ctypes.data)
    try:
        rc = lib.fused_host_run(C.byref(params), C.byref(model), B, C.byref(io), *ptrs)
    finally:
        lib.host_set_how(None)
    if rc < 0:
        raise RuntimeError(lib.fused_host_error().decode())
    out["n_hard"] = int(rc)
    out["how"], out["how_detail"] = how & 0xf, how
    out.update(dbg)
    return out


def run_stream(layout, state, debug=False, strides=None):
    """The streaming step (osc_stream.cuh) on the CPU.  state: numpy arrays in `BatchedOSC.step` field names."""
    lib = load()
    keep = {k_: np.ascontiguousarray(v, dtype=np.float64) if k_ not in (strides or {}).get("views", ()) else v
            for k_, v in state.items()}
    B = int(keep["dq"].shape[0])
    n, D, k, nc = layout.n, layout.D, layout.k, layout.n_ctrl
    out = {"ctrl": np.zeros((B, nc)), "u_all": np.zeros((B, n)), "status": np.zeros(B, dtype=np.uint8)}
    io = _native.Io()
    if "qM" in keep:                       # MuJoCo's sparse inertia (IRLOSC_M_QM), [B][>= nM]
        keep["M"] = keep.pop("qM")
        io.m_layout = _native.M_QM
        strides = dict(strides or {})
        strides.setdefault("m_stride", int(keep["M"].shape[1]))
    else:
        io.m_layout = _native.M_DENSE if keep["M"].ndim == 3 else _native.M_PACKED
    io.j_layout = _native.J_FULL6 if keep["J"].ndim == 4 else _native.J_ROWS
    for name in ("M", "J", "dq", "bias", "ee_xyz", "ee_quat", "target_xyz", "target_quat", "target_vel", "max_vel",
                 "ft_xmat", "ft_raw"):
        setattr(io, name, keep[name].ctypes.data if name in keep else None)
    for name in out:
        setattr(io, name, out[name].ctypes.data)
    for key in ("ldm", "m_stride", "ldj", "j_stride"):
        setattr(io, key, int((strides or {}).get(key, 0)))
    dbg = {}
    ptrs = [None] * 5
    if debug:
        dbg = {"A": np.zeros((B, k, k)), "g": np.zeros((B, k)), "uv": np.zeros((B, n)), "dx": np.zeros((B, k)),
               "J": np.zeros((B, k, n))}
        ptrs = [dbg[x].ctypes.data for x in ("A", "g", "uv", "dx", "J")]
    params = layout.to_c_params()
    nch = C.c_int32(0)
    how = np.zeros(B, dtype=np.int32)
    lib.host_set_how(how.ctypes.data)
    try:
        rc = lib.stream_host_run(C.byref(params), B, C.byref(io), *ptrs, C.byref(nch))
    finally:
        lib.host_set_how(None)
    if rc < 0:
        raise RuntimeError(lib.fused_host_error().decode())
    out["n_hard"] = int(rc)
    out["how"], out["how_detail"] = how & 0xf, how
    out["n_chunks"] = int(nch.value)
    out.update(dbg)
    return out


def sequence_step(layout, model, seq, inp, seq_state):
    """One `irlosc_step_sequence` on the CPU (numpy arrays; seq_state is updated in place)."""
    lib = load()
    lib.sequence_host_step.restype = C.c_int64
    lib.sequence_host_step.argtypes = [C.POINTER(_native.Params), C.POINTER(_native.Model), C.c_int64,
                                       C.POINTER(_native.FusedIo), C.POINTER(_native.Sequence), C.POINTER(_native.SequenceIo)]
    B = int(inp["q"].shape[0])
    keep = {k_: np.ascontiguousarray(v, dtype=np.float64) for k_, v in inp.items()}
    out = {"ctrl": np.zeros((B, layout.n_ctrl)), "u_all": np.zeros((B, layout.n)), "status": np.zeros(B, dtype=np.uint8),
           "ee_xyz": np.zeros((B, layout.D, 3)), "ee_quat": np.zeros((B, layout.D, 4))}
    io = _native.FusedIo()
    for name in ("q", "dq", "target_vel", "max_vel", "ft_raw"):
        setattr(io, name, keep[name].ctypes.data if name in keep else None)
    for name in out:
        setattr(io, name, out[name].ctypes.data)
    sio = _native.SequenceIo()
    for name in ("wp_xyz", "wp_quat", "action", "entered", "timer", "err", "max_vel0", "target_xyz", "target_quat"):
        setattr(sio, name, seq_state[name].ctypes.data)
    params = layout.to_c_params()
    rc = lib.sequence_host_step(C.byref(params), C.byref(model), B, C.byref(io), C.byref(seq.c_struct), C.byref(sio))
    if rc < 0:
        raise RuntimeError(lib.fused_host_error().decode())
    return out


def waypoints_step(layout, model, inp, wp_state, threshold=0.1):
    """One `irlosc_step_waypoints` on the CPU (numpy arrays; wp_state is updated in place)."""
    lib = load()
    lib.waypoints_host_step.restype = C.c_int64
    lib.waypoints_host_step.argtypes = [C.POINTER(_native.Params), C.POINTER(_native.Model), C.c_int64,
                                        C.POINTER(_native.FusedIo), C.POINTER(_native.WaypointsIo)]
    B = int(inp["q"].shape[0])
    keep = {k_: np.ascontiguousarray(v, dtype=np.float64) for k_, v in inp.items()}
    out = {"ctrl": np.zeros((B, layout.n_ctrl)), "u_all": np.zeros((B, layout.n)), "status": np.zeros(B, dtype=np.uint8)}
    io = _native.FusedIo()
    for name in ("q", "dq", "target_vel", "max_vel", "ft_raw"):
        setattr(io, name, keep[name].ctypes.data if name in keep else None)
    for name in out:
        setattr(io, name, out[name].ctypes.data)
    wio = _native.WaypointsIo()
    wio.wps, wio.W, wio.threshold = wp_state["wps"].ctypes.data, int(wp_state["wps"].shape[2]), float(threshold)
    for d, n in enumerate(wp_state["n_wp"]):
        wio.n_wp[d] = int(n)
    wio.wp_idx = wp_state["wp_idx"].ctypes.data
    wio.target_xyz, wio.target_quat = wp_state["target_xyz"].ctypes.data, wp_state["target_quat"].ctypes.data
    params = layout.to_c_params()
    rc = lib.waypoints_host_step(C.byref(params), C.byref(model), B, C.byref(io), C.byref(wio))
    if rc < 0:
        raise RuntimeError(lib.fused_host_error().decode())
    return out


class TileEntry(C.Structure):
    _fields_ = [("array", C.c_int32), ("i", C.c_int32), ("j", C.c_int32)]


def lane_spec(layout):
    """Entry table of the tile layout (csrc/osc_lane.cuh build_tile_spec): (entries [(array, i, j)], gbase)."""
    lib = load()
    lib.lane_host_spec.restype = C.c_int32
    lib.lane_host_spec.argtypes = [C.POINTER(_native.Params), C.POINTER(TileEntry), C.c_int32, C.POINTER(C.c_int32)]
    params = layout.to_c_params()
    buf = (TileEntry * 512)()
    gbase = (C.c_int32 * 12)()
    E = lib.lane_host_spec(C.byref(params), buf, 512, gbase)
    return [(buf[e].array, buf[e].i, buf[e].j) for e in range(max(E, 0))], list(gbase)


def run_lane(layout, state):
    """The lane step (osc_lane.cuh) on the CPU: arrays are packed into tiles with the product's pack table, then the
    kernel's per-instance function reads them.  state: numpy arrays in `BatchedOSC.step` field names."""
    lib = load()
    lib.lane_host_run.restype = C.c_int64
    lib.lane_host_run.argtypes = [C.POINTER(_native.Params), C.c_int64, C.POINTER(_native.Io), C.c_void_p]
    keep = {k_: np.ascontiguousarray(v, dtype=np.float64) for k_, v in state.items()}
    B = int(keep["dq"].shape[0])
    out = {"ctrl": np.zeros((B, layout.n_ctrl)), "u_all": np.zeros((B, layout.n)), "status": np.zeros(B, dtype=np.uint8)}
    io = _native.Io()
    if "qM" in keep:
        keep["M"] = keep.pop("qM")
        io.m_layout = _native.M_QM
        io.m_stride = int(keep["M"].shape[1])
    else:
        io.m_layout = _native.M_DENSE if keep["M"].ndim == 3 else _native.M_PACKED
    io.j_layout = _native.J_FULL6 if keep["J"].ndim == 4 else _native.J_ROWS
    for name in ("M", "J", "dq", "bias", "ee_xyz", "ee_quat", "target_xyz", "target_quat", "target_vel", "max_vel",
                 "ft_xmat", "ft_raw"):
        setattr(io, name, keep[name].ctypes.data if name in keep else None)
    for name in out:
        setattr(io, name, out[name].ctypes.data)
    E = len(lane_spec(layout)[0])
    tiles = np.full(((B + 31) // 32, E, 32), np.nan)
    params = layout.to_c_params()
    how = np.zeros(B, dtype=np.int32)
    lib.host_set_how(how.ctypes.data)
    try:
        rc = lib.lane_host_run(C.byref(params), B, C.byref(io), tiles.ctypes.data)
    finally:
        lib.host_set_how(None)
    if rc < 0:
        raise RuntimeError(lib.fused_host_error().decode())
    out["n_hard"], out["how"], out["how_detail"], out["tiles"] = int(rc), how & 0xf, how, tiles
    return out


# TailHow bits of csrc/osc_tail.cuh (out["how"]); out["how_detail"] keeps the statistics above them: why an instance
# went to the warp (WHY_*), inertia counts (bits 12-19), eigenvector iterations (20-27), refinement steps (28-30)
HOW_INVERSE, HOW_CUT1, HOW_CUT2, HOW_WARP = 1, 2, 4, 8
WHY = {16: "blocks", 32: "base_zero", 64: "lost", 128: "undecided", 256: "many", 512: "no_convergence", 1024: "residual"}

"""`BatchedOSC` - Python face of one `irlosc_handle` (include/irlosc.h).

Evaluates the reference's `OSC.generate` control law (osc.py:120-210) for B
independent robot instances per call:

    step_tiles(...) state in GPU memory in the native batch-interleaved tile layout
                    (`pack_tiles` converts from per-variable tensors): the lane kernel
    step(...)       state already in GPU memory as per-variable tensors (torch CUDA, float64);
                    asynchronous on the current torch stream, returns tensors
    step_host(...)  state in host memory (numpy arrays); host->device copies,
                    kernel and device->host copies are pipelined in chunks
                    inside the library; returns numpy arrays

Field names follow the reference's state enums: M, J, DQ (RobotState,
robot.py:11-14), EE_XYZ, EE_QUAT (DeviceState, device.py:7-18), qfrc_bias
(osc.py:191), plus the Target fields (utils.py:5-67).  Per-device arrays are
indexed in TARGET order.

torch is used for device memory and streams only; all arithmetic happens in
the CUDA library.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _native
from .layout import OscLayout, qm_size

_OPTIONAL = ("bias", "target_vel", "max_vel", "ft_xmat", "ft_raw")
_FIELDS = ("M", "J", "dq", "bias", "ee_xyz", "ee_quat", "target_xyz", "target_quat",
           "target_vel", "max_vel", "ft_xmat", "ft_raw")


class BatchedOSC:
    def __init__(self, layout: OscLayout, device: Optional[int] = None):
        self.layout = layout
        self.lib = _native.load()
        self._handle = C.c_void_p()
        self._torch_device = None
        if device is not None:
            import torch
            torch.cuda.set_device(device)
            self._torch_device = torch.device("cuda", device)
        params = layout.to_c_params()
        _native.check(self.lib.irlosc_create(C.byref(params), C.byref(self._handle)))
        self.n = layout.n
        self.D = layout.D
        self.k = self.lib.irlosc_num_task_rows(self._handle)
        self.n_ctrl = self.lib.irlosc_num_ctrl(self._handle)
        assert self.k == layout.k and self.n_ctrl == layout.n_ctrl

    # ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self.lib.irlosc_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_kernel(self, which: int):
        _native.check(self.lib.irlosc_set_kernel(self._handle, int(which)))

    TILE_KERNELS = {"auto": 0, "lane": 1, "pair": 2}

    def set_tile_kernel(self, which):
        """Kernel of `step_tiles` / `step_tiles_host`: "auto", "lane" (a thread per instance) or "pair" (a lane per arm)."""
        _native.check(self.lib.irlosc_set_tile_kernel(self._handle, self.TILE_KERNELS.get(which, which)))

    def set_sm_margin(self, sms: int):
        """Leave `sms` SMs free so a collective on another stream can overlap the step kernel."""
        _native.check(self.lib.irlosc_set_sm_margin(self._handle, int(sms)))

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.irlosc_kernel_launches(self._handle))

    @property
    def last_kernel(self) -> str:
        return self.lib.irlosc_last_kernel(self._handle).decode()

    # ------------------------------------------------------------------
    def _shapes(self, B: int, m_layout: int, j_layout: int) -> Dict[str, tuple]:
        n, D, k = self.n, self.D, self.k
        return {
            "M": (B, n, n) if m_layout == _native.M_DENSE else (B, n * (n + 1) // 2),
            "J": (B, k, n) if j_layout == _native.J_ROWS else (B, D, 6, n),
            "dq": (B, n), "bias": (B, n),
            "ee_xyz": (B, D, 3), "ee_quat": (B, D, 4),
            "target_xyz": (B, D, 3), "target_quat": (B, D, 4),
            "target_vel": (B, D, 6), "max_vel": (B, D, 2),
            "ft_xmat": (B, D, 9), "ft_raw": (B, D, 6),
        }

    def _infer_layouts(self, state):
        M, J = state["M"], state["J"]
        m_layout = _native.M_DENSE if len(M.shape) == 3 else _native.M_PACKED
        if state.get("_qM"):
            m_layout = _native.M_QM
        j_layout = _native.J_FULL6 if len(J.shape) == 4 else _native.J_ROWS
        return m_layout, j_layout

    def _accept_qM(self, state, shapes_out=None):
        """`state["qM"]` ([B, nM of the scene], MuJoCo's sparse `mjData.qM` - the array the reference expands
        with `mj_fullM`, robot.py:69) may stand in for `state["M"]` (IRLOSC_M_QM)."""
        if "qM" not in state:
            return state
        if "M" in state:
            raise ValueError("give either state['M'] or state['qM'], not both")
        qM = state["qM"]
        nM = qm_size(self.layout.joint_parent) if self.layout.joint_parent is not None else None
        if nM is None:
            raise ValueError("state['qM'] needs a layout with the kinematic tree (joint_parent)")
        if len(qM.shape) != 2 or int(qM.shape[1]) < nM:
            raise ValueError("state['qM'] has shape %s, expected (B, >= %d)" % (tuple(qM.shape), nM))
        state = {k: v for k, v in state.items() if k != "qM"}
        state["M"] = qM
        state["_qM"] = True
        return state

    def _check(self, state, B, shapes, is_ok):
        for name in _FIELDS:
            arr = state.get(name)
            if arr is None:
                if name in _OPTIONAL:
                    continue
                raise ValueError("state['%s'] is required" % name)
            want = shapes[name]
            got = tuple(arr.shape)
            if name == "ft_xmat" and got == (B, self.D, 3, 3):
                got = want
            if got != want:
                raise ValueError("state['%s'] has shape %s, expected %s" % (name, got, want))
            is_ok(name, arr)
        if self.layout.use_g and state.get("bias") is None:
            raise ValueError("use_g is set: state['bias'] (qfrc_bias) is required")
        if self.layout.admittance and (state.get("ft_xmat") is None or state.get("ft_raw") is None):
            raise ValueError("admittance is set: state['ft_xmat'] and state['ft_raw'] are required")

    # ------------------------------------------------------------------
    def step(self, state: Dict, out: Optional[Dict] = None, want_u_all: bool = False,
             want_status: bool = True, strides: Optional[Dict[str, int]] = None,
             gather: Optional[tuple] = None) -> Dict:
        """One control step for B instances resident on the GPU.

        state   : dict of CUDA float64 contiguous tensors (see module docstring)
        out     : optional preallocated {"ctrl", "u_all", "status"} tensors
        strides : optional {"ldm", "m_stride", "ldj", "j_stride"} in doubles when M / J are views into
                  larger buffers, e.g. the robot block of the scene's nv x nv `mj_fullM` output
                  (robot.py:69-71) - `M` is then `[B, nv, nv]` and ldm = nv, m_stride = nv * nv.
        gather  : optional (peer_ptrs, row_offset): device pointers of every rank's gathered
                  `[B_total, n_ctrl]` array (peer-mapped, e.g. torch symmetric memory `buffer_ptrs`);
                  the kernel then also stores its ctrl rows at row_offset + i of each of them
                  (fused NVLink gather, see `irlosc_io.ctrl_gather`).  A third element, the
                  multicast (NVLS) address of the gathered array (symmetric memory
                  `multicast_ptr`), makes the kernel write each row once through the switch
                  instead of once per peer (`irlosc_io.ctrl_multicast`).
        """
        import torch
        state = self._accept_qM(state)
        M = state["M"]
        if not M.is_cuda:
            raise ValueError("BatchedOSC.step expects CUDA tensors; use step_host for host arrays")
        B = int(M.shape[0])
        m_layout, j_layout = self._infer_layouts(state)
        shapes = self._shapes(B, m_layout, j_layout)
        strides = dict(strides or {})
        if m_layout == _native.M_QM:
            shapes["M"] = tuple(M.shape)
            strides["m_stride"] = int(M.shape[1])
        if "ldm" in strides:
            shapes["M"] = tuple(M.shape)            # a view: the caller vouches for ldm / m_stride
        if "ldj" in strides:
            shapes["J"] = tuple(state["J"].shape)

        def ok(name, t):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.device == M.device):
                raise ValueError("state['%s'] must be a contiguous float64 CUDA tensor on %s" % (name, M.device))
        self._check(state, B, shapes, ok)
        out = {} if out is None else out
        if "ctrl" not in out:
            out["ctrl"] = torch.empty(B, self.n_ctrl, dtype=torch.float64, device=M.device)
        if want_u_all and "u_all" not in out:
            out["u_all"] = torch.empty(B, self.n, dtype=torch.float64, device=M.device)
        if want_status and "status" not in out:
            out["status"] = torch.empty(B, dtype=torch.uint8, device=M.device)
        io = _native.Io()
        io.m_layout, io.j_layout = m_layout, j_layout
        for name in _FIELDS:
            t = state.get(name)
            setattr(io, name, t.data_ptr() if t is not None else None)
        io.ctrl = out["ctrl"].data_ptr()
        io.u_all = out["u_all"].data_ptr() if "u_all" in out else None
        io.status = out["status"].data_ptr() if "status" in out else None
        if gather is not None:
            ptrs, offset = gather[0], gather[1]
            if len(gather) > 2 and gather[2]:
                io.ctrl_multicast = int(gather[2])
            if len(ptrs) > _native.MAX_PEERS:
                raise ValueError("at most %d peers" % _native.MAX_PEERS)
            io.n_gather, io.gather_offset = len(ptrs), int(offset)
            for gi, ptr in enumerate(ptrs):
                io.ctrl_gather[gi] = int(ptr)
        io.ldm, io.m_stride = int(strides.get("ldm", 0)), int(strides.get("m_stride", 0))
        io.ldj, io.j_stride = int(strides.get("ldj", 0)), int(strides.get("j_stride", 0))
        stream = torch.cuda.current_stream(M.device).cuda_stream
        with torch.cuda.device(M.device):
            _native.check(self.lib.irlosc_step(self._handle, B, C.byref(io), C.c_void_p(stream)))
        return out

    def step_host(self, state: Dict, out: Optional[Dict] = None, want_u_all: bool = False,
                  want_status: bool = True) -> Dict:
        """Same step with HOST numpy arrays (float64, C-contiguous); blocks until results are valid."""
        state = self._accept_qM(state)
        M = state["M"]
        B = int(M.shape[0])
        m_layout, j_layout = self._infer_layouts(state)
        shapes = self._shapes(B, m_layout, j_layout)
        if m_layout == _native.M_QM:
            shapes["M"] = tuple(M.shape)

        def ok(name, a):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
                raise ValueError("state['%s'] must be a C-contiguous float64 numpy array" % name)
        self._check(state, B, shapes, ok)
        out = {} if out is None else out
        if "ctrl" not in out:
            out["ctrl"] = np.empty((B, self.n_ctrl), dtype=np.float64)
        if want_u_all and "u_all" not in out:
            out["u_all"] = np.empty((B, self.n), dtype=np.float64)
        if want_status and "status" not in out:
            out["status"] = np.empty((B,), dtype=np.uint8)
        io = _native.Io()
        io.m_layout, io.j_layout = m_layout, j_layout
        if m_layout == _native.M_QM:
            io.m_stride = int(M.shape[1])
        for name in _FIELDS:
            a = state.get(name)
            setattr(io, name, a.ctypes.data if a is not None else None)
        io.ctrl = out["ctrl"].ctypes.data
        io.u_all = out["u_all"].ctypes.data if "u_all" in out else None
        io.status = out["status"].ctypes.data if "status" in out else None
        if self._torch_device is not None:
            import torch
            with torch.cuda.device(self._torch_device):
                _native.check(self.lib.irlosc_step_host(self._handle, B, C.byref(io)))
        else:
            _native.check(self.lib.irlosc_step_host(self._handle, B, C.byref(io)))
        return out

    # ------------------------------------------------------------------ batch-interleaved tiles (native layout)
    @property
    def tile_entries(self) -> int:
        """E: doubles per instance in the tile layout (0: this controller has none)."""
        return int(self.lib.irlosc_tile_entries(self._handle))

    def tile_spec(self):
        """[(array, i, j)] for every tile entry (`irlosc_tile_spec`, IRLOSC_ARR_* ids in `_native`)."""
        E = self.tile_entries
        buf = (_native.TileEntry * max(E, 1))()
        self.lib.irlosc_tile_spec(self._handle, buf, E)
        return [(buf[e].array, buf[e].i, buf[e].j) for e in range(E)]

    def tiles_shape(self, B: int) -> tuple:
        return ((B + _native.TILE - 1) // _native.TILE, self.tile_entries, _native.TILE)

    def _io_inputs(self, state, B, on_device):
        """irlosc_io with the input pointers of `state` (same checks as step / step_host)."""
        import torch
        state = self._accept_qM(state)
        M = state["M"]
        m_layout, j_layout = self._infer_layouts(state)
        shapes = self._shapes(B, m_layout, j_layout)
        if m_layout == _native.M_QM:
            shapes["M"] = tuple(M.shape)

        def ok_dev(name, t):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.device == M.device):
                raise ValueError("state['%s'] must be a contiguous float64 CUDA tensor on %s" % (name, M.device))

        def ok_host(name, a):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
                raise ValueError("state['%s'] must be a C-contiguous float64 numpy array" % name)
        self._check(state, B, shapes, ok_dev if on_device else ok_host)
        io = _native.Io()
        io.m_layout, io.j_layout = m_layout, j_layout
        if m_layout == _native.M_QM:
            io.m_stride = int(M.shape[1])
        for name in _FIELDS:
            t = state.get(name)
            setattr(io, name, (t.data_ptr() if on_device else t.ctypes.data) if t is not None else None)
        return io

    def pack_tiles(self, state: Dict, tiles=None):
        """Per-variable CUDA tensors (the `step` fields, any M / J layout) -> tiles `[ceil(B/32), E, 32]` on the GPU
        (`irlosc_pack_tiles`, asynchronous on the current stream)."""
        import torch
        B = int(state["dq"].shape[0])
        dev = state["dq"].device
        if tiles is None:
            tiles = torch.empty(self.tiles_shape(B), dtype=torch.float64, device=dev)
        io = self._io_inputs(state, B, True)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _native.check(self.lib.irlosc_pack_tiles(self._handle, B, C.byref(io), C.c_void_p(tiles.data_ptr()), C.c_void_p(stream)))
        return tiles

    def pack_tiles_host(self, state: Dict, tiles: Optional[np.ndarray] = None) -> np.ndarray:
        """Same conversion between host numpy arrays (`irlosc_pack_tiles_host`; a data-layout helper)."""
        B = int(state["dq"].shape[0])
        if tiles is None:
            tiles = np.empty(self.tiles_shape(B), dtype=np.float64)
        io = self._io_inputs(state, B, False)
        _native.check(self.lib.irlosc_pack_tiles_host(self._handle, B, C.byref(io), C.c_void_p(tiles.ctypes.data)))
        return tiles

    def step_tiles(self, tiles, B: int, out: Optional[Dict] = None, want_u_all: bool = False, want_status: bool = True,
                   target_vel=None, gather: Optional[tuple] = None) -> Dict:
        """One control step for B instances stored as batch-interleaved tiles on the GPU (`irlosc_step_tiles`): one
        launch of the lane kernel, asynchronous on the current torch stream.  `gather` as in `step`."""
        import torch
        if not (tiles.is_cuda and tiles.dtype == torch.float64 and tiles.is_contiguous()):
            raise ValueError("tiles must be a contiguous float64 CUDA tensor")
        if tuple(tiles.shape) != self.tiles_shape(B):
            raise ValueError("tiles has shape %s, expected %s" % (tuple(tiles.shape), self.tiles_shape(B)))
        dev = tiles.device
        out = {} if out is None else out
        if "ctrl" not in out:
            out["ctrl"] = torch.empty(B, self.n_ctrl, dtype=torch.float64, device=dev)
        if want_u_all and "u_all" not in out:
            out["u_all"] = torch.empty(B, self.n, dtype=torch.float64, device=dev)
        if want_status and "status" not in out:
            out["status"] = torch.empty(B, dtype=torch.uint8, device=dev)
        io = _native.TilesIo()
        io.tiles = tiles.data_ptr()
        io.target_vel = target_vel.data_ptr() if target_vel is not None else None
        io.ctrl = out["ctrl"].data_ptr()
        io.u_all = out["u_all"].data_ptr() if "u_all" in out else None
        io.status = out["status"].data_ptr() if "status" in out else None
        if gather is not None:
            ptrs, offset = gather[0], gather[1]
            if len(gather) > 2 and gather[2]:
                io.ctrl_multicast = int(gather[2])
            if len(ptrs) > _native.MAX_PEERS:
                raise ValueError("at most %d peers" % _native.MAX_PEERS)
            io.n_gather, io.gather_offset = len(ptrs), int(offset)
            for gi, ptr in enumerate(ptrs):
                io.ctrl_gather[gi] = int(ptr)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _native.check(self.lib.irlosc_step_tiles(self._handle, B, C.byref(io), C.c_void_p(stream)))
        return out

    def step_tiles_host(self, tiles: np.ndarray, B: int, out: Optional[Dict] = None, want_u_all: bool = False,
                        want_status: bool = True, target_vel: Optional[np.ndarray] = None) -> Dict:
        """`step_tiles` with HOST numpy tiles (pinned memory makes the copies asynchronous); blocks until valid."""
        if not (isinstance(tiles, np.ndarray) and tiles.dtype == np.float64 and tiles.flags["C_CONTIGUOUS"]):
            raise ValueError("tiles must be a C-contiguous float64 numpy array")
        if tuple(tiles.shape) != self.tiles_shape(B):
            raise ValueError("tiles has shape %s, expected %s" % (tuple(tiles.shape), self.tiles_shape(B)))
        out = {} if out is None else out
        if "ctrl" not in out:
            out["ctrl"] = np.empty((B, self.n_ctrl), dtype=np.float64)
        if want_u_all and "u_all" not in out:
            out["u_all"] = np.empty((B, self.n), dtype=np.float64)
        if want_status and "status" not in out:
            out["status"] = np.empty((B,), dtype=np.uint8)
        io = _native.TilesIo()
        io.tiles = tiles.ctypes.data
        io.target_vel = target_vel.ctypes.data if target_vel is not None else None
        io.ctrl = out["ctrl"].ctypes.data
        io.u_all = out["u_all"].ctypes.data if "u_all" in out else None
        io.status = out["status"].ctypes.data if "status" in out else None
        if self._torch_device is not None:
            import torch
            with torch.cuda.device(self._torch_device):
                _native.check(self.lib.irlosc_step_tiles_host(self._handle, B, C.byref(io)))
        else:
            _native.check(self.lib.irlosc_step_tiles_host(self._handle, B, C.byref(io)))
        return out

    # ------------------------------------------------------------------ fused state provider
    _FUSED_FIELDS = ("q", "dq", "target_xyz", "target_quat", "target_vel", "max_vel", "ft_raw")

    def set_model(self, model: "_native.Model"):
        """Attach the rigid-body description (`rigid_model.reduce_model`) the fused step needs."""
        _native.check(self.lib.irlosc_set_model(self._handle, C.byref(model)))
        self._has_model = True

    def _fused_shapes(self, B: int) -> Dict[str, tuple]:
        n, D = self.n, self.D
        return {"q": (B, n), "dq": (B, n), "target_xyz": (B, D, 3), "target_quat": (B, D, 4),
                "target_vel": (B, D, 6), "max_vel": (B, D, 2), "ft_raw": (B, D, 6)}

    def _fused_check(self, state, B, is_ok):
        shapes = self._fused_shapes(B)
        for name in self._FUSED_FIELDS:
            arr = state.get(name)
            if arr is None:
                if name in ("target_vel", "max_vel", "ft_raw"):
                    continue
                raise ValueError("state['%s'] is required" % name)
            if tuple(arr.shape) != shapes[name]:
                raise ValueError("state['%s'] has shape %s, expected %s" % (name, tuple(arr.shape), shapes[name]))
            is_ok(name, arr)
        if self.layout.admittance and state.get("ft_raw") is None:
            raise ValueError("admittance is set: state['ft_raw'] is required")

    def step_fused(self, state: Dict, out: Optional[Dict] = None, want_u_all: bool = False,
                   want_status: bool = True, want_ee: bool = False) -> Dict:
        """One control step from joint states: what `Robot.get_all_states` + `OSC.generate` do together
        (robot.py:125-136, osc.py:120-210), with M, J, qfrc_bias and the EE poses computed on the GPU
        from `state["q"]`, `state["dq"]` (CUDA float64 `[B, n]`) instead of read from a simulator."""
        import torch
        q = state["q"]
        if not q.is_cuda:
            raise ValueError("BatchedOSC.step_fused expects CUDA tensors; use step_fused_host for host arrays")
        B = int(q.shape[0])

        def ok(name, t):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.device == q.device):
                raise ValueError("state['%s'] must be a contiguous float64 CUDA tensor on %s" % (name, q.device))
        self._fused_check(state, B, ok)
        out = {} if out is None else out
        new = lambda *shape, dtype=torch.float64: torch.empty(*shape, dtype=dtype, device=q.device)
        if "ctrl" not in out:
            out["ctrl"] = new(B, self.n_ctrl)
        if want_u_all and "u_all" not in out:
            out["u_all"] = new(B, self.n)
        if want_status and "status" not in out:
            out["status"] = new(B, dtype=torch.uint8)
        if want_ee:
            out.setdefault("ee_xyz", new(B, self.D, 3))
            out.setdefault("ee_quat", new(B, self.D, 4))
        io = _native.FusedIo()
        for name in self._FUSED_FIELDS:
            t = state.get(name)
            setattr(io, name, t.data_ptr() if t is not None else None)
        for name in ("ctrl", "u_all", "status", "ee_xyz", "ee_quat"):
            setattr(io, name, out[name].data_ptr() if name in out else None)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        with torch.cuda.device(q.device):
            _native.check(self.lib.irlosc_step_fused(self._handle, B, C.byref(io), C.c_void_p(stream)))
        return out

    def step_fused_host(self, state: Dict, out: Optional[Dict] = None, want_u_all: bool = False,
                        want_status: bool = True, want_ee: bool = False) -> Dict:
        """`step_fused` with HOST numpy arrays; blocks until the results are valid."""
        q = state["q"]
        B = int(q.shape[0])

        def ok(name, a):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
                raise ValueError("state['%s'] must be a C-contiguous float64 numpy array" % name)
        self._fused_check(state, B, ok)
        out = {} if out is None else out
        if "ctrl" not in out:
            out["ctrl"] = np.empty((B, self.n_ctrl), dtype=np.float64)
        if want_u_all and "u_all" not in out:
            out["u_all"] = np.empty((B, self.n), dtype=np.float64)
        if want_status and "status" not in out:
            out["status"] = np.empty((B,), dtype=np.uint8)
        if want_ee:
            out.setdefault("ee_xyz", np.empty((B, self.D, 3), dtype=np.float64))
            out.setdefault("ee_quat", np.empty((B, self.D, 4), dtype=np.float64))
        io = _native.FusedIo()
        for name in self._FUSED_FIELDS:
            a = state.get(name)
            setattr(io, name, a.ctypes.data if a is not None else None)
        for name in ("ctrl", "u_all", "status", "ee_xyz", "ee_quat"):
            setattr(io, name, out[name].ctypes.data if name in out else None)
        if self._torch_device is not None:
            import torch
            with torch.cuda.device(self._torch_device):
                _native.check(self.lib.irlosc_step_fused_host(self._handle, B, C.byref(io)))
        else:
            _native.check(self.lib.irlosc_step_fused_host(self._handle, B, C.byref(io)))
        return out

    def step_sequence(self, state: Dict, seq, seq_state: Dict, out: Optional[Dict] = None, want_u_all: bool = False,
                      want_status: bool = True) -> Dict:
        """One control step of B episodes of `seq` (`sequence.ActionSequence`): the state machine of
        examples/insertion_task.py:190-204,279-297 runs per instance inside the fused step.  `state`
        holds q, dq (+ optional max_vel, ft_raw); the targets live in `seq_state` (updated in place)."""
        import torch
        q = state["q"]
        B = int(q.shape[0])
        st2 = dict(state, target_xyz=seq_state["target_xyz"], target_quat=seq_state["target_quat"])

        def ok(name, t):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.device == q.device):
                raise ValueError("state['%s'] must be a contiguous float64 CUDA tensor on %s" % (name, q.device))
        self._fused_check(st2, B, ok)
        out = {} if out is None else out
        if "ctrl" not in out:
            out["ctrl"] = torch.empty(B, self.n_ctrl, dtype=torch.float64, device=q.device)
        if want_u_all and "u_all" not in out:
            out["u_all"] = torch.empty(B, self.n, dtype=torch.float64, device=q.device)
        if want_status and "status" not in out:
            out["status"] = torch.empty(B, dtype=torch.uint8, device=q.device)
        io = _native.FusedIo()
        for name in self._FUSED_FIELDS:
            t = st2.get(name)
            setattr(io, name, t.data_ptr() if t is not None else None)
        for name in ("ctrl", "u_all", "status"):
            setattr(io, name, out[name].data_ptr() if name in out else None)
        sio = _native.SequenceIo()
        for name in ("wp_xyz", "wp_quat", "action", "entered", "timer", "err", "max_vel0", "target_xyz", "target_quat"):
            t = seq_state[name]
            want = torch.int32 if name in ("action", "entered", "timer") else torch.float64
            if not (t.is_cuda and t.dtype == want and t.is_contiguous() and t.shape[0] == B):
                raise ValueError("seq_state['%s'] must be a contiguous %s CUDA tensor with leading axis %d" % (name, want, B))
            setattr(sio, name, t.data_ptr())
        stream = torch.cuda.current_stream(q.device).cuda_stream
        with torch.cuda.device(q.device):
            _native.check(self.lib.irlosc_step_sequence(self._handle, B, C.byref(io), C.byref(seq.c_struct),
                                                        C.byref(sio), C.c_void_p(stream)))
        return out

    def step_waypoints(self, state: Dict, wp_state: Dict, threshold: float = 0.1, out: Optional[Dict] = None,
                       want_u_all: bool = False, want_status: bool = True) -> Dict:
        """One control step of B gain_test-style episodes (examples/gain_test.py:134-162): every arm steers to
        `wp_state["wps"][b, d, wp_idx[b, d]]` and moves on to its next waypoint (wrapping) once
        `|EE_XYZ - target| < threshold`.  `wp_state`: wps [B, D, W, 3] float64, n_wp (D ints), wp_idx [B, D] int32,
        target_xyz [B, D, 3], target_quat [B, D, 4] (updated in place)."""
        import torch
        q = state["q"]
        B = int(q.shape[0])
        st2 = dict(state, target_xyz=wp_state["target_xyz"], target_quat=wp_state["target_quat"])

        def ok(name, t):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.device == q.device):
                raise ValueError("state['%s'] must be a contiguous float64 CUDA tensor on %s" % (name, q.device))
        self._fused_check(st2, B, ok)
        wps, idx = wp_state["wps"], wp_state["wp_idx"]
        if not (wps.is_cuda and wps.dtype == torch.float64 and wps.is_contiguous() and wps.dim() == 4 and
                tuple(wps.shape[:2]) == (B, self.D) and wps.shape[3] == 3):
            raise ValueError("wp_state['wps'] must be a contiguous float64 CUDA tensor [B, D, W, 3]")
        if not (idx.is_cuda and idx.dtype == torch.int32 and idx.is_contiguous() and tuple(idx.shape) == (B, self.D)):
            raise ValueError("wp_state['wp_idx'] must be a contiguous int32 CUDA tensor [B, D]")
        out = {} if out is None else out
        if "ctrl" not in out:
            out["ctrl"] = torch.empty(B, self.n_ctrl, dtype=torch.float64, device=q.device)
        if want_u_all and "u_all" not in out:
            out["u_all"] = torch.empty(B, self.n, dtype=torch.float64, device=q.device)
        if want_status and "status" not in out:
            out["status"] = torch.empty(B, dtype=torch.uint8, device=q.device)
        io = _native.FusedIo()
        for name in self._FUSED_FIELDS:
            t = st2.get(name)
            setattr(io, name, t.data_ptr() if t is not None else None)
        for name in ("ctrl", "u_all", "status"):
            setattr(io, name, out[name].data_ptr() if name in out else None)
        wio = _native.WaypointsIo()
        wio.wps, wio.W, wio.threshold = wps.data_ptr(), int(wps.shape[2]), float(threshold)
        for d, n in enumerate(wp_state["n_wp"]):
            wio.n_wp[d] = int(n)
        wio.wp_idx = idx.data_ptr()
        wio.target_xyz, wio.target_quat = wp_state["target_xyz"].data_ptr(), wp_state["target_quat"].data_ptr()
        stream = torch.cuda.current_stream(q.device).cuda_stream
        with torch.cuda.device(q.device):
            _native.check(self.lib.irlosc_step_waypoints(self._handle, B, C.byref(io), C.byref(wio), C.c_void_p(stream)))
        return out

    def calc_error(self, ee_xyz, ee_quat, target_xyz, target_quat):
        """Batched `OSC.calc_error` (osc.py:101-118) on CUDA tensors -> (B, D, 6)."""
        import torch
        B = int(ee_xyz.shape[0])
        err = torch.empty(B, self.D, 6, dtype=torch.float64, device=ee_xyz.device)
        stream = torch.cuda.current_stream(ee_xyz.device).cuda_stream
        with torch.cuda.device(ee_xyz.device):
            _native.check(self.lib.irlosc_calc_error(
                self._handle, B, ee_xyz.data_ptr(), ee_quat.data_ptr(), target_xyz.data_ptr(),
                target_quat.data_ptr(), err.data_ptr(), C.c_void_p(stream)))
        return err


class _PinnedBlock:
    """Owns one cudaHostAlloc block; freed when the last array viewing it dies."""

    def __init__(self, nbytes: int):
        self.lib = _native.load()
        self.ptr = C.c_void_p()
        _native.check(self.lib.irlosc_host_alloc(C.byref(self.ptr), int(nbytes)))

    def __del__(self):
        try:
            self.lib.irlosc_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """Uninitialised numpy array in pinned host memory (`irlosc_host_alloc`), for `step_host`."""
    count = int(np.prod(shape))
    nbytes = max(count * np.dtype(dtype).itemsize, 1)
    block = _PinnedBlock(nbytes)
    cbuf = (C.c_char * nbytes).from_address(block.ptr.value)
    cbuf._block = block          # the numpy array references cbuf, cbuf keeps the block alive
    return np.frombuffer(cbuf, dtype=dtype, count=count).reshape(shape)

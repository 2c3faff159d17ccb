"""CPU: libirlosc.so builds for sm_100a, loads, and exports every symbol of include/irlosc.h.
No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(native_lib):
    from irl_control_b200 import _native
    header = open(os.path.join(ROOT, "include", "irlosc.h")).read()
    declared = set(re.findall(r"\b(irlosc_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for sym in declared:
        assert hasattr(native_lib, sym), sym
    assert native_lib.irlosc_abi_version() == _native.ABI_VERSION


def test_struct_sizes_match_header(native_lib, tmp_path):
    """Compile a 3-line C program against the header and compare sizeof with the ctypes mirror."""
    import subprocess
    from irl_control_b200 import _native
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "irlosc.h"\nint main(){printf("%zu %zu %zu\\n", '
                   'sizeof(irlosc_device_params), sizeof(irlosc_params), sizeof(irlosc_io));'
                   'printf("%zu %zu %zu %zu\\n", sizeof(irlosc_joint_model), sizeof(irlosc_frame_model), '
                   'sizeof(irlosc_model), sizeof(irlosc_fused_io));'
                   'printf("%zu %zu %zu %zu\\n", sizeof(irlosc_action), sizeof(irlosc_sequence), sizeof(irlosc_sequence_io), '
                   'sizeof(irlosc_waypoints_io));'
                   'return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert [int(x) for x in out] == [C.sizeof(_native.DeviceParams), C.sizeof(_native.Params), C.sizeof(_native.Io),
                                     C.sizeof(_native.JointModel), C.sizeof(_native.FrameModel),
                                     C.sizeof(_native.Model), C.sizeof(_native.FusedIo), C.sizeof(_native.Action),
                                     C.sizeof(_native.Sequence), C.sizeof(_native.SequenceIo), C.sizeof(_native.WaypointsIo)]


def test_library_is_sm100a_only(native_lib):
    import subprocess
    from irl_control_b200 import _native
    out = subprocess.check_output(["cuobjdump", "--list-elf", _native.library_path()]).decode()
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_create_validates_parameters_without_a_gpu(native_lib):
    from irl_control_b200 import _native
    p = _native.Params()
    h = C.c_void_p()
    p.abi_version = 99
    assert native_lib.irlosc_create(C.byref(p), C.byref(h)) == 1
    assert b"abi_version" in native_lib.irlosc_last_error()
    p.abi_version = _native.ABI_VERSION
    p.n = 25
    p.n_devices = 9
    assert native_lib.irlosc_create(C.byref(p), C.byref(h)) == 1
    assert b"n_devices" in native_lib.irlosc_last_error()
    assert not h.value


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from irl_control_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "library_path", lambda: str(tmp_path / "libirlosc.so"))
    with pytest.raises(_native.NativeLibraryError):
        _native.load()


def test_product_never_loads_the_host_harness():
    """tests/host_fused runs the fused kernel's per-instance function on the CPU for the tests; the
    package must not know about it."""
    pkg_dir = os.path.join(ROOT, "irl_control_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg_dir, fn)).read()
            assert "libfused_host" not in text and "host_fused" not in text and "import fused_host" not in text, fn


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "irl_control_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg_dir, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), fn

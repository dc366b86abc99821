"""CPU: the C-ABI library loads and exports exactly the symbols include/wmd.h declares (no compute calls)."""
import ctypes
import os
import re

from wavelet_monodepth_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(REPO, "include", "wmd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(wmd_[a-z0-9_]+)\s*\(", text))


def test_header_binding_and_library_agree():
    declared = header_symbols()
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name


def test_version_and_status_strings():
    lib = _lib.load()
    assert lib.wmd_version() == 100
    assert lib.wmd_status_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert lib.wmd_status_string(code) not in (b"ok", b"unknown status")
    assert lib.wmd_launch_count() == 0 or lib.wmd_launch_count() > 0


def test_host_only_size_queries():
    lib = _lib.load()
    assert lib.wmd_range_ws_bytes(4, 320 * 1024) == 65536 + 4 * 64 * 2 * 4
    assert lib.wmd_compact_ws_bytes(32, 320, 1024) == ((32 * 320 * 1024 + 2047) // 2048) * 4


def test_descriptor_layout_matches_c_struct():
    # the packed order in _lib mirrors include/wmd.h; natural alignment on LP64
    assert ctypes.sizeof(_lib.ConvDesc) % 8 == 0 and ctypes.sizeof(_lib.HeadDesc) % 8 == 0
    assert _lib.ConvDesc.x0.offset == 16 and _lib.ConvDesc.map0.offset == 32
    assert _lib.HeadDesc.t.offset == 16


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.wmd_idwt_haar_f32(None, None, None, None, 1.0, 0, 1, 1, 2, 2, None) == -1
    assert lib.wmd_dwt_haar_f32(1, 1, 1, 1, 1, 3, 4, None) == -2       # odd height
    assert lib.wmd_conv_rows_f32(None, None) == -1
    d = _lib.HeadDesc()
    assert lib.wmd_head_conv3x3_f32(ctypes.byref(d), None) == -1

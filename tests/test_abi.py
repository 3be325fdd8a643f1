"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/ffx.h declares,
and refuses to compute without a CUDA device (no CPU fallback)."""

import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def ffx():
    import __graft_entry__ as g

    g.build()
    from fast_forward import _ffx

    return _ffx


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ffx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ffx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(ffx):
    names = declared_symbols()
    assert len(names) >= 18
    handle = ctypes.CDLL(ffx.LIB_PATH)
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/ffx.h but not exported"
    # and the binding table covers exactly the header
    assert sorted(ffx.SYMBOLS) == names


def test_abi_version_and_error_string(ffx):
    lib = ffx.lib()
    assert lib.ffx_abi_version() == 1
    assert isinstance(lib.ffx_last_error(), bytes)
    assert lib.ffx_launch_count() >= 0


def test_no_cpu_fallback(ffx):
    """Without a CUDA device every compute entry point must fail loudly."""
    if ffx.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(ffx.FFXError) as e:
        ffx.DeviceIndex(768, capacity=16)
    assert e.value.code == -2 and "no CUDA device" in str(e.value)


def test_bad_arguments_are_rejected_before_cuda(ffx):
    lib = ffx.lib()
    h = ctypes.c_void_p()
    assert lib.ffx_index_create(0, 7, 768, 1, ctypes.byref(h)) == -1
    assert b"row kind" in lib.ffx_last_error()
    assert lib.ffx_index_create(0, 0, 0, 1, ctypes.byref(h)) == -1
    assert lib.ffx_rerank_host(None, 2, None, 1, None, None, None, 0.1, 0, None, None, None, None) == -1
    assert lib.ffx_merge_topk(0, None, None, 1, 1, 1, None, None, None) == -1
    assert lib.ffx_index_destroy(None) == 0
    assert lib.ffx_index_num_rows(None) == -1


def test_header_is_plain_c():
    """include/ffx.h is the C ABI: it must compile as C99 on its own (no C++, no torch types)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    header = os.path.join(ROOT, "include", "ffx.h")
    out = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", header],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_host_side_id_coding_needs_no_gpu(ffx):
    """The ffx_dict_* entry points are pure host code: they work in this (GPU-less) container."""
    import ctypes as C

    h = C.c_void_p()
    assert ffx.lib().ffx_dict_create(C.byref(h)) == 0
    assert ffx.lib().ffx_dict_size(h) == 0
    assert ffx.lib().ffx_dict_destroy(h) == 0

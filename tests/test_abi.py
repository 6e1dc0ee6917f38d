"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/mavg.h
declares, struct layouts agree with the ctypes mirror, and compute entry points fail loudly
(MAVG_ERR_NO_DEVICE, never a CPU fallback) when no GPU is present.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "mavg.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mavg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(mavg):
    from digital_signal_processsing_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mavg.h but not exported by libmavg.so"
    # and the ctypes table covers exactly the header
    assert sorted(n for n, _, _ in _lib.SYMBOLS) == declared


def test_no_torch_or_oracle_in_the_abi(mavg):
    from digital_signal_processsing_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "libcuda.so" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in syms


def test_struct_layout_matches_header(mavg, tmp_path):
    """sizeof/offsetof from the real header (compiled with gcc) == the ctypes mirror."""
    from digital_signal_processsing_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "mavg.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mavg_desc), sizeof(mavg_info),'
        ' sizeof(mavg_timing), sizeof(mavg_tuning), offsetof(mavg_desc, first_frame), offsetof(mavg_desc, tuning),'
        ' offsetof(mavg_info, shard_frames));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.Desc), ctypes.sizeof(_lib.Info), ctypes.sizeof(_lib.Timing), ctypes.sizeof(_lib.Tuning),
            _lib.Desc.first_frame.offset, _lib.Desc.tuning.offset, _lib.Info.shard_frames.offset]
    assert got == want


def test_version_and_strerror(mavg):
    from digital_signal_processsing_b200 import _lib
    lib = _lib.load()
    assert lib.mavg_version() == 300
    assert lib.mavg_strerror(0) == b"ok"
    assert b"block size" in lib.mavg_strerror(_lib.ERR_BLOCK_SIZE)
    assert lib.mavg_strerror(-99) == b"unknown status"


def test_argument_validation_without_gpu(mavg):
    """Validation happens before any device work, so it is testable on the CPU box."""
    from digital_signal_processsing_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    d = _lib.Desc()
    assert lib.mavg_plan_create(None, ctypes.byref(h)) == _lib.ERR_INVALID_ARG
    d.struct_size = 3
    assert lib.mavg_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.ERR_INVALID_ARG
    d.struct_size = ctypes.sizeof(_lib.Desc)
    d.channels, d.frames, d.window = 1, 100, 0
    assert lib.mavg_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.ERR_INVALID_ARG      # k == 0
    d.window, d.channels = 5, 0
    assert lib.mavg_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.ERR_INVALID_ARG      # channels == 0
    d.channels = 1
    for bad in (16, 48, 2048, 100):  # the reference's rule, basics/profilable_sm_vload4.cu:231
        d.block_size = bad
        assert lib.mavg_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.ERR_BLOCK_SIZE
    assert b"multiple of 32" in lib.mavg_last_error()
    assert lib.mavg_plan_destroy(None) == 0
    assert lib.mavg_run_host(None, None, None) == _lib.ERR_INVALID_ARG
    assert lib.mavg_run_device(None, None, None) == _lib.ERR_INVALID_ARG
    assert lib.mavg_run_host_sweep(None, 0, None, None) == _lib.ERR_INVALID_ARG
    one = (ctypes.c_void_p * 1)(None)
    assert lib.mavg_run_host_sweep(one, 1, ctypes.c_void_p(8), one) == _lib.ERR_INVALID_ARG    # a null plan in the list
    assert lib.mavg_run_cascade(None, None, None, None, 3) == _lib.ERR_INVALID_ARG
    assert lib.mavg_host_register(None, 64) == _lib.ERR_INVALID_ARG
    assert lib.mavg_host_unregister(None) == 0
    assert lib.mavg_plan_info(None, None) == _lib.ERR_INVALID_ARG
    assert lib.mavg_get_timing(None, None) == _lib.ERR_INVALID_ARG
    assert lib.mavg_prefix_sum(0, None, None, 10, 1, None) == _lib.ERR_INVALID_ARG


def test_fails_loudly_without_gpu(mavg):
    if mavg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(mavg.MavgError) as e:
        mavg.Plan(1024, 5)
    assert e.value.status == -4 and "no CPU fallback" in str(e.value)
    with pytest.raises(mavg.MavgError):
        mavg.moving_average(__import__("numpy").zeros(64, dtype="float32"), 3)


def test_product_does_not_import_the_oracle():
    """Only tests/, smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "digital_signal_processsing_b200")
    for base in (pkg, os.path.join(ROOT, "host"), os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                    assert "mavg_oracle.h" not in text and "libmavg_oracle" not in text, f


def test_links_and_runs_from_plain_c(mavg, tmp_path):
    """A C99 program (no C++, no Python) includes include/mavg.h, links against libmavg.so and gets status codes
    back -- the shape of a cgo / JNI / Rust FFI binding."""
    from digital_signal_processsing_b200 import _lib
    src = tmp_path / "client.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "mavg.h"\n'
        'int main(void) {\n'
        '    mavg_desc d; mavg_plan *p = NULL; int rc;\n'
        '    memset(&d, 0, sizeof d);\n'
        '    d.struct_size = sizeof d; d.dtype = MAVG_I16; d.channels = 2; d.frames = 1000; d.window = 0;\n'
        '    rc = mavg_plan_create(&d, &p);\n'
        '    printf("%d %d %s|%s\\n", mavg_version(), rc, mavg_strerror(rc), mavg_last_error());\n'
        '    return (rc == MAVG_ERR_INVALID_ARG && p == NULL) ? 0 : 1;\n'
        '}\n')
    exe = tmp_path / "client"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-lmavg", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split()[0] == "300" and "window" in out.stdout

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden_md5.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def oracle():
    """CPU checkers: builds the port (always) and the reference build (when /root/reference exists)."""
    from oracle import bindings
    bindings.build(ref=os.path.exists("/root/reference/h4m_audio_decode.c"), port=True)
    return bindings


@pytest.fixture(scope="session")
def emul_lib():
    """tests/emul: the kernel's block arithmetic + work order run serially on the CPU (test-only)."""
    import ctypes
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul")])
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "emul", "libhvqm4_emul.so"))
    lib.h4e_seq_create.restype = ctypes.c_void_p
    lib.h4e_seq_create.argtypes = [ctypes.c_int] * 5
    lib.h4e_seq_destroy.argtypes = [ctypes.c_void_p]
    lib.h4e_parse_begin.restype = ctypes.c_size_t
    lib.h4e_parse_begin.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    lib.h4e_parse_finish.restype = ctypes.c_uint32
    lib.h4e_parse_finish.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.h4e_seq_set_split_schedule.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.h4e_seq_errors.restype = ctypes.c_uint32
    lib.h4e_seq_errors.argtypes = [ctypes.c_void_p]
    lib.emul_recon_picture.argtypes = [ctypes.c_void_p] * 4
    lib.h4e_set_band_rows.argtypes = [ctypes.c_int]
    lib.emul_sweep_picture.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 3
    lib.emul_row_picture.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 2
    lib.emul_weighted.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5
    lib.emul_predict.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.emul_predict_diag.argtypes = lib.emul_predict.argtypes
    return lib


@pytest.fixture(scope="session")
def native_lib():
    """The product library (built in-tree by nvcc; loading it needs no GPU)."""
    from hvqm4_b200 import build
    build.build_native()
    from hvqm4_b200 import api
    return api

"""ctypes binding of the seeded synthetic HVQM4 stream generator (tools/h4mgen.c).

No .h4m game assets exist offline, so every test and benchmark input is produced
here.  The generator speaks the container and picture syntax the reference
decoder consumes (/root/reference/h4m_audio_decode.c:1970-2050, 2175-2247,
2427-2537); see the header of tools/h4mgen.c for the per-field citations.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_ROOT, "tools", "libh4mgen.so")
_SRC_PATH = os.path.join(_ROOT, "tools", "h4mgen.c")

DENSE = 0
REALISTIC = 1


class _Params(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("version", ctypes.c_int32),
        ("n_gops", ctypes.c_int32),
        ("profile", ctypes.c_int32),
        ("usec_per_frame", ctypes.c_int32),
        ("seed", ctypes.c_uint64),
        ("gop", ctypes.c_char_p),
    ]


def build(force: bool = False) -> str:
    """Compile tools/libh4mgen.so (plain gcc, seconds)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(_SRC_PATH):
        subprocess.check_call(
            ["gcc", "-O2", "-shared", "-fPIC", "-fvisibility=hidden", _SRC_PATH, "-o", _LIB_PATH]
        )
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.h4mgen_generate.argtypes = [
            ctypes.POINTER(_Params),
            ctypes.POINTER(ctypes.POINTER(ctypes.c_uint8)),
            ctypes.POINTER(ctypes.c_uint64),
        ]
        _lib.h4mgen_generate.restype = ctypes.c_int
        _lib.h4mgen_free.argtypes = [ctypes.POINTER(ctypes.c_uint8)]
    return _lib


def generate(width: int, height: int, version: int = 15, gop: str = "IPPP", n_gops: int = 1,
             seed: int = 1, profile: int = DENSE) -> bytes:
    """Returns a complete .h4m file image (plus 8 bytes of trailing slack)."""
    lib = _load()
    prm = _Params(width, height, version, n_gops, profile, 33367, seed, gop.encode())
    data = ctypes.POINTER(ctypes.c_uint8)()
    n = ctypes.c_uint64()
    rc = lib.h4mgen_generate(ctypes.byref(prm), ctypes.byref(data), ctypes.byref(n))
    if rc != 0:
        raise ValueError(f"h4mgen_generate failed with {rc}")
    try:
        return ctypes.string_at(data, n.value + 8)
    finally:
        lib.h4mgen_free(data)


# The five configurations BASELINE.json names (sizes/GOPs per SURVEY.md section 8d).
CONFIGS = {
    "cfg1_320x240_v15_I30": dict(width=320, height=240, version=15, gop="I" * 30, n_gops=1),
    "cfg2_640x480_v15_IP15": dict(width=640, height=480, version=15, gop="I" + "P" * 14, n_gops=2),
    "cfg3_640x480_v15_IPB": dict(width=640, height=480, version=15, gop="I" + "PBB" * 5, n_gops=2),
    "cfg4_320x240_v13_IPB": dict(width=320, height=240, version=13, gop="I" + "PBB" * 5, n_gops=2),
    "cfg5_640x480_v15_IPB_stream": dict(width=640, height=480, version=15, gop="I" + "PBB" * 5, n_gops=1),
}

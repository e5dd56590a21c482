"""ctypes bindings for the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package (hvqm4_b200/) never does.

  RefDecoder    -> oracle/_ref/libhvqm4_ref.so, the unmodified reference compiled
                   from /root/reference by oracle/Makefile (kind "reference")
  PortDecoder   -> oracle/liboracle_port.so, our own C restatement (kind "port")
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libhvqm4_ref.so")
PORT_LIB = os.path.join(_HERE, "liboracle_port.so")

I_FRAME, P_FRAME, B_FRAME = 0x10, 0x20, 0x30


def build(ref: bool = True, port: bool = True) -> None:
    targets = []
    if port and os.path.exists(os.path.join(_HERE, "hvqm4_oracle.c")):
        targets.append("port")
    if ref:
        targets.append("ref")
    if targets:
        subprocess.check_call(["make", "-s", "-C", _HERE] + targets)


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def have_port() -> bool:
    return os.path.exists(PORT_LIB)


class _Decoder:
    """Common walker API: open(data) -> iterate frames -> planar YUV bytes."""

    _prefix = ""
    _path = ""
    _libs: dict = {}

    def __init__(self, data: bytes):
        lib = self._lib()
        self._buf = ctypes.create_string_buffer(data, len(data))   # keep alive
        self._h = getattr(lib, self._prefix + "open")(self._buf, len(data))
        if not self._h:
            raise ValueError("not a HVQM4 stream")
        info = (ctypes.c_int32 * 6)()
        getattr(lib, self._prefix + "info")(self._h, info)
        self.width, self.height, self.n_frames, self.version, self.picsize, self.n_gops = list(info)
        self._out = (ctypes.c_uint8 * self.picsize)()

    @classmethod
    def _lib(cls):
        lib = cls._libs.get(cls._path)
        if lib is None:
            lib = ctypes.CDLL(cls._path)
            p = cls._prefix
            getattr(lib, p + "open").restype = ctypes.c_void_p
            getattr(lib, p + "open").argtypes = [ctypes.c_char_p, ctypes.c_size_t]
            getattr(lib, p + "close").argtypes = [ctypes.c_void_p]
            getattr(lib, p + "info").argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)]
            getattr(lib, p + "decode_next").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)]
            getattr(lib, p + "decode_next").restype = ctypes.c_int
            getattr(lib, p + "section_usage").argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
            getattr(lib, p + "get_map").argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)]
            getattr(lib, p + "get_nest").argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            getattr(lib, p + "bench").restype = ctypes.c_double
            getattr(lib, p + "bench").argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
            getattr(lib, p + "bench_mp").argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
            getattr(lib, p + "bench_mp").restype = ctypes.c_int
            getattr(lib, p + "yuv_to_rgb").argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
            getattr(lib, p + "yuv_to_rgb").restype = ctypes.c_int
            getattr(lib, p + "decode_audio").argtypes = [ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                                         ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
            getattr(lib, p + "decode_audio").restype = ctypes.c_int
            cls._libs[cls._path] = lib
        return lib

    @classmethod
    def yuv_to_rgb(cls, yuv: bytes, width: int, height: int) -> bytes:
        """Planar Y|U|V 4:2:0 -> interleaved RGB exactly as the reference's dumpRGB (h4m:895-926)."""
        assert len(yuv) == width * height * 3 // 2
        out = (ctypes.c_uint8 * (width * height * 3))()
        rc = getattr(cls._lib(), cls._prefix + "yuv_to_rgb")(yuv, width, height, out)
        if rc:
            raise RuntimeError(f"yuv_to_rgb failed ({rc})")
        return bytes(out)

    @classmethod
    def decode_audio(cls, state, channels: int, first: bool, sample_count: int, data: bytes):
        """One audio record as the reference's decode_audio (h4m:185-258).  state = [hist0, idx0, hist1, idx1, ...]
        (updated in place); data = the record payload behind its sample count.  Returns the int16 samples."""
        st = (ctypes.c_int32 * (2 * channels))(*state)
        out = (ctypes.c_int16 * max(1, (sample_count + 1) * channels))()
        n = getattr(cls._lib(), cls._prefix + "decode_audio")(st, channels, int(first), sample_count, data, len(data), out)
        if n < 0:
            raise RuntimeError(f"decode_audio failed ({n})")
        state[:] = list(st)
        return list(out[:n])

    def close(self):
        if self._h:
            getattr(self._lib(), self._prefix + "close")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode_next(self):
        """-> (frame_type, disp_id, display_index, yuv bytes) or None at end of stream."""
        meta = (ctypes.c_int32 * 4)()
        rc = getattr(self._lib(), self._prefix + "decode_next")(self._h, self._out, meta)
        if rc == 0:
            return None
        if rc < 0:
            raise ValueError(f"container error {rc}")
        return meta[0], meta[1], meta[2], bytes(self._out)

    def frames(self):
        while True:
            f = self.decode_next()
            if f is None:
                return
            yield f

    def section_usage(self):
        c = (ctypes.c_int32 * 17)()
        s = (ctypes.c_int32 * 17)()
        getattr(self._lib(), self._prefix + "section_usage")(self._h, c, s)
        return list(c), list(s)

    def get_map(self, plane: int):
        dims = (ctypes.c_int32 * 2)()
        getattr(self._lib(), self._prefix + "get_map")(self._h, plane, None, dims)
        buf = (ctypes.c_uint8 * (dims[0] * dims[1] * 2))()
        getattr(self._lib(), self._prefix + "get_map")(self._h, plane, buf, dims)
        return dims[0], dims[1], bytes(buf)

    def get_nest(self) -> bytes:
        buf = (ctypes.c_uint8 * (70 * 38))()
        getattr(self._lib(), self._prefix + "get_nest")(self._h, buf)
        return bytes(buf)

    @classmethod
    def md5s(cls, data: bytes):
        d = cls(data)
        out = [(t, disp, hashlib.md5(yuv).hexdigest()) for t, disp, _, yuv in d.frames()]
        d.close()
        return out

    @classmethod
    def bench(cls, data: bytes, reps: int = 1):
        """Single process: (seconds inside Decode?pic calls, frames)."""
        n = ctypes.c_int64()
        t = getattr(cls._lib(), cls._prefix + "bench")(data, len(data), reps, ctypes.byref(n))
        return t, n.value

    @classmethod
    def bench_mp(cls, data: bytes, nproc: int, reps: int = 1):
        """nproc forked workers: (wall seconds, sum of in-decode seconds, total frames)."""
        out = (ctypes.c_double * 3)()
        rc = getattr(cls._lib(), cls._prefix + "bench_mp")(data, len(data), nproc, reps, out)
        if rc:
            raise RuntimeError(f"bench_mp failed {rc}")
        return out[0], out[1], int(out[2])


    @classmethod
    def bench_mp_streams(cls, datas, nproc: int, reps: int):
        """nproc forked workers cycle through `datas` (worker i: streams i, i + nproc, ...), `reps` GOP decodes each:
        (wall seconds, sum of in-decode seconds, total frames)."""
        n = len(datas)
        bufs = [ctypes.create_string_buffer(d, len(d)) for d in datas]
        ptrs = (ctypes.c_char_p * n)(*[ctypes.cast(b, ctypes.c_char_p) for b in bufs])
        lens = (ctypes.c_size_t * n)(*[len(d) for d in datas])
        out = (ctypes.c_double * 3)()
        fn = getattr(cls._lib(), cls._prefix + "bench_mp_streams")
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        fn.restype = ctypes.c_int
        rc = fn(ptrs, lens, n, nproc, reps, out)
        if rc:
            raise RuntimeError(f"bench_mp_streams failed {rc}")
        return out[0], out[1], int(out[2])


class RefDecoder(_Decoder):
    _prefix = "ref_"
    _path = REF_LIB


class PortDecoder(_Decoder):
    _prefix = "port_"
    _path = PORT_LIB

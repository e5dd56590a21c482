"""Python mirror of the C ABI in include/hvqm4.h (ctypes; no torch types cross the boundary).

Two layers, both thin:

* ``SeqDecoder``  -- the reference's SDK call protocol, one stream, host buffers:
  InitSeqObj -> BuffSize -> SetBuffer -> DecodeIpic/Ppic/Bpic
  (/root/reference/h4m_audio_decode.c:2409-2418, 2099-2104).  ``Player`` adds the
  reference's past/present/future rotation (h4m:2087-2093, 2131-2137) on top, so a
  parity test reads like the reference's own ``decode_video``.
* ``Batch``       -- the batched multi-stream runtime (HVQM4Batch*): device-resident
  surfaces, host thread pool, one upload + one kernel launch per step.

The shared library must have been built (``hvqm4_b200.build.build_native()``); importing
this module never falls back to a CPU implementation -- there is none.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_size_t, c_uint8, c_uint16, c_uint32, c_uint64, c_void_p

from .build import LIB

I_FRAME, P_FRAME, B_FRAME = 0x10, 0x20, 0x30

ERR_TRUNCATED = 1 << 0
ERR_OVERFLOW = 1 << 6
ERR_NO_DEVICE = 1 << 16
ERR_CUDA = 1 << 17
ERR_ARGUMENT = 1 << 18
ERR_NOMEM = 1 << 19


class HVQM4Error(RuntimeError):
    def __init__(self, bits: int, what: str = ""):
        self.bits = bits
        super().__init__(f"hvqm4_b200 error bits 0x{bits:x} {what}".strip())


class SeqObj(ctypes.Structure):
    _fields_ = [("state", c_void_p), ("width", c_uint16), ("height", c_uint16), ("h_samp", c_uint8), ("v_samp", c_uint8)]


class VideoInfo(ctypes.Structure):
    _fields_ = [("hres", c_uint16), ("vres", c_uint16), ("h_samp", c_uint8), ("v_samp", c_uint8), ("video_mode", c_uint8)]


class FileInfo(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("version", "width", "height", "h_samp", "v_samp", "n_gops", "n_video_frames", "usec_per_frame",
                                       "n_audio_frames", "audio_channels", "audio_bits", "audio_format", "audio_sample_rate")]


class FrameRef(ctypes.Structure):
    _fields_ = [("offset", c_uint32), ("bytes", c_uint32), ("frame_type", c_uint16), ("gop", c_uint16), ("disp_id", c_uint32)]


class AudioRef(ctypes.Structure):
    _fields_ = [("offset", c_uint32), ("bytes", c_uint32), ("gop", c_uint16), ("first", c_uint16), ("samples", c_uint32)]


class AudioState(ctypes.Structure):
    _fields_ = [("hist", ctypes.c_int16 * 2), ("idx", ctypes.c_int8 * 2), ("pad", ctypes.c_int8 * 2)]


# every symbol include/hvqm4.h declares: (restype, argtypes)
SIGNATURES = {
    "HVQM4InitDecoder": (None, []),
    "HVQM4InitSeqObj": (None, [POINTER(SeqObj), POINTER(VideoInfo)]),
    "HVQM4BuffSize": (c_uint32, [POINTER(SeqObj)]),
    "HVQM4SetBuffer": (None, [POINTER(SeqObj), c_void_p]),
    "HVQM4DecodeIpic": (None, [POINTER(SeqObj), c_void_p, c_void_p]),
    "HVQM4DecodePpic": (None, [POINTER(SeqObj), c_void_p, c_void_p, c_void_p]),
    "HVQM4DecodeBpic": (None, [POINTER(SeqObj), c_void_p, c_void_p, c_void_p, c_void_p]),
    "HVQM4SetVersion": (c_int, [POINTER(SeqObj), c_int]),
    "HVQM4SetFrameBytes": (None, [POINTER(SeqObj), c_uint32]),
    "HVQM4GetLastError": (c_uint32, [POINTER(SeqObj)]),
    "HVQM4GetLastCudaError": (c_int, []),
    "HVQM4ReleaseBuffer": (None, [POINTER(SeqObj)]),
    "HVQM4InvalidateFrame": (None, [POINTER(SeqObj), c_void_p]),
    "HVQM4ConvertRGB": (c_int, [POINTER(SeqObj), c_void_p, c_void_p]),
    "HVQM4BatchCreate": (c_void_p, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "HVQM4BatchDestroy": (None, [c_void_p]),
    "HVQM4BatchDecode": (c_int, [c_void_p, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_void_p), POINTER(c_uint32)]),
    "HVQM4BatchSetEntropyMode": (c_int, [c_void_p, c_int]),
    "HVQM4BatchSetHostShare": (c_int, [c_void_p, c_int]),
    "HVQM4DevEntropyProfile": (None, [POINTER(c_uint64)]),
    "HVQM4BatchSync": (c_int, [c_void_p]),
    "HVQM4BatchReadFrame": (c_int, [c_void_p, c_int, c_void_p]),
    "HVQM4BatchReadFrameAsync": (c_int, [c_void_p, c_int, c_void_p]),
    "HVQM4BatchReadFramesAsync": (c_int, [c_void_p, c_int, POINTER(c_int32), c_void_p, c_size_t]),
    "HVQM4BatchReadFramesRGBAsync": (c_int, [c_void_p, c_int, POINTER(c_int32), c_void_p, c_size_t]),
    "HVQM4BatchFramePtr": (c_void_p, [c_void_p, c_int]),
    "HVQM4BatchRecord": (c_int, [c_void_p, c_int]),
    "HVQM4BatchReplay": (c_float, [c_void_p, c_int]),
    "HVQM4BatchStats": (None, [c_void_p, POINTER(c_uint64)]),
    "HVQM4KernelLaunches": (ctypes.c_longlong, []),
    "HVQM4SweepLaunches": (ctypes.c_longlong, []),
    "HVQM4SweepErrors": (ctypes.c_int, []),
    "HVQM4RowLaunches": (ctypes.c_longlong, []),
    "HVQM4RowErrors": (ctypes.c_int, []),
    "HVQM4SetReconMode": (None, [c_int]),
    "HVQM4HostAlloc": (c_void_p, [c_size_t]),
    "HVQM4HostFree": (None, [c_void_p]),
    "HVQM4HostRegister": (c_int, [c_void_p, c_size_t]),
    "HVQM4HostUnregister": (c_int, [c_void_p]),
    "HVQM4ParseFile": (c_int, [c_char_p, c_size_t, POINTER(FileInfo), POINTER(FrameRef), c_int]),
    "HVQM4ParseFileAudio": (c_int, [c_char_p, c_size_t, POINTER(AudioRef), c_int]),
    "HVQM4DecodeAudioBatch": (c_int, [c_int, c_int, POINTER(AudioState), POINTER(c_int32), POINTER(c_void_p), POINTER(c_uint32),
                                      POINTER(c_void_p), POINTER(c_uint32), POINTER(c_uint32)]),
    "HVQM4PlayerOpen": (c_void_p, [c_void_p, c_size_t]),
    "HVQM4PlayerClose": (None, [c_void_p]),
    "HVQM4PlayerInfo": (c_int, [c_void_p, POINTER(FileInfo)]),
    "HVQM4PlayerNextFrame": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_uint32), POINTER(c_uint32)]),
    "HVQM4PlayerFrameRGB": (c_int, [c_void_p, c_void_p]),
    "HVQM4PlayerNextAudio": (c_int, [c_void_p, c_void_p, c_uint32]),
    "HVQM4PlayerErrors": (c_uint32, [c_void_p]),
}

_lib = None


def lib():
    """Loads libhvqm4_b200.so; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise ImportError(
                f"{LIB} is missing: build it with `python -m hvqm4_b200.build` (nvcc, sm_100a). "
                "hvqm4_b200 has no CPU reconstruction path."
            )
        l = ctypes.CDLL(LIB)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def parse_file(data: bytes):
    """.h4m container walk -> (FileInfo, [FrameRef])."""
    info = FileInfo()
    n = lib().HVQM4ParseFile(data, len(data), ctypes.byref(info), None, 0)
    if n < 0:
        raise ValueError(f"malformed .h4m container ({n})")
    frames = (FrameRef * n)()
    lib().HVQM4ParseFile(data, len(data), ctypes.byref(info), frames, n)
    return info, list(frames)


def parse_file_audio(data: bytes):
    """The audio records of a .h4m container -> [AudioRef]."""
    n = lib().HVQM4ParseFileAudio(data, len(data), None, 0)
    if n < 0:
        raise ValueError(f"malformed .h4m container ({n})")
    refs = (AudioRef * max(n, 1))()
    lib().HVQM4ParseFileAudio(data, len(data), refs, n)
    return list(refs[:n])


def decode_audio_batch(channels: int, states, first, payloads, capacity: int = 1 << 16):
    """HVQM4DecodeAudioBatch: payloads[i] = one audio record payload (sample count + data) of an
    independent stream, states[i] an AudioState carried between the records of a GOP block.
    Returns (error bits, [interleaved int16 samples per stream])."""
    n = len(payloads)
    st = (AudioState * n)(*states)
    bufs = [ctypes.create_string_buffer(p, len(p)) for p in payloads]
    ptrs = (c_void_p * n)(*[ctypes.addressof(b) for b in bufs])
    lens = (c_uint32 * n)(*[len(p) for p in payloads])
    outs = [(ctypes.c_int16 * (capacity * channels))() for _ in range(n)]
    optrs = (c_void_p * n)(*[ctypes.addressof(o) for o in outs])
    caps = (c_uint32 * n)(*([capacity] * n))
    got = (c_uint32 * n)()
    rc = lib().HVQM4DecodeAudioBatch(n, channels, st, (c_int32 * n)(*[int(f) for f in first]), ptrs, lens, optrs, caps, got)
    for i in range(n):
        states[i] = st[i]
    return rc, [list(outs[i][:got[i] * channels]) for i in range(n)]


class FilePlayer:
    """The reference program as a library (HVQM4Player*): pictures in file order, display index, RGB, audio."""

    def __init__(self, data: bytes):
        self._buf = ctypes.create_string_buffer(data, len(data) + 8)
        self._h = lib().HVQM4PlayerOpen(ctypes.addressof(self._buf), len(data))
        if not self._h:
            raise HVQM4Error(ERR_ARGUMENT, "HVQM4PlayerOpen")
        self.info = FileInfo()
        lib().HVQM4PlayerInfo(self._h, ctypes.byref(self.info))
        self.frame_bytes = self.info.width * self.info.height * 3 // 2

    def frames(self):
        """Yields (frame_type, display_index, planar bytes) per video record."""
        ptr, disp, ftype = c_void_p(), c_uint32(), c_uint32()
        while True:
            rc = lib().HVQM4PlayerNextFrame(self._h, ctypes.byref(ptr), ctypes.byref(disp), ctypes.byref(ftype))
            if rc == 0:
                return
            if rc < 0:
                raise HVQM4Error(ERR_CUDA, f"HVQM4PlayerNextFrame ({rc})")
            yield ftype.value, disp.value, ctypes.string_at(ptr.value, self.frame_bytes)

    def rgb(self) -> bytes:
        out = (c_uint8 * (self.info.width * self.info.height * 3))()
        rc = lib().HVQM4PlayerFrameRGB(self._h, out)
        if rc:
            raise HVQM4Error(rc, "HVQM4PlayerFrameRGB")
        return bytes(out)

    def audio(self, capacity: int = 1 << 16):
        """Yields the interleaved int16 samples of every audio record."""
        out = (ctypes.c_int16 * (capacity * max(1, self.info.audio_channels)))()
        while True:
            n = lib().HVQM4PlayerNextAudio(self._h, out, capacity)
            if n == 0:
                return
            if n < 0:
                raise HVQM4Error(ERR_ARGUMENT, f"HVQM4PlayerNextAudio ({n})")
            yield list(out[:n * self.info.audio_channels])

    def errors(self) -> int:
        return lib().HVQM4PlayerErrors(self._h)

    def close(self):
        if self._h:
            lib().HVQM4PlayerClose(self._h)
            self._h = None


class SeqDecoder:
    """One stream through the SDK-compatible entry points, host frame buffers."""

    def __init__(self, width: int, height: int, version: int = 15, h_samp: int = 2, v_samp: int = 2):
        l = lib()
        l.HVQM4InitDecoder()
        self.seq = SeqObj()
        vi = VideoInfo(width, height, h_samp, v_samp, 0)
        l.HVQM4InitSeqObj(ctypes.byref(self.seq), ctypes.byref(vi))
        self._work = ctypes.create_string_buffer(l.HVQM4BuffSize(ctypes.byref(self.seq)))
        l.HVQM4SetBuffer(ctypes.byref(self.seq), self._work)
        if l.HVQM4SetVersion(ctypes.byref(self.seq), version) != 0:
            raise HVQM4Error(ERR_ARGUMENT, "unsupported geometry or version")
        self.width, self.height = width, height
        self.frame_bytes = width * height * 3 // 2

    def close(self):
        if self.seq.state:
            lib().HVQM4ReleaseBuffer(ctypes.byref(self.seq))
            self.seq.state = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self):
        e = lib().HVQM4GetLastError(ctypes.byref(self.seq))
        if e:
            raise HVQM4Error(e)

    def decode(self, frame_type: int, pic: bytes, present, past=None, future=None):
        """pic = record bytes from offset 4; present/past/future: writable buffers (e.g. ctypes arrays)."""
        l = lib()
        buf = ctypes.create_string_buffer(pic, len(pic) + 8)
        l.HVQM4SetFrameBytes(ctypes.byref(self.seq), len(pic))
        if frame_type == I_FRAME:
            l.HVQM4DecodeIpic(ctypes.byref(self.seq), buf, present)
        elif frame_type == P_FRAME:
            l.HVQM4DecodePpic(ctypes.byref(self.seq), buf, present, past)
        elif frame_type == B_FRAME:
            l.HVQM4DecodeBpic(ctypes.byref(self.seq), buf, present, past, future)
        else:
            raise ValueError("frame_type")
        self._check()


    def to_rgb(self, frame) -> bytes:
        """The reference's dumpRGB (h4m:895-926) of a planar frame (a buffer passed to decode(), or bytes)."""
        out = (c_uint8 * (self.width * self.height * 3))()
        if isinstance(frame, (bytes, bytearray)):
            frame = (c_uint8 * len(frame)).from_buffer_copy(frame)
        rc = lib().HVQM4ConvertRGB(ctypes.byref(self.seq), frame, out)
        if rc:
            raise HVQM4Error(rc, "HVQM4ConvertRGB")
        return bytes(out)


class Player:
    """The reference's decode_video() loop minus file output: demux + buffer rotation + SDK calls."""

    def __init__(self, data: bytes, rgb: bool = False):
        self.data = data
        self.rgb = rgb
        self.last_rgb = None
        self.info, self.frames = parse_file(data)
        self.dec = SeqDecoder(self.info.width, self.info.height, self.info.version, self.info.h_samp, self.info.v_samp)
        n = self.dec.frame_bytes
        self.past = (c_uint8 * n)()
        self.present = (c_uint8 * n)()
        self.future = (c_uint8 * n)()

    def __iter__(self):
        for fr in self.frames:
            t = fr.frame_type
            if t != B_FRAME:
                self.past, self.future = self.future, self.past
            pic = self.data[fr.offset: fr.offset + fr.bytes]
            self.dec.decode(t, pic, self.present, self.past, self.future)
            self.last_rgb = self.dec.to_rgb(self.present) if self.rgb else None   # dumpRGB, h4m:2126
            yield t, fr.disp_id, bytes(self.present)
            if t != B_FRAME:
                self.present, self.future = self.future, self.present

    def close(self):
        self.dec.close()


class Batch:
    """n_streams independent streams of one geometry on one GPU (HVQM4Batch*)."""

    def __init__(self, n_streams: int, width: int, height: int, version: int = 15, device: int = -1, host_threads: int = 0,
                 gpu_entropy: bool = False, host_share: int = 0):
        self._h = lib().HVQM4BatchCreate(device, n_streams, width, height, version, host_threads)
        if not self._h:
            raise HVQM4Error(ERR_NO_DEVICE, "HVQM4BatchCreate failed (no CUDA device, or unsupported geometry)")
        if gpu_entropy:
            rc = lib().HVQM4BatchSetEntropyMode(self._h, 1)
            if rc:
                raise HVQM4Error(rc, "HVQM4BatchSetEntropyMode")
            if host_share:      # streams [0, host_share) go through the host stage, next to the parse kernel
                rc = lib().HVQM4BatchSetHostShare(self._h, host_share)
                if rc:
                    raise HVQM4Error(rc, "HVQM4BatchSetHostShare")
        self.n_streams = n_streams
        self.width, self.height = width, height
        self.frame_bytes = width * height * 3 // 2
        self._keep = None

    def close(self):
        if self._h:
            lib().HVQM4BatchDestroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, stream_ids, frame_types, frame_ptrs, frame_bytes):
        """One step.  frame_ptrs: addresses (ints) of picture headers that stay valid during the call -- and, when the
        pictures lie in memory registered with HVQM4HostRegister and the entropy stage runs on the GPU, unmodified until
        sync() returns (the GPU fetches them after the call; include/hvqm4.h)."""
        n = len(stream_ids)
        ids = (c_int32 * n)(*stream_ids)
        tys = (c_int32 * n)(*frame_types)
        ptrs = (c_void_p * n)(*frame_ptrs)
        lens = (c_uint32 * n)(*frame_bytes)
        rc = lib().HVQM4BatchDecode(self._h, n, ids, tys, ptrs, lens)
        if rc:
            raise HVQM4Error(rc, "HVQM4BatchDecode")

    def decode_prepared(self, step):
        """step = (n, ids, tys, ptrs, lens) ctypes arrays built once by prepare_step()."""
        rc = lib().HVQM4BatchDecode(self._h, *step)
        if rc:
            raise HVQM4Error(rc, "HVQM4BatchDecode")

    @staticmethod
    def prepare_step(stream_ids, frame_types, frame_ptrs, frame_bytes):
        n = len(stream_ids)
        return (n, (c_int32 * n)(*stream_ids), (c_int32 * n)(*frame_types), (c_void_p * n)(*frame_ptrs), (c_uint32 * n)(*frame_bytes))

    def sync(self):
        rc = lib().HVQM4BatchSync(self._h)
        if rc:
            raise HVQM4Error(rc, "HVQM4BatchSync")

    def read_frame(self, stream_id: int) -> bytes:
        out = (c_uint8 * self.frame_bytes)()
        rc = lib().HVQM4BatchReadFrame(self._h, stream_id, out)
        if rc:
            raise HVQM4Error(rc, "HVQM4BatchReadFrame")
        return bytes(out)

    def read_frame_async(self, stream_id: int, host_ptr: int):
        rc = lib().HVQM4BatchReadFrameAsync(self._h, stream_id, host_ptr)
        if rc:
            raise HVQM4Error(rc, "HVQM4BatchReadFrameAsync")

    def read_frames_async(self, ids_array, n: int, host_base: int, host_stride: int):
        rc = lib().HVQM4BatchReadFramesAsync(self._h, n, ids_array, host_base, host_stride)
        if rc:
            raise HVQM4Error(rc, "HVQM4BatchReadFramesAsync")

    def read_frames_rgb_async(self, ids_array, n: int, host_base: int, host_stride: int):
        """The reference's dumpRGB of the last decoded picture of n streams, into (pinned) host memory."""
        rc = lib().HVQM4BatchReadFramesRGBAsync(self._h, n, ids_array, host_base, host_stride)
        if rc:
            raise HVQM4Error(rc, "HVQM4BatchReadFramesRGBAsync")

    def read_frames_rgb(self, stream_ids) -> list:
        n = len(stream_ids)
        size = self.width * self.height * 3
        out = (c_uint8 * (n * size))()
        self.read_frames_rgb_async((c_int32 * n)(*stream_ids), n, ctypes.addressof(out), size)
        self.sync()
        raw = bytes(out)
        return [raw[i * size:(i + 1) * size] for i in range(n)]

    def frame_ptr(self, stream_id: int) -> int:
        return lib().HVQM4BatchFramePtr(self._h, stream_id)

    def record(self, enable: bool):
        lib().HVQM4BatchRecord(self._h, int(enable))

    def replay(self, repeats: int = 1) -> float:
        ms = lib().HVQM4BatchReplay(self._h, repeats)
        if ms < 0:
            raise HVQM4Error(ERR_CUDA, "HVQM4BatchReplay")
        return ms

    def stats(self) -> dict:
        out = (c_uint64 * 8)()
        lib().HVQM4BatchStats(self._h, out)
        keys = ("pictures", "launches", "symbol_bytes", "algorithmic_bytes", "host_ns", "inter_mcbs", "total_mcbs", "band_launches")
        return dict(zip(keys, list(out)))


def set_recon_mode(mode: int) -> None:
    """0 auto, 1..4 fused band kernel, 5 sweep kernel, 6 row kernel, 7 band kernel with a shared-memory tile, <0 map + record kernels (see include/hvqm4.h)."""
    lib().HVQM4SetReconMode(mode)


def sweep_launches() -> int:
    return lib().HVQM4SweepLaunches()


def sweep_errors() -> int:
    return lib().HVQM4SweepErrors()


def row_launches() -> int:
    return lib().HVQM4RowLaunches()


def row_errors() -> int:
    return lib().HVQM4RowErrors()


def kernel_launches() -> int:
    return lib().HVQM4KernelLaunches()


def decode_streams(files, device: int = -1, host_threads: int = 0, gpu_entropy: bool = False):
    """Decodes several .h4m images of identical geometry/GOP structure in lock step through
    the batch runtime.  Yields, per step, a list of (frame_type, disp_id, yuv bytes), one per stream."""
    parsed = [parse_file(f) for f in files]
    info0 = parsed[0][0]
    n = len(files)
    bufs = [ctypes.create_string_buffer(f, len(f) + 8) for f in files]
    bases = [ctypes.addressof(b) for b in bufs]
    batch = Batch(n, info0.width, info0.height, info0.version, device, host_threads, gpu_entropy=gpu_entropy)
    try:
        steps = len(parsed[0][1])
        for k in range(steps):
            frs = [p[1][k] for p in parsed]
            batch.decode(list(range(n)), [f.frame_type for f in frs], [bases[i] + frs[i].offset for i in range(n)],
                         [f.bytes for f in frs])
            batch.sync()
            yield [(frs[i].frame_type, frs[i].disp_id, batch.read_frame(i)) for i in range(n)]
    finally:
        batch.close()

"""Shared helpers for the tests (pure Python container walk, emulation driver)."""
import ctypes
import hashlib
import struct

import numpy as np

I_FRAME, P_FRAME, B_FRAME = 0x10, 0x20, 0x30


def demux(data: bytes):
    """-> (version, width, height, [(frame_type, disp_id, picture bytes)]) following
    /root/reference/h4m_audio_decode.c:2192-2211 (header), 2429-2438 (GOP), 2456-2458 (record)."""
    version = 15 if data[:9] == b"HVQM4 1.5" else 13
    n_gops = struct.unpack(">I", data[0x18:0x1C])[0]
    w, h = struct.unpack(">HH", data[0x34:0x38])
    pos = 0x44
    recs = []
    for _ in range(n_gops):
        nv, na = struct.unpack(">II", data[pos + 8:pos + 16])
        pos += 20
        while nv or na:
            id1, id2, size = struct.unpack(">HHI", data[pos:pos + 8])
            pos += 8
            if id1 == 1:
                recs.append((id2, struct.unpack(">I", data[pos:pos + 4])[0], data[pos + 4:pos + size]))
                nv -= 1
            else:
                na -= 1
            pos += size
    return version, w, h, recs


def emul_decode(lib, data: bytes, sweep=None, stats=None, row=None):
    """Host stage (entropy.c) + CPU emulation of the kernel work order; yields (type, yuv bytes, err).
    sweep = (band rows, shared-memory bytes, look-ahead bands): emulate the sweep kernel's plan and work order
    (tests/emul/sweep_emul.cpp); pictures its plan does not serve go the band kernel's way, as in the product,
    and are counted in stats["fallback"].
    row = (shared-memory bytes, look-ahead rows): the same for the row kernel (tests/emul/row_emul.cpp)."""
    version, w, h, recs = demux(data)
    lib.h4e_set_band_rows(1 if (sweep is not None or row is not None) else 8)     # record bands of one macroblock row for those kernels
    seq = lib.h4e_seq_create(w, h, 2, 2, int(version == 15))
    lib.h4e_set_band_rows(8)
    assert seq
    fb = w * h * 3 // 2
    bufs = [np.zeros(fb + 64, np.uint8) for _ in range(3)]
    past, present, future = 0, 1, 2
    try:
        for ty, _, pic in recs:
            if ty != B_FRAME:
                past, future = future, past
            padded = pic + b"\0" * 8
            n = lib.h4e_parse_begin(seq, ty, padded, len(pic))
            assert n > 0
            blob = np.zeros(n, np.uint8)
            err = lib.h4e_parse_finish(seq, blob.ctypes.data)
            fut = bufs[present] if ty == P_FRAME else bufs[future]
            rc = 1
            if sweep is not None:
                rc = lib.emul_sweep_picture(blob.ctypes.data, bufs[present].ctypes.data, bufs[past].ctypes.data, fut.ctypes.data, *sweep)
                assert rc in (0, 1), f"sweep emulation failed ({rc})"
                if stats is not None:
                    stats["sweep" if rc == 0 else "fallback"] = stats.get("sweep" if rc == 0 else "fallback", 0) + 1
            if row is not None:
                rc = lib.emul_row_picture(blob.ctypes.data, bufs[present].ctypes.data, bufs[past].ctypes.data, fut.ctypes.data, *row)
                assert rc in (0, 1), f"row emulation failed ({rc})"
                if stats is not None:
                    stats["row" if rc == 0 else "fallback"] = stats.get("row" if rc == 0 else "fallback", 0) + 1
            if rc == 1:
                rc = lib.emul_recon_picture(blob.ctypes.data, bufs[present].ctypes.data, bufs[past].ctypes.data, fut.ctypes.data)
            assert rc == 0, f"segment table / prefix sum disagree ({rc})"
            yield ty, bufs[present][:fb].tobytes(), err
            if ty != B_FRAME:
                present, future = future, present
    finally:
        lib.h4e_seq_destroy(seq)


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def all_triples_pictures():
    """Eight 2048x1024 planar 4:2:0 pictures that together contain every (y, u, v) byte triple:
    the chroma planes enumerate the 65536 (u, v) pairs eight times over, and the 4 luma samples
    of a 2x2 tile x 8 copies x 8 pictures enumerate the 256 luma values.  Yields (yuv, w, h)."""
    import numpy as np
    w, h = 2048, 1024
    idx = np.arange((w // 2) * (h // 2), dtype=np.uint32)
    u = (idx & 0xFF).astype(np.uint8).tobytes()
    v = ((idx >> 8) & 0xFF).astype(np.uint8).tobytes()
    copy = (idx >> 16).reshape(h // 2, w // 2)                      # 0..7
    for pic in range(8):
        y = np.empty((h, w), np.uint8)
        for dy in range(2):
            for dx in range(2):
                y[dy::2, dx::2] = ((pic * 8 + copy) * 4 + dy * 2 + dx).astype(np.uint8)
        yield y.tobytes() + u + v, w, h


def with_audio(data: bytes, channels: int, per_gop: int, samples: int, seed: int):
    """Splices `per_gop` IMA-ADPCM audio records (random codes, valid seeds) into every GOP block
    of a generated .h4m file, the way the container interleaves them (h4m:2456-2507), and patches
    the header counts.  Returns (file, [(gop, first, payload)])."""
    import random
    rng = random.Random(seed)
    n_gops = struct.unpack(">I", data[0x18:0x1C])[0]
    out = bytearray(data[:0x44])
    pos, records = 0x44, []
    for g in range(n_gops):
        hdr = bytearray(data[pos:pos + 20])
        nv, na = struct.unpack(">II", hdr[8:16])
        assert na == 0
        pos += 20
        body = bytearray()
        for k in range(nv):
            id1, id2, size = struct.unpack(">HHI", data[pos:pos + 8])
            body += data[pos:pos + 8 + size]
            pos += 8 + size
            if k < per_gop:
                first = k == 0
                n = samples + rng.randrange(8)
                seed_bytes = b""
                if first:
                    for _ in range(channels):
                        seed_bytes += bytes([rng.randrange(256), (rng.randrange(2) << 7) | rng.randrange(89)])
                codes = bytes(rng.randrange(256) for _ in range(((n - (1 if first else 0)) * channels + 1) // 2 + rng.randrange(3)))
                payload = struct.pack(">I", n) + seed_bytes + codes
                body += struct.pack(">HHI", 0, 0, len(payload)) + payload
                records.append((g, first, payload))
        struct.pack_into(">I", hdr, 4, len(body))
        struct.pack_into(">I", hdr, 12, min(per_gop, nv))
        out += hdr + body
    struct.pack_into(">I", out, 0x14, len(out) - 0x44)
    struct.pack_into(">I", out, 0x20, len(records))
    out[0x3C] = channels
    out[0x3D] = 16
    struct.pack_into(">I", out, 0x40, 22050)
    return bytes(out), records

"""Host serial stage (entropy.c) + the kernel's block arithmetic and work order, run on the CPU
through tests/emul, against the golden vectors and the oracle port.  No GPU."""
import ctypes

import numpy as np
import pytest

from hvqm4_b200 import synth
from tests.h4m_util import demux, emul_decode, md5


@pytest.mark.parametrize("name", [
    "cfg1_320x240_v15_I30", "cfg3_640x480_v15_IPB", "cfg4_320x240_v13_IPB", "cfg5_stream511",
    "realistic_640x480_v15_IPB", "min_280x152_v15_IPB", "ragged_328x248_v15_IPB", "wide_1024x576_v13_IPB",
    "hd_1280x720_v15_IPB", "tiny_16x16_v15_IPB", "small_64x48_v13_IPB", "mirror_h_200x152_v15_IPB", "mirror_v_320x104_v15_IPB",
    "stress_640x480_v15_IPB", "stress_320x240_v13_IPB", "stress_328x248_v15_IPB", "stress_64x48_v15_IPB", "cap16_320x240_v15_I",
    "cap17_320x240_v15_I", "portrait_240x320_v15_IPB", "portrait_480x640_v13_IPB", "portrait_stress_240x320_v15_IPB",
    "portrait_small_64x96_v15_IPB"])
def test_emulated_pipeline_matches_golden(emul_lib, golden, name):
    case = golden[name]
    got = list(emul_decode(emul_lib, synth.generate(**case["args"])))
    assert all(err == 0 for _, _, err in got)
    assert [md5(yuv) for _, yuv, _ in got] == case["md5"]


def test_emulated_pipeline_matches_port_on_fresh_seeds(emul_lib, oracle):
    for seed in range(3):
        data = synth.generate(320, 240, 15 if seed % 2 else 13, "IPBBPBB", 1, seed=7000 + seed, profile=seed % 2)
        want = [yuv for _, _, _, yuv in oracle.PortDecoder(data).frames()]
        got = [yuv for _, yuv, _ in emul_decode(emul_lib, data)]
        assert got == want


@pytest.mark.parametrize("name", ["cfg3_640x480_v15_IPB", "cfg4_320x240_v13_IPB", "cfg5_stream511", "realistic_640x480_v15_IPB",
                                  "wide_1024x576_v13_IPB", "hd_1280x720_v15_IPB", "small_64x48_v13_IPB", "stress_640x480_v15_IPB",
                                  "stress_320x240_v13_IPB"])
@pytest.mark.parametrize("kernel", ["sweep", "row"])
def test_emulated_pipeline_kernels_match_golden(emul_lib, golden, name, kernel):
    """The shared-memory pipeline kernels (sweep.cu, row.cu): plan, slots, rings, lists and tasks run serially on the CPU.
    Three slots / rows of look-ahead and then one, so that both a full and an empty pipeline are walked."""
    case = golden[name]
    for look in (6, 1):
        stats = {}
        kw = dict(sweep=(1, 232448, look)) if kernel == "sweep" else dict(row=(232448, look))
        got = list(emul_decode(emul_lib, synth.generate(**case["args"]), stats=stats, **kw))
        assert [md5(yuv) for _, yuv, _ in got] == case["md5"], (kernel, look)
        if kernel == "row" and case["args"]["width"] % 32 == 0 and case["args"]["width"] <= 1024:
            if "stress" in name:      # a row with more chunk descriptors than a slot holds leaves the picture to the band kernel
                assert stats.get("row", 0) >= len(got) // 2, stats
            else:
                assert stats.get("row", 0) == len(got), stats      # every picture of these streams is served by the row kernel


@pytest.mark.parametrize("args", [
    dict(width=320, height=240, version=15, gop="IPBBPBB", n_gops=1, seed=811, profile=0),
    dict(width=640, height=480, version=13, gop="IPB", n_gops=1, seed=812, profile=1),
    dict(width=328, height=248, version=15, gop="IPB", n_gops=1, seed=813, profile=0)])
def test_split_schedule_builds_the_same_symbol_buffer(emul_lib, args):
    """The GPU build of the entropy stage resolves vectors serially and schedules records row by
    row (pb_mvs + schedule_rows); a host thread runs the fused pb_pass2.  One lane must give the
    same bytes either way."""
    version, w, h, recs = demux(synth.generate(**args))
    seqs = [emul_lib.h4e_seq_create(w, h, 2, 2, 1 if version == 15 else 0) for _ in range(2)]
    emul_lib.h4e_seq_set_split_schedule(seqs[1], 1)
    for ty, _, pic in recs:
        buf = bytes(pic) + b"\0" * 8
        blobs = []
        for seq in seqs:
            n = emul_lib.h4e_parse_begin(seq, ty, buf, len(pic))
            assert n
            blob = np.zeros(n, np.uint8)
            assert emul_lib.h4e_parse_finish(seq, blob.ctypes.data) == 0
            blobs.append(blob.tobytes())
        assert blobs[0] == blobs[1]
    for seq in seqs:
        emul_lib.h4e_seq_destroy(seq)


def test_split_schedule_equals_serial_walk_on_damaged_pictures(emul_lib):
    """Same comparison on bit-flipped and truncated pictures: exhausted sections, zero-length runs,
    clamped types and poisoned vectors must come out of the prefix-sum forms exactly as out of
    the serial loops -- same bytes, same error bits."""
    rng = np.random.default_rng(5)
    checked = flagged = 0
    for seed, profile in ((31, 0), (32, 1), (33, 1)):
        version, w, h, recs = demux(synth.generate(320, 240, 15, "IPBPB", 1, seed=seed, profile=profile))
        for trial in range(10):
            seqs = [emul_lib.h4e_seq_create(w, h, 2, 2, 1) for _ in range(2)]
            emul_lib.h4e_seq_set_split_schedule(seqs[1], 1)
            for ty, _, pic in recs:
                bad = bytearray(pic)
                if trial % 5 == 4:
                    bad = bad[: max(80, len(bad) * (1 + trial) // 12)]
                else:
                    for _ in range(1 + 6 * (trial % 4)):
                        bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
                buf = bytes(bad) + b"\0" * 24
                out = []
                for seq in seqs:
                    n = emul_lib.h4e_parse_begin(seq, ty, buf, len(bad))
                    blob = np.zeros(max(n, 1), np.uint8)
                    err = emul_lib.h4e_parse_finish(seq, blob.ctypes.data) if n else -1
                    out.append((n, err, blob.tobytes()))
                assert out[0] == out[1], (seed, trial, ty)
                checked += 1
                flagged += out[0][1] != 0
            for seq in seqs:
                emul_lib.h4e_seq_destroy(seq)
    assert checked == 150 and flagged > 20


def test_truncated_and_corrupt_pictures_raise_error_bits_not_crashes(emul_lib):
    """The reference has no input validation; the host stage must stay in bounds."""
    data = synth.generate(320, 240, 15, "IPB", 1, seed=42, profile=0)
    version, w, h, recs = demux(data)
    rng = np.random.default_rng(0)
    for ty, _, pic in recs:
        for trial in range(6):
            seq = emul_lib.h4e_seq_create(w, h, 2, 2, 1)
            bad = bytearray(pic)
            if trial < 3:
                bad = bad[: max(8, len(bad) * (trial + 1) // 5)]          # truncation
            else:
                for _ in range(40):                                        # bit flips
                    bad[rng.integers(0, len(bad))] ^= 1 << rng.integers(0, 8)
            buf = bytes(bad) + b"\0" * 8
            n = emul_lib.h4e_parse_begin(seq, ty, buf, len(bad))
            if n:
                blob = np.zeros(n, np.uint8)
                err = emul_lib.h4e_parse_finish(seq, blob.ctypes.data)
                if trial < 3:
                    assert err != 0
            emul_lib.h4e_seq_destroy(seq)


def test_unsupported_geometry_is_rejected(emul_lib):
    assert not emul_lib.h4e_seq_create(322, 240, 2, 2, 1)     # not a multiple of 8
    assert not emul_lib.h4e_seq_create(320, 240, 1, 1, 1)     # 4:4:4
    seq = emul_lib.h4e_seq_create(240, 320, 2, 2, 1)          # portrait: decoded like the reference does (38 x 70 nest)
    assert seq
    emul_lib.h4e_seq_destroy(seq)


def test_kernel_leaf_arithmetic_equals_reference_leaves(emul_lib, oracle):
    """The packed 16-bit weighted fill and the packed half-sample filters of recon_core.h (the
    code the CUDA kernels run) against the reference's own WeightImBlock (h4m:299-383) and
    _MotionComp (h4m:1242-1294): extremes, the sat_mean8 wrap, and 20 000 random neighbourhoods."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    ref = ctypes.CDLL(oracle.REF_LIB)
    rng = np.random.default_rng(7)
    want = (ctypes.c_uint8 * 16)()
    got = (ctypes.c_uint8 * 16)()
    ext = [0, 1, 2, 127, 128, 254, 255]
    cases = [(v, t, b, l, r) for v in ext for t in ext for b in ext for l in (0, 128, 255) for r in (0, 3, 255)]
    cases += [tuple(int(x) for x in rng.integers(0, 256, 5)) for _ in range(20000)]
    for V, T, B, L, R in cases:
        ref.ref_WeightImBlock(want, 4, V, T, B, L, R)
        emul_lib.emul_weighted(got, V, T, B, L, R)
        assert bytes(got) == bytes(want), (V, T, B, L, R)
    src = rng.integers(0, 256, (12, 16), dtype=np.uint8)
    src[:4] = 255
    src[4:6] = 0
    for oy in range(6):
        for ox in range(8):
            for hx in (0, 1):
                for hy in (0, 1):
                    p = src.ctypes.data + oy * 16 + ox
                    ref.ref_MotionComp4x4(want, 4, ctypes.c_void_p(p), 16, hx, hy)
                    emul_lib.emul_predict(got, ctypes.c_void_p(p), 16, hx, hy)
                    assert bytes(got) == bytes(want), (oy, ox, hx, hy)
                    emul_lib.emul_predict_diag(got, ctypes.c_void_p(p), 16, hx, hy)
                    assert bytes(got) == bytes(want), ("diag form", oy, ox, hx, hy)

"""The CPU checkers themselves: oracle port vs golden vectors (always) and vs the compiled
reference (where oracle/_ref exists).  No GPU."""
import ctypes
import hashlib

import numpy as np
import pytest

from hvqm4_b200 import synth


def test_generator_is_deterministic(golden):
    for name, case in golden.items():
        data = synth.generate(**case["args"])
        assert hashlib.sha256(data).hexdigest() == case["stream_sha256"], name


@pytest.mark.parametrize("name", [
    "cfg1_320x240_v15_I30", "cfg2_640x480_v15_IP15", "cfg3_640x480_v15_IPB", "cfg4_320x240_v13_IPB",
    "cfg5_stream0", "cfg5_stream1023", "realistic_640x480_v15_IPB", "realistic_320x240_v13_IPB",
    "min_280x152_v15_IPB", "ragged_328x248_v15_IPB", "wide_1024x576_v13_IPB"])
def test_port_matches_golden(oracle, golden, name):
    case = golden[name]
    got = oracle.PortDecoder.md5s(synth.generate(**case["args"]))
    assert [t for t, _, _ in got] == case["frame_types"]
    assert [d for _, d, _ in got] == case["disp_ids"]
    assert [m for _, _, m in got] == case["md5"]


def test_reference_build_matches_golden(oracle, golden):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    for name in ("cfg1_320x240_v15_I30", "cfg4_320x240_v13_IPB", "cfg5_stream1", "stress_320x240_v13_IPB", "cap17_320x240_v15_I"):
        case = golden[name]
        got = oracle.RefDecoder.md5s(synth.generate(**case["args"]))
        assert [m for _, _, m in got] == case["md5"], name


def test_port_matches_reference_maps_and_sections(oracle):
    """Beyond pixels: block maps, nest and per-section consumption agree picture by picture."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    for seed, (w, h, v, gop, prof) in enumerate([(320, 240, 15, "IPBBPBB", 0), (640, 480, 13, "IPBB", 0), (328, 248, 15, "IPB", 1),
                                                 (320, 240, 15, "IPBBPBB", 2), (640, 480, 13, "IPBB", 2), (320, 240, 15, "I", 4),
                                                 (240, 320, 15, "IPBBPBB", 0), (96, 160, 13, "IPB", 2)]):
        data = synth.generate(w, h, v, gop, 2, seed=900 + seed, profile=prof)
        a, b = oracle.RefDecoder(data), oracle.PortDecoder(data)
        for fa, fb in zip(a.frames(), b.frames()):
            assert fa == fb
            for p in range(3):
                assert a.get_map(p) == b.get_map(p)
            assert a.get_nest() == b.get_nest()
            ca, sa = a.section_usage()
            cb, sb = b.section_usage()
            assert (ca, sa) == (cb, sb)
            # generator contract: every present section is consumed to exactly its declared size
            assert ca == sa


def test_reference_constants_and_sat_mean8_quirk(oracle):
    """SURVEY section 10 items 1 and 5, pinned on the reference's own leaf functions."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    lib = ctypes.CDLL(oracle.REF_LIB)
    div = (ctypes.c_int32 * 16)()
    mcdiv = (ctypes.c_int32 * 512)()
    lib.ref_tables(div, mcdiv)
    assert list(div) == [0] + [0x1000 // (i * 16) * 16 for i in range(1, 16)]
    assert list(mcdiv) == [0] + [0x1000 // i for i in range(1, 512)]
    out = (ctypes.c_uint8 * 16)()

    def closed_form(V, T, B, L, R):
        rt = [2 * T - B - V, V - B, V - T, 2 * B - T - V]
        ct = [2 * L - R - V, V - R, V - L, 2 * R - L - V]
        res = []
        for r in range(4):
            for c in range(4):
                s = 8 * V + rt[r] + ct[c]
                q = ((s + 4) & 0xFFFFFFFF) // 8            # unsigned division, h4m:293-296
                res.append(min(q, 255))
        return res

    rng = np.random.default_rng(1)
    cases = [(0, 0, 255, 0, 255), (0, 0, 2, 0, 1), (255, 0, 0, 0, 0), (0, 255, 255, 255, 255), (3, 9, 20, 1, 30)]
    cases += [tuple(int(x) for x in rng.integers(0, 256, 5)) for _ in range(500)]
    for V, T, B, L, R in cases:
        lib.ref_WeightImBlock(out, 4, V, T, B, L, R)
        assert list(out) == closed_form(V, T, B, L, R), (V, T, B, L, R)
    # the quirk itself: 10V-B-R = -510 -> 255 (not 0); 10V-B-R = -3 -> 0
    lib.ref_WeightImBlock(out, 4, 0, 0, 255, 0, 255)
    assert out[5] == 255
    lib.ref_WeightImBlock(out, 4, 0, 0, 2, 0, 1)
    assert out[5] == 0


def test_rgb_conversion_port_equals_reference_on_every_triple(oracle):
    """dumpRGB (h4m:895-926) on pictures that together hold all 2^24 (y, u, v) triples: the port
    must equal the reference build (which also pins that FMA contraction cannot matter here)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    from tests.h4m_util import all_triples_pictures
    n = 0
    for yuv, w, h in all_triples_pictures():
        assert oracle.RefDecoder.yuv_to_rgb(yuv, w, h) == oracle.PortDecoder.yuv_to_rgb(yuv, w, h)
        n += 1
    assert n == 8


def test_audio_port_equals_reference_decode_audio(oracle):
    """IMA-ADPCM track (decode_audio, h4m:185-258): the port against the reference's own function on
    random records, mono and stereo, seeded first record then two continuation records."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    import random
    rng = random.Random(11)
    for ch in (1, 2):
        for _ in range(100):
            s_ref, s_port = [0] * (2 * ch), [0] * (2 * ch)
            for k in range(3):
                n = rng.randint(1, 400)
                data = bytearray(rng.randrange(256) for _ in range(2 * ch + n * ch))
                if k == 0:
                    for c in range(ch):
                        data[2 * c + 1] = (data[2 * c + 1] & 0x80) | rng.randint(0, 88)
                a = oracle.RefDecoder.decode_audio(s_ref, ch, k == 0, n, bytes(data))
                b = oracle.PortDecoder.decode_audio(s_port, ch, k == 0, n, bytes(data))
                assert a == b and s_ref == s_port and len(a) == n * ch


def test_spliced_audio_container_is_what_the_reference_walks(oracle):
    """tests/h4m_util.with_audio builds files the reference's container walk accepts: its video
    frames decode to the same pictures as without the audio records."""
    from hvqm4_b200 import synth
    from tests.h4m_util import with_audio
    data = synth.generate(320, 240, 15, "IPBB", 2, seed=77, profile=1)
    spliced, records = with_audio(data, 2, 3, 200, seed=5)
    assert len(records) == 6 and [r[1] for r in records] == [True, False, False] * 2
    checker = oracle.RefDecoder if oracle.have_ref() else oracle.PortDecoder
    assert [f[3] for f in checker(spliced).frames()] == [f[3] for f in checker(data).frames()]

"""Host serial stage (entropy.c) + the kernel's block arithmetic and work order, run on the CPU
through tests/emul, against the golden vectors and the oracle port.  No GPU."""
import ctypes

import numpy as np
import pytest

from hvqm4_b200 import synth
from tests.h4m_util import demux, emul_decode, md5


@pytest.mark.parametrize("name", [
    "cfg1_320x240_v15_I30", "cfg3_640x480_v15_IPB", "cfg4_320x240_v13_IPB", "cfg5_stream511",
    "realistic_640x480_v15_IPB", "min_280x152_v15_IPB", "ragged_328x248_v15_IPB", "wide_1024x576_v13_IPB"])
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
    assert not emul_lib.h4e_seq_create(240, 320, 2, 2, 1)     # portrait (untested upstream, README:23)

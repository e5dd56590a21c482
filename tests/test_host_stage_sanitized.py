"""The serial stage (hvqm4_b200/csrc/entropy.c) under AddressSanitizer + UndefinedBehaviorSanitizer.

The reference reads its input without bounds (h4m:552-602) and lets malformed streams run into undefined
behaviour; this stage claims the opposite: no read past the picture bytes it was given
(not even the three slack bytes of h4m:2080-2082), no write past the symbol buffer size it announced, no
undefined arithmetic, error bits instead.  tests/san/entropy_san.c hands every picture over in a heap block
of exactly its size and the blob in a heap block of exactly `h4e_parse_begin`'s answer, intact and damaged
(byte flips, corrupted header / section table, truncation), in the host pass structure and in the pass
structure of the GPU build (the device parser is the same source file).  Every parsed picture is then
reconstructed by tests/emul -- the block functions and addressing of the CUDA kernels (recon_core.h) --
into surfaces of exactly frame bytes + 64: a vector, window or record of a damaged picture that left a
surface or the blob would be a report.  compute-sanitizer is not available on the GPU pool, so this is
the sanitizer coverage the GPU parser and the kernels' arithmetic get."""
import os
import shutil
import subprocess

import pytest

from hvqm4_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "san", "entropy_san.c")
EXE = os.path.join(ROOT, "tests", "san", "entropy_san.bin")
CSRC = os.path.join(ROOT, "hvqm4_b200", "csrc")


@pytest.fixture(scope="module")
def harness():
    if not shutil.which("gcc") or not shutil.which("g++"):
        pytest.skip("gcc / g++ not available")
    san = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer"]
    inc = ["-I", CSRC, "-I", os.path.join(ROOT, "include")]
    obj = lambda name: os.path.join(ROOT, "tests", "san", name + ".san.o")
    steps = [
        ["gcc", "-std=gnu11"] + san + inc + ["-c", os.path.join(CSRC, "entropy.c"), "-o", obj("entropy")],
        ["gcc", "-std=gnu11"] + san + inc + ["-c", SRC, "-o", obj("harness")],
        # the kernels' block functions and addressing (recon_core.h) as the serial emulation runs them
        ["g++", "-std=c++17", "-Wno-unknown-pragmas"] + san + ["-c", os.path.join(ROOT, "tests", "emul", "recon_emul.cpp"), "-o", obj("recon")],
        ["g++"] + san + [obj("entropy"), obj("harness"), obj("recon"), "-o", EXE],
    ]
    for cmd in steps:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 and ("asan" in r.stderr or "ubsan" in r.stderr or "sanitize" in r.stderr):
            pytest.skip("sanitizer runtimes not installed: " + r.stderr.strip()[-200:])
        assert r.returncode == 0, r.stderr[-2000:]
    return EXE


CASES = [
    # width, height, version, decode order, profile (0 dense, 1 realistic)
    (640, 480, 15, "IPBBPB", 0),
    (640, 480, 15, "IPBBPB", 1),
    (320, 240, 13, "IPBBPBB", 0),
    (328, 248, 15, "IPBB", 0),          # ragged: sizes that are no multiple of the segment width
    (16, 16, 15, "IPB", 1),             # smaller than the nest (MakeNest mirror / zero fill, h4m:1166-1239)
]


@pytest.mark.parametrize("split", [0, 1], ids=["host-passes", "gpu-passes"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}x{c[1]}_v{c[2]}_{c[3]}_p{c[4]}")
def test_serial_stage_is_clean_under_asan_and_ubsan(harness, tmp_path, case, split):
    w, h, ver, gop, profile = case
    data = synth.generate(w, h, ver, gop, 1, seed=7300 + 10 * CASES.index(case), profile=profile)
    _, frames = api.parse_file(data)
    stream, listing = tmp_path / "s.h4m", tmp_path / "s.txt"
    stream.write_bytes(data)
    listing.write_text("".join(f"{f.offset} {f.bytes} {f.frame_type}\n" for f in frames))
    rounds = 25
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([harness, str(stream), str(listing), str(w), str(h), str(int(ver == 15)), str(rounds), str(31 + split), str(split)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-3000:]
    words = r.stdout.split()
    parsed, flagged, refused, reconstructed = int(words[0]), int(words[2]), int(words[4]), int(words[6])
    # round 0 is the intact stream: no error bits there, so at most (rounds - 1) x pictures can be flagged
    assert parsed + refused == rounds * len(frames)
    assert flagged <= (rounds - 1) * len(frames)
    assert reconstructed == parsed

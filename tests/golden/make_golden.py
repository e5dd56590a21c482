"""Generates tests/golden/golden_md5.json.

Run in the authoring container, where /root/reference exists: the UNMODIFIED reference
decoder (compiled by oracle/Makefile into oracle/_ref/libhvqm4_ref.so) decodes the
seeded synthetic streams below and the per-frame MD5 of its planar YUV `present` buffer
is recorded.  The streams themselves are not stored: tools/h4mgen.c regenerates them
bit-identically from (geometry, gop, seed, profile); their SHA-256 is stored to catch
generator drift.  The reference ships no golden vectors of its own.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hvqm4_b200 import synth  # noqa: E402
from oracle import bindings  # noqa: E402

# name -> generator arguments.  The first five are BASELINE.json's configs (config 5 is
# represented by four of its 1024 seeds); the rest widen coverage.
CASES = {
    "cfg1_320x240_v15_I30": dict(width=320, height=240, version=15, gop="I" * 30, n_gops=1, seed=101, profile=0),
    "cfg2_640x480_v15_IP15": dict(width=640, height=480, version=15, gop="I" + "P" * 14, n_gops=2, seed=102, profile=0),
    "cfg3_640x480_v15_IPB": dict(width=640, height=480, version=15, gop="I" + "PBB" * 5, n_gops=2, seed=103, profile=0),
    "cfg4_320x240_v13_IPB": dict(width=320, height=240, version=13, gop="I" + "PBB" * 5, n_gops=2, seed=104, profile=0),
    "cfg5_stream0": dict(width=640, height=480, version=15, gop="I" + "PBB" * 5, n_gops=1, seed=5000, profile=0),
    "cfg5_stream1": dict(width=640, height=480, version=15, gop="I" + "PBB" * 5, n_gops=1, seed=5001, profile=0),
    "cfg5_stream511": dict(width=640, height=480, version=15, gop="I" + "PBB" * 5, n_gops=1, seed=5511, profile=0),
    "cfg5_stream1023": dict(width=640, height=480, version=15, gop="I" + "PBB" * 5, n_gops=1, seed=6023, profile=0),
    "realistic_640x480_v15_IPB": dict(width=640, height=480, version=15, gop="I" + "PBB" * 5, n_gops=2, seed=201, profile=1),
    "realistic_320x240_v13_IPB": dict(width=320, height=240, version=13, gop="I" + "PBB" * 3, n_gops=2, seed=202, profile=1),
    "min_280x152_v15_IPB": dict(width=280, height=152, version=15, gop="IPBB", n_gops=2, seed=203, profile=0),
    "ragged_328x248_v15_IPB": dict(width=328, height=248, version=15, gop="IPBBP", n_gops=1, seed=204, profile=0),
    "wide_1024x576_v13_IPB": dict(width=1024, height=576, version=13, gop="IPBB", n_gops=1, seed=205, profile=0),
    # 160 macroblocks per row: more than one column tile of the band kernel (recon.cu kTileMcbs)
    "hd_1280x720_v15_IPB": dict(width=1280, height=720, version=15, gop="IPBB", n_gops=1, seed=206, profile=0),
    # 512 macroblocks per row: four column tiles of the band kernel, queues at their fixed maximum size
    "uhd_4096x2160_v15_IPB": dict(width=4096, height=2160, version=15, gop="IPB", n_gops=1, seed=211, profile=0),
    # pictures smaller than the 70x38-block nest: MakeNest mirrors, then zero-fills (h4m:1173-1203)
    "tiny_16x16_v15_IPB": dict(width=16, height=16, version=15, gop="IPBBPB", n_gops=2, seed=207, profile=0),
    "small_64x48_v13_IPB": dict(width=64, height=48, version=13, gop="IPBBPB", n_gops=2, seed=208, profile=0),
    "mirror_h_200x152_v15_IPB": dict(width=200, height=152, version=15, gop="IPBB", n_gops=2, seed=209, profile=0),
    "mirror_v_320x104_v15_IPB": dict(width=320, height=104, version=15, gop="IPBB", n_gops=2, seed=210, profile=0),
    # the generator's stress profile: what the reference accepts and an encoder rarely emits (block types 7 and 9..255 in
    # I-picture luma = that many bases, nibbles 7 and 9..15 elsewhere -- up to 14 bases per predicted block --, scale
    # symbols to 255, dc_shift 0..3, unk_shift 6..12, escape chains, run counts >= 255, rb 0..3; h4m:1358-1459, 654-677)
    "stress_640x480_v15_IPB": dict(width=640, height=480, version=15, gop="I" + "PBB" * 3, n_gops=2, seed=301, profile=2),
    "stress_320x240_v13_IPB": dict(width=320, height=240, version=13, gop="I" + "PBB" * 5, n_gops=2, seed=302, profile=2),
    "stress_328x248_v15_IPB": dict(width=328, height=248, version=15, gop="IPBBPB", n_gops=2, seed=303, profile=2),
    "stress_64x48_v15_IPB": dict(width=64, height=48, version=15, gop="IPBBPB", n_gops=3, seed=304, profile=2),
    # every luma block of the I pictures carries 16 / 17 bases: at / beyond the symbol capacity of the GPU entropy stage
    "cap16_320x240_v15_I": dict(width=320, height=240, version=15, gop="II", n_gops=1, seed=305, profile=3),
    "cap17_320x240_v15_I": dict(width=320, height=240, version=15, gop="II", n_gops=1, seed=306, profile=4),
    # portrait pictures: the nest is 38 x 70, the axes of the basis descriptors swap and the window of predicted-AOT
    # macroblocks sits at (-16, -32) (h4m:700-711, 743-754, 965-975, 1865-1868); upstream calls the orientation untested
    # (README:23): what counts here is what the reference build does
    "portrait_240x320_v15_IPB": dict(width=240, height=320, version=15, gop="I" + "PBB" * 3, n_gops=2, seed=401, profile=0),
    "portrait_480x640_v13_IPB": dict(width=480, height=640, version=13, gop="IPBBPBB", n_gops=1, seed=402, profile=0),
    "portrait_stress_240x320_v15_IPB": dict(width=240, height=320, version=15, gop="IPBBPB", n_gops=2, seed=403, profile=2),
    "portrait_small_64x96_v15_IPB": dict(width=64, height=96, version=15, gop="IPBBPB", n_gops=2, seed=404, profile=0),
}


def main():
    bindings.build(ref=True, port=True)
    if not bindings.have_ref():
        raise SystemExit("oracle/_ref is not built and /root/reference is absent: cannot make golden vectors")
    out = {}
    for name, args in CASES.items():
        data = synth.generate(**args)
        frames = bindings.RefDecoder.md5s(data)
        out[name] = {
            "args": args,
            "stream_sha256": hashlib.sha256(data).hexdigest(),
            "frame_types": [t for t, _, _ in frames],
            "disp_ids": [d for _, d, _ in frames],
            "md5": [m for _, _, m in frames],
        }
        print(name, len(frames), "frames")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_md5.json")
    with open(path, "w") as f:
        json.dump({"generator": "tools/h4mgen.c", "oracle": "oracle/_ref (unmodified reference)", "cases": out}, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()

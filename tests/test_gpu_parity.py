"""Parity proper: the CUDA path, called through the C ABI (include/hvqm4.h), against the
oracle and the golden vectors.  Bit-exact: every comparison is on bytes / MD5s."""
import ctypes
import hashlib

import numpy as np
import pytest

from hvqm4_b200 import synth
from tests.h4m_util import md5

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["two_kernels", "band_kernel", "band_tile", "sweep_kernel", "row_kernel"])
def recon_mode(request, native_lib):
    """Every parity test runs under all three reconstruction schedules (include/hvqm4.h HVQM4SetReconMode)."""
    native_lib.set_recon_mode({"two_kernels": -1, "band_kernel": 4, "band_tile": 7, "sweep_kernel": 5, "row_kernel": 6}[request.param])
    yield request.param
    native_lib.set_recon_mode(0)

SDK_CASES = ["cfg1_320x240_v15_I30", "cfg2_640x480_v15_IP15", "cfg3_640x480_v15_IPB", "cfg4_320x240_v13_IPB",
             "realistic_640x480_v15_IPB", "realistic_320x240_v13_IPB", "min_280x152_v15_IPB",
             "ragged_328x248_v15_IPB", "wide_1024x576_v13_IPB", "hd_1280x720_v15_IPB",
             "tiny_16x16_v15_IPB", "small_64x48_v13_IPB", "mirror_h_200x152_v15_IPB", "mirror_v_320x104_v15_IPB",
             "uhd_4096x2160_v15_IPB",
             "stress_640x480_v15_IPB", "stress_320x240_v13_IPB", "stress_328x248_v15_IPB", "stress_64x48_v15_IPB",
             "cap16_320x240_v15_I", "cap17_320x240_v15_I",
             "portrait_240x320_v15_IPB", "portrait_480x640_v13_IPB", "portrait_stress_240x320_v15_IPB", "portrait_small_64x96_v15_IPB"]


@pytest.mark.parametrize("name", SDK_CASES)
def test_sdk_entry_points_match_golden(native_lib, golden, name, recon_mode):
    """HVQM4InitSeqObj/BuffSize/SetBuffer/DecodeIpic/Ppic/Bpic with host buffers, driven like the
    reference's decode_video(): per-frame MD5 must equal the reference decoder's."""
    case = golden[name]
    player = native_lib.Player(synth.generate(**case["args"]))
    got = [(t, d, md5(yuv)) for t, d, yuv in player]
    player.close()
    assert [t for t, _, _ in got] == case["frame_types"]
    assert [d for _, d, _ in got] == case["disp_ids"]
    assert [m for _, _, m in got] == case["md5"]


def test_sdk_entry_points_match_oracle_port_bytes(native_lib, oracle, recon_mode):
    """Same seeds, fresh streams, byte comparison against the oracle port (reports the first
    differing block instead of just a hash)."""
    for seed in range(4):
        data = synth.generate(320, 240, 15 if seed & 1 else 13, "IPBBPBB", 1, seed=8100 + seed, profile=seed >> 1)
        want = [yuv for _, _, _, yuv in oracle.PortDecoder(data).frames()]
        player = native_lib.Player(data)
        for i, (t, _, yuv) in enumerate(player):
            if yuv != want[i]:
                a, b = np.frombuffer(yuv, np.uint8), np.frombuffer(want[i], np.uint8)
                idx = np.nonzero(a != b)[0]
                pytest.fail(f"seed {seed} frame {i} type {t:#x}: {len(idx)} bytes differ, first at {idx[:8]}")
        player.close()


def test_batch_runtime_matches_golden_cfg5(native_lib, golden, recon_mode):
    """Config 5 in miniature: independent streams batched per launch; every stream's every
    frame must match the reference MD5."""
    names = ["cfg5_stream0", "cfg5_stream1", "cfg5_stream511", "cfg5_stream1023"]
    files = [synth.generate(**golden[n]["args"]) for n in names]
    step = 0
    for frames in native_lib.decode_streams(files, host_threads=4):
        for n, (t, d, yuv) in zip(names, frames):
            assert t == golden[n]["frame_types"][step]
            assert md5(yuv) == golden[n]["md5"][step], (n, step)
        step += 1
    assert step == 16


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_band_kernel_register_budgets_match_golden(native_lib, golden, mode):
    """The fused band kernel exists with 2, 3 and 4 CTAs per SM (92 / 80 / 64 registers; HVQM4SetReconMode 2..4,
    1 = by grid size, which large batches resolve to 3): same pictures from every instantiation."""
    names = ["cfg5_stream0", "cfg5_stream1", "cfg5_stream511", "cfg5_stream1023"]
    files = [synth.generate(**golden[n]["args"]) for n in names]
    native_lib.set_recon_mode(mode)
    try:
        before = native_lib.lib().HVQM4KernelLaunches()
        for step, frames in enumerate(native_lib.decode_streams(files, host_threads=4)):
            for n, (_, _, yuv) in zip(names, frames):
                assert md5(yuv) == golden[n]["md5"][step], (n, step, mode)
        assert native_lib.lib().HVQM4KernelLaunches() - before == 16        # one band kernel per step
    finally:
        native_lib.set_recon_mode(0)


def test_batch_runtime_many_streams_vs_oracle(native_lib, oracle, recon_mode):
    """48 streams with distinct seeds in one batch (partial warps, several CTAs per SM)."""
    n = 48
    files = [synth.generate(320, 240, 15, "IPBBP", 1, seed=9000 + i, profile=i % 2) for i in range(n)]
    want = [[yuv for _, _, _, yuv in oracle.PortDecoder(f).frames()] for f in files]
    for step, frames in enumerate(native_lib.decode_streams(files)):
        for i, (_, _, yuv) in enumerate(frames):
            assert yuv == want[i][step], (i, step)


@pytest.mark.parametrize("profile", [0, 1])
def test_sweep_kernel_full_size_vs_oracle(native_lib, oracle, profile):
    """The sweep kernel at BASELINE.json's size, where the reference windows fill shared memory (dense content: vectors
    of +-64 rows, B pictures swept twice): 12 streams of 640x480 against the oracle, frame by frame; the kernel must
    actually have run and none of its CTAs may have given up on its copy pipeline."""
    n = 12
    files = [synth.generate(640, 480, 15, "IPBBPBB", 1, seed=7100 + i, profile=profile) for i in range(n)]
    want = [[yuv for _, _, _, yuv in oracle.PortDecoder(f).frames()] for f in files]
    native_lib.set_recon_mode(5)
    try:
        before = native_lib.sweep_launches()
        for step, frames in enumerate(native_lib.decode_streams(files)):
            for i, (_, _, yuv) in enumerate(frames):
                assert yuv == want[i][step], (i, step)
        assert native_lib.sweep_launches() - before == 7
        assert native_lib.sweep_errors() == 0
    finally:
        native_lib.set_recon_mode(0)


@pytest.mark.parametrize("profile", [0, 1])
def test_row_kernel_full_size_vs_oracle(native_lib, oracle, profile):
    """The row kernel at BASELINE.json's size: 160 streams of 640x480 (more than one picture per CTA, pictures split
    between CTAs) against the oracle, frame by frame; the kernel must actually have run and none of its CTAs may have
    given up on its copy pipeline."""
    distinct = 10
    files = [synth.generate(640, 480, 15, "IPBBPBB", 1, seed=7200 + i, profile=profile) for i in range(distinct)]
    want = [[yuv for _, _, _, yuv in oracle.PortDecoder(f).frames()] for f in files]
    native_lib.set_recon_mode(6)
    try:
        before = native_lib.row_launches()
        for step, frames in enumerate(native_lib.decode_streams([files[i % distinct] for i in range(160)])):
            for i, (_, _, yuv) in enumerate(frames):
                assert yuv == want[i % distinct][step], (i, step)
        assert native_lib.row_launches() - before == 7
        assert native_lib.row_errors() == 0
    finally:
        native_lib.set_recon_mode(0)


def test_record_replay_is_idempotent_and_counts_launches(native_lib, golden):
    """Reconstruction-only replay (what bench.py times) must reproduce the same pictures."""
    native_lib.set_recon_mode(-1)
    case = golden["cfg5_stream1"]
    data = synth.generate(**case["args"])
    info, frames = native_lib.parse_file(data)
    buf = ctypes.create_string_buffer(data, len(data) + 8)
    base = ctypes.addressof(buf)
    n = 3
    batch = native_lib.Batch(n, info.width, info.height, info.version, host_threads=2)
    batch.record(True)
    for fr in frames:
        batch.decode(list(range(n)), [fr.frame_type] * n, [base + fr.offset] * n, [fr.bytes] * n)
    batch.sync()
    batch.record(False)
    last = md5(batch.read_frame(1))
    before = native_lib.kernel_launches()
    ms = batch.replay(2)
    assert ms > 0
    # every dense picture has records: one map launch + one record launch per step
    assert native_lib.kernel_launches() - before == 2 * 2 * len(frames)
    batch.sync()
    assert md5(batch.read_frame(1)) == last == case["md5"][-1]
    st = batch.stats()
    assert st["pictures"] == n * len(frames) and st["algorithmic_bytes"] > st["pictures"] * batch.frame_bytes
    batch.close()
    native_lib.set_recon_mode(0)


def test_full_size_properties_640x480(native_lib, recon_mode):
    """Size-independent properties at BASELINE.json's full size (no oracle needed):
    (1) a stream decoded alone and inside a batch of different streams gives identical frames;
    (2) decoding the same stream twice is deterministic;
    (3) 1.3 and 1.5 agree on I pictures and differ on P/B pictures (chroma phase rule,
        /root/reference/h4m_audio_decode.c:1337-1343)."""
    gop = "I" + "PBB" * 3
    a15 = synth.generate(640, 480, 15, gop, 1, seed=77, profile=0)
    a13 = synth.generate(640, 480, 13, gop, 1, seed=77, profile=0)
    others = [synth.generate(640, 480, 15, gop, 1, seed=500 + i, profile=0) for i in range(5)]
    alone = [yuv for _, _, yuv in native_lib.Player(a15)]
    again = [yuv for _, _, yuv in native_lib.Player(a15)]
    assert alone == again
    in_batch = [frames[2][2] for frames in native_lib.decode_streams(others[:2] + [a15] + others[2:])]
    assert in_batch == alone
    v13 = [yuv for _, _, yuv in native_lib.Player(a13)]
    assert v13[0] == alone[0]
    assert any(x != y for x, y in zip(v13[1:], alone[1:]))


def test_device_pointer_mode_skips_copies(native_lib, golden):
    """Zero-copy mode of the SDK calls: present/past/future may be device pointers."""
    import torch
    case = golden["cfg4_320x240_v13_IPB"]
    data = synth.generate(**case["args"])
    info, frames = native_lib.parse_file(data)
    dec = native_lib.SeqDecoder(info.width, info.height, info.version)
    fb = dec.frame_bytes
    surf = [torch.zeros(fb + 64, dtype=torch.uint8, device="cuda") for _ in range(3)]
    past, present, future = 0, 1, 2
    for i, fr in enumerate(frames[:8]):
        if fr.frame_type != 0x30:
            past, future = future, past
        dec.decode(fr.frame_type, data[fr.offset:fr.offset + fr.bytes], surf[present].data_ptr(), surf[past].data_ptr(), surf[future].data_ptr())
        got = surf[present][:fb].cpu().numpy().tobytes()
        assert md5(got) == case["md5"][i], i
        if fr.frame_type != 0x30:
            present, future = future, present
    dec.close()


def test_corrupt_pictures_do_not_fault_the_gpu(native_lib):
    """Bit-flipped and truncated pictures through the SDK entry points: the host stage must flag
    them (HVQM4GetLastError) and the kernels must stay inside the surfaces (no CUDA error, and the
    library keeps decoding valid pictures afterwards)."""
    import numpy as np
    data = synth.generate(320, 240, 15, "IPB", 1, seed=4242, profile=0)
    info, frames = native_lib.parse_file(data)
    rng = np.random.default_rng(7)
    flagged = 0
    for trial in range(12):
        dec = native_lib.SeqDecoder(info.width, info.height, info.version)
        bufs = [(ctypes.c_uint8 * dec.frame_bytes)() for _ in range(3)]
        past, present, future = 0, 1, 2
        for fr in frames:
            if fr.frame_type != 0x30:
                past, future = future, past
            pic = bytearray(data[fr.offset:fr.offset + fr.bytes])
            if fr.frame_type != 0x10:                      # keep the I picture intact, damage P and B
                if trial % 2:
                    pic = pic[: max(80, len(pic) * (1 + trial % 5) // 6)]
                else:
                    for _ in range(60):
                        pic[int(rng.integers(76, len(pic)))] ^= 1 << int(rng.integers(0, 8))
            try:
                dec.decode(fr.frame_type, bytes(pic), bufs[present], bufs[past], bufs[future])
            except native_lib.HVQM4Error as e:
                assert e.bits < (1 << 16), f"runtime error, not a stream error: {e}"
                flagged += 1
            if fr.frame_type != 0x30:
                present, future = future, present
        dec.close()
    assert flagged > 0
    assert native_lib.lib().HVQM4GetLastCudaError() == 0
    # the library still decodes a valid stream bit-exactly afterwards
    a = [yuv for _, _, yuv in native_lib.Player(data)]
    b = [yuv for _, _, yuv in native_lib.Player(data)]
    assert a == b


def test_gpu_entropy_stage_matches_golden(native_lib, golden):
    """The bitstream stage compiled as device code (one warp per picture, entropy_dev.cu): the
    frames must equal the reference MD5s exactly like with the host stage."""
    names = ["cfg5_stream0", "cfg5_stream1", "cfg5_stream511", "cfg5_stream1023"]
    files = [synth.generate(**golden[n]["args"]) for n in names]
    step = 0
    for frames in native_lib.decode_streams(files, gpu_entropy=True):
        for n, (t, d, yuv) in zip(names, frames):
            assert md5(yuv) == golden[n]["md5"][step], (n, step)
        step += 1
    assert step == 16


@pytest.mark.parametrize("name", ["cfg4_320x240_v13_IPB", "realistic_640x480_v15_IPB", "ragged_328x248_v15_IPB", "wide_1024x576_v13_IPB",
                                  "hd_1280x720_v15_IPB", "tiny_16x16_v15_IPB", "small_64x48_v13_IPB", "mirror_h_200x152_v15_IPB",
                                  "uhd_4096x2160_v15_IPB", "stress_640x480_v15_IPB", "stress_320x240_v13_IPB", "stress_328x248_v15_IPB",
                                  "stress_64x48_v15_IPB", "portrait_240x320_v15_IPB", "portrait_480x640_v13_IPB",
                                  "portrait_stress_240x320_v15_IPB", "portrait_small_64x96_v15_IPB"])
def test_gpu_entropy_stage_other_geometries(native_lib, golden, name):
    case = golden[name]
    data = synth.generate(**case["args"])
    got = [md5(frames[0][2]) for frames in native_lib.decode_streams([data, data], gpu_entropy=True)]
    assert got == case["md5"]


@pytest.mark.parametrize("mode", [-1, 4, 6, 7])
def test_batch_runtime_matches_golden_stress(native_lib, golden, mode):
    """The generator's stress profile (long basis lists, wide shifts, escape chains, long runs) through the batch runtime
    under every reconstruction schedule that serves batches."""
    native_lib.set_recon_mode(mode)
    try:
        for name in ("stress_640x480_v15_IPB", "stress_320x240_v13_IPB", "portrait_240x320_v15_IPB"):
            case = golden[name]
            data = synth.generate(**case["args"])
            got = [[md5(f[2]) for f in frames] for frames in native_lib.decode_streams([data] * 5)]
            assert got == [[m] * 5 for m in case["md5"]], (name, mode)
    finally:
        native_lib.set_recon_mode(0)


def test_gpu_entropy_stage_symbol_capacity(native_lib, golden):
    """The GPU entropy stage holds 16 symbols per block of a section's plane (api.cpp).  I pictures whose every luma block
    carries 16 bases sit exactly at that capacity and decode; with 17 bases per block the section no longer fits: the
    picture is reported as truncated (error bit, no CUDA fault) -- the host stage decodes both.  The heavy stream shares
    its step with light ones: the device symbol arena of a step is sized for twice the frame bytes per stream on average."""
    light = synth.generate(320, 240, 15, "II", 1, seed=307, profile=1)
    for name, fits in (("cap16_320x240_v15_I", True), ("cap17_320x240_v15_I", False)):
        case = golden[name]
        data = synth.generate(**case["args"])
        assert [md5(frames[0][2]) for frames in native_lib.decode_streams([data, data])] == case["md5"]      # host stage
        files = [data] + [light] * 5
        parsed = [native_lib.parse_file(f) for f in files]
        bufs = [ctypes.create_string_buffer(f, len(f) + 8) for f in files]
        batch = native_lib.Batch(len(files), 320, 240, 15, gpu_entropy=True)
        try:
            errors = 0
            for k in range(2):
                frs = [p[1][k] for p in parsed]
                batch.decode(list(range(len(files))), [f.frame_type for f in frs],
                             [ctypes.addressof(bufs[i]) + frs[i].offset for i in range(len(files))], [f.bytes for f in frs])
                try:
                    batch.sync()
                except native_lib.HVQM4Error as e:
                    errors |= e.bits
                if fits:
                    assert md5(batch.read_frame(0)) == case["md5"][k]
            assert (errors == 0) == fits, (name, hex(errors))
        finally:
            batch.close()


def _decode_damaged(native_lib, files, gpu_entropy):
    """Decodes damaged streams through the batch runtime, tolerating stream-error bits."""
    parsed = [native_lib.parse_file(f) for f in files]
    info = parsed[0][0]
    n = len(files)
    bufs = [ctypes.create_string_buffer(f, len(f) + 8) for f in files]
    bases = [ctypes.addressof(b) for b in bufs]
    batch = native_lib.Batch(n, info.width, info.height, info.version, gpu_entropy=gpu_entropy)
    frames, bits = [], 0
    try:
        for k in range(len(parsed[0][1])):
            frs = [p[1][k] for p in parsed]
            for call in (lambda: batch.decode(list(range(n)), [f.frame_type for f in frs],
                                              [bases[i] + frs[i].offset for i in range(n)], [f.bytes for f in frs]),
                         batch.sync):
                try:
                    call()
                except native_lib.HVQM4Error as e:
                    assert e.bits < (1 << 16), f"runtime error, not a stream error: {e}"
                    bits |= e.bits
            frames.append([batch.read_frame(i) for i in range(n)])
    finally:
        batch.close()
    return frames, bits


@pytest.mark.parametrize("profile", [0, 1])
def test_gpu_entropy_stage_equals_host_stage_on_damaged_streams(native_lib, profile):
    """Bit-flipped pictures (I, P and B): whatever the host stage makes of them -- clamped types,
    exhausted sections, poisoned vectors -- the warp-parallel GPU build of the same code must
    make the same frames and raise the same error bits."""
    import numpy as np
    rng = np.random.default_rng(99 + profile)
    files = []
    for i in range(6):
        data = bytearray(synth.generate(320, 240, 15, "IPBBPB", 1, seed=900 + i, profile=profile))
        _, recs = native_lib.parse_file(bytes(data))
        for fr in recs:
            if i == 0:
                continue                                    # one intact stream
            for _ in range(1 + 4 * (i % 3)):
                at = fr.offset + int(rng.integers(76, fr.bytes))
                data[at] ^= 1 << int(rng.integers(0, 8))
        files.append(bytes(data))
    host, host_bits = _decode_damaged(native_lib, files, False)
    dev, dev_bits = _decode_damaged(native_lib, files, True)
    assert host == dev
    assert host_bits == dev_bits and host_bits != 0
    assert native_lib.lib().HVQM4GetLastCudaError() == 0


@pytest.mark.parametrize("recon", [0, 6])
def test_damaged_headers_and_section_tables_host_equals_gpu_stage(native_lib, recon):
    """The part of a picture the other fuzz test leaves alone: the 8-byte header (dc_shift, unk_shift, nest origin /
    residual-bit counts) and the section table, of I, P and B pictures.  Hostile shifts are clamped, a nest origin outside
    the map is pulled back, sections that point outside the record read as empty -- whatever the host stage makes of it,
    the GPU build makes the same frames and error bits, no kernel faults (also not the row kernel with its tensor copies,
    which leaves pictures with poisoned or out-of-plane vectors to the band kernel)."""
    import numpy as np
    rng = np.random.default_rng(4242)
    files = []
    for i in range(8):
        data = bytearray(synth.generate(320, 240, 15, "IPBBPB", 1, seed=950 + i, profile=i % 3))
        _, recs = native_lib.parse_file(bytes(data))
        for k, fr in enumerate(recs):
            if i == 0:
                continue                                    # one intact stream
            if (i + k) % 2:
                continue                                    # every other picture of a stream
            if i % 2:                                       # header bytes: any value
                data[fr.offset + int(rng.integers(0, 8))] = int(rng.integers(0, 256))
            else:                                           # section table: flip bits of an offset word
                at = fr.offset + 8 + int(rng.integers(0, 68))
                data[at] ^= 1 << int(rng.integers(0, 8))
        files.append(bytes(data))
    native_lib.set_recon_mode(recon)
    try:
        host, host_bits = _decode_damaged(native_lib, files, False)
        dev, dev_bits = _decode_damaged(native_lib, files, True)
    finally:
        native_lib.set_recon_mode(0)
    assert host == dev
    assert host_bits == dev_bits and host_bits != 0
    assert native_lib.lib().HVQM4GetLastCudaError() == 0
    assert native_lib.row_errors() == 0


def test_gpu_entropy_stage_matches_oracle_on_fresh_seeds(native_lib, oracle):
    files = [synth.generate(320, 240, 15 if seed % 2 else 13, "IPBBPBB", 1, seed=7100 + seed, profile=(seed // 2) % 2)
             for seed in range(2)]
    # same geometry and version per batch: two batches
    for data in files:
        want = [yuv for _, _, _, yuv in oracle.PortDecoder(data).frames()]
        got = [frames[0][2] for frames in native_lib.decode_streams([data, data, data], gpu_entropy=True)]
        assert got == want


def test_rgb_kernel_equals_reference_on_every_triple(native_lib, oracle):
    """HVQM4ConvertRGB (rgb.cu) against the reference's dumpRGB (h4m:895-926) on pictures that
    together hold all 2^24 (y, u, v) byte triples: bit-exact, so the float path has no room."""
    from tests.h4m_util import all_triples_pictures
    checker = oracle.RefDecoder if oracle.have_ref() else oracle.PortDecoder
    dec = None
    for yuv, w, h in all_triples_pictures():
        dec = dec or native_lib.SeqDecoder(w, h, 15)
        assert dec.to_rgb(yuv) == checker.yuv_to_rgb(yuv, w, h)
    dec.close()


def test_rgb_of_decoded_frames_sdk_and_batch(native_lib, oracle):
    """The reference converts every frame it decodes (h4m:2126): same bytes from the SDK-mode
    call on the host frame buffer and from the batched read-back, ragged geometry included."""
    checker = oracle.RefDecoder if oracle.have_ref() else oracle.PortDecoder
    for args in (dict(width=320, height=240, version=15, gop="IPBB", n_gops=1, seed=61, profile=0),
                 dict(width=328, height=248, version=13, gop="IPB", n_gops=1, seed=62, profile=1)):
        data = synth.generate(**args)
        w, h = args["width"], args["height"]
        want = [checker.yuv_to_rgb(yuv, w, h) for _, _, _, yuv in checker(data).frames()]
        player = native_lib.Player(data, rgb=True)
        got = []
        for _ in player:
            got.append(player.last_rgb)
        player.close()
        assert got == want
        # batched: three streams in lock step, RGB of all three read back per step
        info, frames = native_lib.parse_file(data)
        buf = ctypes.create_string_buffer(data, len(data) + 8)
        base = ctypes.addressof(buf)
        batch = native_lib.Batch(3, w, h, info.version, gpu_entropy=(args["seed"] == 61))
        try:
            for k, fr in enumerate(frames):
                batch.decode([0, 1, 2], [fr.frame_type] * 3, [base + fr.offset] * 3, [fr.bytes] * 3)
                rgb = batch.read_frames_rgb([2, 0, 1])
                assert rgb == [want[k]] * 3, k
        finally:
            batch.close()


@pytest.mark.parametrize("gpu_entropy,host_share", [(False, 0), (True, 0), (True, 20)])
def test_pipelined_steps_with_async_readback_match_oracle(native_lib, oracle, gpu_entropy, host_share):
    """The way bench.py's end-to-end arm drives the batch runtime: every step of a GOP submitted
    back to back, each followed by an asynchronous read-back of all frames into its own pinned
    buffer, one sync at the end.  Uploads, entropy stage, reconstruction and read-backs of
    neighbouring steps overlap (staging ring, alternating parser slots, spare B surface), so every
    frame of every step is compared with the oracle: a missing dependency shows up as a torn frame.
    host_share: the first streams of the batch are parsed by the host threads next to the parse kernel
    (HVQM4BatchSetHostShare), both stages feed the same reconstruction launch."""
    S, distinct, gop = 96, 3, "IPBBPBBPBB"
    files = [synth.generate(640, 480, 15, gop, 1, seed=8300 + i, profile=i & 1) for i in range(distinct)]
    want = [[md5(yuv) for _, _, _, yuv in oracle.PortDecoder(f).frames()] for f in files]
    parsed = [native_lib.parse_file(f) for f in files]
    bufs = [ctypes.create_string_buffer(f, len(f) + 8) for f in files]
    bases = [ctypes.addressof(b) for b in bufs]
    n_steps = len(parsed[0][1])
    batch = native_lib.Batch(S, 640, 480, 15, gpu_entropy=gpu_entropy, host_share=host_share)
    fb = batch.frame_bytes
    pinned = native_lib.lib().HVQM4HostAlloc(n_steps * S * fb)
    assert pinned
    try:
        ids = list(range(S))
        ids_arr = (ctypes.c_int32 * S)(*ids)
        for rep in range(2):            # second GOP: the rings have wrapped and every surface has been in every role
            for k in range(n_steps):
                frs = [parsed[i % distinct][1][k] for i in range(S)]
                step = native_lib.Batch.prepare_step(ids, [f.frame_type for f in frs], [bases[i % distinct] + frs[i].offset for i in range(S)],
                                                     [f.bytes for f in frs])
                batch.decode_prepared(step)
                batch.read_frames_async(ids_arr, S, pinned + k * S * fb, fb)
            batch.sync()
            raw = ctypes.string_at(pinned, n_steps * S * fb)
            for k in range(n_steps):
                for i in range(S):
                    got = md5(raw[(k * S + i) * fb:(k * S + i + 1) * fb])
                    assert got == want[i % distinct][k], (rep, k, i)
    finally:
        batch.close()
        native_lib.lib().HVQM4HostFree(pinned)


def test_audio_kernel_equals_oracle(native_lib, oracle):
    """HVQM4DecodeAudioBatch (audio.cu) against the reference's decode_audio (h4m:185-258): 48
    independent streams per call, mono and stereo, a seeded first record and two continuation
    records each (the predictor state travels through HVQM4AudioState); a short record raises
    HVQM4_ERR_TRUNCATED and a seed index > 88 HVQM4_ERR_ARGUMENT."""
    import random
    import struct
    checker = oracle.RefDecoder if oracle.have_ref() else oracle.PortDecoder
    rng = random.Random(21)
    for ch in (1, 2):
        n = 48
        states = [native_lib.AudioState() for _ in range(n)]
        ref_states = [[0] * (2 * ch) for _ in range(n)]
        for k in range(3):
            payloads, want = [], []
            for i in range(n):
                samples = rng.randint(1, 700)
                body = bytearray(rng.randrange(256) for _ in range(2 * ch + samples * ch))
                if k == 0:
                    for c in range(ch):
                        body[2 * c + 1] = (body[2 * c + 1] & 0x80) | rng.randint(0, 88)
                payloads.append(struct.pack(">I", samples) + bytes(body))
                want.append(checker.decode_audio(ref_states[i], ch, k == 0, samples, bytes(body)))
            rc, got = native_lib.decode_audio_batch(ch, states, [k == 0] * n, payloads)
            assert rc == 0
            assert got == want
            for i in range(n):
                assert [states[i].hist[c] for c in range(ch)] == [ref_states[i][2 * c] for c in range(ch)]
                assert [states[i].idx[c] for c in range(ch)] == [ref_states[i][2 * c + 1] for c in range(ch)]
    st = [native_lib.AudioState()]
    rc, got = native_lib.decode_audio_batch(1, st, [True], [struct.pack(">I", 100) + bytes([1, 5]) + bytes(10)])
    assert rc == native_lib.ERR_TRUNCATED and len(got[0]) == 21          # seed sample + 2 per byte
    rc, _ = native_lib.decode_audio_batch(1, st, [True], [struct.pack(">I", 4) + bytes([1, 0x7F]) + bytes(10)])
    assert rc & native_lib.ERR_ARGUMENT


def test_file_player_matches_the_reference_program(native_lib, oracle):
    """HVQM4Player* = the reference program as a library: every video record of a two-GOP file with
    an interleaved stereo audio track, in file order, with the display index of the reference's
    output file name (h4m:2122), its RGB conversion (h4m:2126) and the audio track (h4m:185-258)."""
    from tests.h4m_util import with_audio
    checker = oracle.RefDecoder if oracle.have_ref() else oracle.PortDecoder
    for version, gop in ((15, "IPBBPBB"), (13, "IPB")):
        data = synth.generate(320, 240, version, gop, 2, seed=90 + version, profile=0)
        spliced, records = with_audio(data, 2, 2, 300, seed=version)
        want = list(checker(spliced).frames())
        player = native_lib.FilePlayer(spliced)
        assert (player.info.n_audio_frames, player.info.audio_channels, player.info.audio_sample_rate) == (len(records), 2, 22050)
        n = 0
        for (t, disp, yuv), (wt, wd, wdisp, wyuv) in zip(player.frames(), want):
            assert (t, disp) == (wt, wdisp) and wdisp == (n // len(gop)) * len(gop) + wd, n
            assert yuv == wyuv, n
            if n % 3 == 0:
                assert player.rgb() == checker.yuv_to_rgb(wyuv, 320, 240), n
            n += 1
        assert n == len(want) == 2 * len(gop)
        state = [0, 0, 0, 0]
        pcm = list(player.audio())
        assert len(pcm) == len(records)
        for got, (g, first, payload) in zip(pcm, records):
            if first:
                state = [0, 0, 0, 0]
            samples = int.from_bytes(payload[:4], "big")
            assert got == checker.decode_audio(state, 2, first, samples, payload[4:])
        assert player.errors() == 0
        player.close()


@pytest.mark.parametrize("gpu_entropy,host_share", [(False, 0), (True, 0), (True, 7), (True, 24)])
def test_staggered_streams_and_partial_steps_match_oracle(native_lib, oracle, gpu_entropy, host_share):
    """Streams that are out of phase with each other (stream i joins at step i % 4 and sits out every
    step where (step + i) % 5 == 0): every step mixes I, P and B pictures and addresses a different
    subset of the batch, read-backs are asynchronous, nothing is synchronised until the end.  The
    per-stream surface rotation, the alternating parser slots (a stream may use the same slot twice
    in a row), the I-picture fences and the read-back dependencies all have to hold."""
    n, gop = 24, "IPBBPBPB"
    files = [synth.generate(320, 240, 15, gop, 2, seed=8600 + i, profile=i % 2) for i in range(n)]
    want = [[md5(yuv) for _, _, _, yuv in oracle.PortDecoder(f).frames()] for f in files]
    parsed = [native_lib.parse_file(f)[1] for f in files]
    bufs = [ctypes.create_string_buffer(f, len(f) + 8) for f in files]
    bases = [ctypes.addressof(b) for b in bufs]
    batch = native_lib.Batch(n, 320, 240, 15, gpu_entropy=gpu_entropy, host_share=host_share)
    fb = batch.frame_bytes
    total_frames = sum(len(p) for p in parsed)
    pinned = native_lib.lib().HVQM4HostAlloc(total_frames * fb)
    assert pinned
    try:
        cursor = [0] * n
        slots, slot = [], 0          # (stream, frame index) of every read-back, in order
        step = 0
        while any(cursor[i] < len(parsed[i]) for i in range(n)):
            ids = [i for i in range(n) if step >= i % 4 and (step + i) % 5 != 0 and cursor[i] < len(parsed[i])]
            if ids:
                frs = [parsed[i][cursor[i]] for i in ids]
                batch.decode(ids, [f.frame_type for f in frs], [bases[i] + f.offset for i, f in zip(ids, frs)], [f.bytes for f in frs])
                arr = (ctypes.c_int32 * len(ids))(*ids)
                batch.read_frames_async(arr, len(ids), pinned + slot * fb, fb)
                for i in ids:
                    slots.append((i, cursor[i]))
                    cursor[i] += 1
                slot += len(ids)
            step += 1
        batch.sync()
        raw = ctypes.string_at(pinned, total_frames * fb)
        assert len(slots) == total_frames
        for k, (i, f) in enumerate(slots):
            assert md5(raw[k * fb:(k + 1) * fb]) == want[i][f], (i, f)
    finally:
        batch.close()
        native_lib.lib().HVQM4HostFree(pinned)


def test_sdk_mode_notices_a_reference_frame_the_application_wrote_into(native_lib, oracle):
    """Drop-in mode keeps device twins of the host frame buffers.  An application that writes into a reference frame
    between two decode calls must get pictures predicted from what it wrote -- the same as after an explicit
    HVQM4InvalidateFrame -- and a frame it scribbled over and restored must still predict the reference's picture."""
    data = synth.generate(320, 240, 15, "IPPP", 1, seed=8900, profile=0)
    info, frames = native_lib.parse_file(data)
    fb = info.width * info.height * 3 // 2
    pics = [data[f.offset:f.offset + f.bytes] for f in frames]
    want = [yuv for _, _, _, yuv in oracle.PortDecoder(data).frames()]
    lib = native_lib.lib()

    def run(invalidate):
        dec = native_lib.SeqDecoder(info.width, info.height, info.version)
        try:
            a, b = (ctypes.c_uint8 * (fb + 64))(), (ctypes.c_uint8 * (fb + 64))()
            dec.decode(frames[0].frame_type, pics[0], a)
            assert bytes(a)[:fb] == want[0]
            dec.decode(frames[1].frame_type, pics[1], b, a)
            assert bytes(b)[:fb] == want[1]
            ctypes.memset(b, 0x33, fb)                 # scribbled over ...
            ctypes.memmove(b, want[1], fb)             # ... and restored
            if invalidate:
                lib.HVQM4InvalidateFrame(ctypes.byref(dec.seq), b)
            dec.decode(frames[2].frame_type, pics[2], a, b)
            assert bytes(a)[:fb] == want[2]
            ctypes.memset(a, 0x80, fb)                 # really changed: the next picture is predicted from grey
            if invalidate:
                lib.HVQM4InvalidateFrame(ctypes.byref(dec.seq), a)
            dec.decode(frames[3].frame_type, pics[3], b, a)
            return bytes(b)[:fb]
        finally:
            dec.close()

    from_grey = run(True)
    assert from_grey != want[3]
    assert run(False) == from_grey                     # noticed without being told


def test_entropy_mode_switch_after_the_first_picture_is_refused(native_lib, golden):
    """The host and the GPU entropy stage keep separate per-stream state: switching once pictures have been decoded
    would reconstruct the next P/B pictures against the wrong nest.  Same mode again is fine."""
    case = golden["cfg4_320x240_v13_IPB"]
    data = synth.generate(**case["args"])
    info, frames = native_lib.parse_file(data)
    buf = ctypes.create_string_buffer(data, len(data) + 8)
    lib = native_lib.lib()
    batch = native_lib.Batch(1, info.width, info.height, info.version)
    try:
        fr = frames[0]
        batch.decode([0], [fr.frame_type], [ctypes.addressof(buf) + fr.offset], [fr.bytes])
        batch.sync()
        assert lib.HVQM4BatchSetEntropyMode(batch._h, 1) != 0
        assert lib.HVQM4BatchSetEntropyMode(batch._h, 0) == 0
        assert lib.HVQM4BatchSetHostShare(batch._h, 1) != 0       # the share is fixed with the first picture as well
        assert lib.HVQM4BatchSetHostShare(batch._h, 0) == 0
        assert lib.HVQM4BatchSetHostShare(batch._h, 2) != 0       # more streams than the batch has
        for k, fr in enumerate(frames[1:], 1):
            batch.decode([0], [fr.frame_type], [ctypes.addressof(buf) + fr.offset], [fr.bytes])
            batch.sync()
            assert md5(batch.read_frame(0)) == case["md5"][k]
    finally:
        batch.close()


def test_unregister_waits_for_pictures_still_being_fetched(native_lib, oracle):
    """In gather mode the GPU reads the registered picture bytes after HVQM4BatchDecode has returned; unregistering right
    behind the call must wait for those reads (and the application may overwrite the bytes once sync() has returned)."""
    n = 24
    data = synth.generate(320, 240, 15, "I", 1, seed=8800, profile=0)
    want = [md5(yuv) for _, _, _, yuv in oracle.PortDecoder(data).frames()][0]
    fr = native_lib.parse_file(data)[1][0]
    image = ctypes.create_string_buffer(data, len(data) + 64)
    base = ctypes.addressof(image)
    lib = native_lib.lib()
    assert lib.HVQM4HostRegister(base, len(data)) == 0
    batch = native_lib.Batch(n, 320, 240, 15, gpu_entropy=True)
    try:
        batch.decode(list(range(n)), [fr.frame_type] * n, [base + fr.offset] * n, [fr.bytes] * n)
        assert lib.HVQM4HostUnregister(base) == 0          # no sync in between
        ctypes.memset(base, 0xEE, len(data))               # the application reuses its buffer
        batch.sync()
        assert all(md5(batch.read_frame(i)) == want for i in range(n))
    finally:
        batch.close()


@pytest.mark.parametrize("host_share", [0, 5])
def test_registered_host_memory_is_gathered_by_the_gpu(native_lib, oracle, host_share):
    """HVQM4HostRegister: with the bitstreams in page-locked, mapped application memory the GPU entropy mode
    fetches the pictures itself (dev_gather_kernel) -- same frames as the oracle, every source alignment
    modulo 16 occurs (the pictures sit at arbitrary offsets inside the file images), and a step that mixes
    registered and unregistered pictures falls back to the host copy.  The files are registered one by one:
    neighbours share pages, which are registered once and released with their last user."""
    n, gop = 20, "IPBBPB"
    files = [synth.generate(320, 240, 15, gop, 1, seed=8700 + i, profile=i % 2) for i in range(n)]
    want = [[md5(yuv) for _, _, _, yuv in oracle.PortDecoder(f).frames()] for f in files]
    parsed = [native_lib.parse_file(f)[1] for f in files]
    # one shared image with the files at odd offsets, registered file by file; plus an unregistered copy of stream 0
    offs, blob = [], bytearray()
    for i, f in enumerate(files):
        blob += bytes(i % 7 + 1)
        offs.append(len(blob))
        blob += f
    blob += bytes(64)
    image = ctypes.create_string_buffer(bytes(blob), len(blob))
    loose = ctypes.create_string_buffer(files[0], len(files[0]) + 8)
    base = ctypes.addressof(image)
    lib = native_lib.lib()
    for i, f in enumerate(files):
        assert lib.HVQM4HostRegister(base + offs[i], len(f)) == 0
    batch = native_lib.Batch(n, 320, 240, 15, gpu_entropy=True, host_share=host_share)
    try:
        assert {(base + offs[i] + parsed[i][k].offset) & 15 for i in range(n) for k in range(len(gop))} == set(range(16))
        launches0 = native_lib.kernel_launches()
        for k in range(len(gop)):
            ptrs = [base + offs[i] + parsed[i][k].offset for i in range(n)]
            if k == 3:
                ptrs[0] = ctypes.addressof(loose) + parsed[0][k].offset      # mixed step: host copy path
            batch.decode(list(range(n)), [parsed[i][k].frame_type for i in range(n)], ptrs, [parsed[i][k].bytes for i in range(n)])
            batch.sync()
            for i in range(n):
                assert md5(batch.read_frame(i)) == want[i][k], (i, k)
        # gather + parse + band kernel per gathered step, parse + band for the mixed one
        assert native_lib.kernel_launches() - launches0 == 3 * (len(gop) - 1) + 2
    finally:
        batch.close()
        for i in range(n):
            assert lib.HVQM4HostUnregister(base + offs[i]) == 0
        assert lib.HVQM4HostUnregister(base + offs[0]) != 0
        # every page was released with its last user: the whole image can be registered again
        assert lib.HVQM4HostRegister(base, len(blob)) == 0
        assert lib.HVQM4HostUnregister(base) == 0


def test_read_back_into_memory_that_shares_a_page_with_a_registered_range(native_lib, oracle):
    """HVQM4HostRegister pins whole pages, so a heap block next to a registered range shares the range's last page, and
    cudaMemcpyAsync refuses a destination that straddles page-locked and pageable memory (seen as a 1-in-17 failure of
    the test above, whenever Python happened to put a read-back buffer there).  The library then takes the frame over
    its own page-locked bounce buffer: single-frame and pitched multi-frame read-backs into exactly such a destination."""
    n = 6
    data = synth.generate(320, 240, 15, "I", 1, seed=8801, profile=0)
    want = [md5(yuv) for _, _, _, yuv in oracle.PortDecoder(data).frames()][0]
    fr = native_lib.parse_file(data)[1][0]
    fb = 320 * 240 * 3 // 2
    arena = ctypes.create_string_buffer(len(data) + (n + 2) * fb + 4 * 4096)
    base = (ctypes.addressof(arena) + 4095) & ~4095
    length = ((len(data) + 4095) & ~4095) + 100              # the range ends 100 bytes into its last page
    ctypes.memmove(base, data, len(data))
    lib = native_lib.lib()
    assert lib.HVQM4HostRegister(base, length) == 0
    dst = base + ((length + 4095) & ~4095) - 2000            # begins in that page, ends n frames further on in pageable memory
    assert dst >= base + length
    batch = native_lib.Batch(n, 320, 240, 15, gpu_entropy=True)
    try:
        batch.decode(list(range(n)), [fr.frame_type] * n, [base + fr.offset] * n, [fr.bytes] * n)
        batch.sync()
        assert lib.HVQM4BatchReadFrame(batch._h, 0, ctypes.c_void_p(dst)) == 0
        assert md5(ctypes.string_at(dst, fb)) == want
        ctypes.memset(dst, 0, n * fb)
        ids = (ctypes.c_int32 * n)(*range(n))
        batch.read_frames_async(ids, n, dst, fb)
        batch.sync()
        assert all(md5(ctypes.string_at(dst + i * fb, fb)) == want for i in range(n))
    finally:
        batch.close()
        assert lib.HVQM4HostUnregister(base) == 0

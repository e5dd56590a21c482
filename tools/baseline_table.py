"""Measures the rows of BASELINE.md section 5 (configs 1-4; config 5 is bench.py's line) on one B200:
    python tools/baseline_table.py > gpurun_out/baseline_table.json
per config: MD5 parity against the committed goldens (the reference's frames), single-stream SDK-mode frames/s (host
buffers, synchronous: the latency path), reconstruction-only frames/s of 1024 copies of the stream batched per launch
(symbol buffers resident in HBM, CUDA events), the same end to end (host bitstreams in, frames back in pinned host
memory, GPU entropy stage), and the reference decoder on one host core."""
import ctypes
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hvqm4_b200 import api, synth  # noqa: E402
from oracle import bindings  # noqa: E402

golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_md5.json")))["cases"]
CONFIGS = [("1. 320x240 1.5 I-only x30", "cfg1_320x240_v15_I30"), ("2. 640x480 1.5 I/P GOP-15", "cfg2_640x480_v15_IP15"),
           ("3. 640x480 1.5 I/P/B", "cfg3_640x480_v15_IPB"), ("4. 320x240 1.3 I/P/B", "cfg4_320x240_v13_IPB")]
S = 1024
PEAK = 6553.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, KeyError, ValueError):
    pass
bindings.build(ref=os.path.exists("/root/reference/h4m_audio_decode.c"), port=True)
dec = bindings.RefDecoder if bindings.have_ref() else bindings.PortDecoder
rows = []
for label, name in CONFIGS:
    case = golden[name]
    data = synth.generate(**case["args"])
    w, h = case["args"]["width"], case["args"]["height"]
    # parity + SDK-mode single stream
    for _ in range(2):
        pl = api.Player(data)
        t0 = time.perf_counter()
        got = [hashlib.md5(yuv).hexdigest() for _, _, yuv in pl]
        t_sdk = time.perf_counter() - t0
        pl.close()
    parity = got == case["md5"]
    # batched: S copies of the stream, one picture per stream per step
    info, frames = api.parse_file(data)
    buf = ctypes.create_string_buffer(data, len(data) + 8)
    base = ctypes.addressof(buf)
    ids = list(range(S))
    batch = api.Batch(S, w, h, info.version)
    batch.record(True)
    for fr in frames:
        batch.decode(ids, [fr.frame_type] * S, [base + fr.offset] * S, [fr.bytes] * S)
    batch.sync()
    batch.record(False)
    st = batch.stats()
    batch.replay(2)
    reps = 5
    ms = batch.replay(reps)
    fps = S * len(frames) * reps / (ms * 1e-3)
    gbs = st["algorithmic_bytes"] * reps / (ms * 1e-3) / 1e9
    batch_parity = hashlib.md5(batch.read_frame(S - 1)).hexdigest() == case["md5"][-1]
    batch.close()
    # end to end
    fb = w * h * 3 // 2
    pinned = api.lib().HVQM4HostAlloc(S * fb)
    gb = api.Batch(S, w, h, info.version, gpu_entropy=True)
    ids_arr = (ctypes.c_int32 * S)(*ids)
    steps = [api.Batch.prepare_step(ids, [fr.frame_type] * S, [base + fr.offset] * S, [fr.bytes] * S) for fr in frames]

    def gop():
        for stp in steps:
            gb.decode_prepared(stp)
            gb.read_frames_async(ids_arr, S, pinned, fb)
    gop()
    gb.sync()
    k = 3
    t0 = time.perf_counter()
    for _ in range(k):
        gop()
    gb.sync()
    e2e = S * len(frames) * k / (time.perf_counter() - t0)
    e2e_parity = hashlib.md5(ctypes.string_at(pinned + (S - 1) * fb, fb)).hexdigest() == case["md5"][-1]
    gb.close()
    api.lib().HVQM4HostFree(pinned)
    t_ref, n_ref = dec.bench(data, 2)
    rows.append({"config": label, "frames": len(frames), "md5_parity": bool(parity and batch_parity and e2e_parity),
                 "sdk_single_stream_fps": len(frames) / t_sdk, "recon_fps_1024_streams": fps, "recon_mpix_s": fps * w * h / 1e6,
                 "recon_gbs": gbs, "frac_of_8TBs": gbs / 8000.0, "frac_of_measured": gbs / PEAK, "e2e_fps_1024_streams": e2e,
                 "ref_cpu_fps_1core": n_ref / t_ref, "checker": dec.__name__})
    print(label, rows[-1], file=sys.stderr)
print(json.dumps({"peak_gbs": PEAK, "streams": S, "rows": rows}, indent=1))

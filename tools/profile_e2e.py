"""End-to-end timing experiments: entropy stage (host or GPU) + H2D + kernels (+ optional D2H).
    python tools/profile_e2e.py [S] [threads] [d2h 0|1] [profile] [gops] [gpu_entropy 0|1] [host_share streams]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hvqm4_b200 import api, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 0
D2H = int(sys.argv[3]) if len(sys.argv) > 3 else 1
PROFILE = int(sys.argv[4]) if len(sys.argv) > 4 else 0
GOPS = int(sys.argv[5]) if len(sys.argv) > 5 else 3
GPU_ENTROPY = int(sys.argv[6]) if len(sys.argv) > 6 else 0
SHARE = int(sys.argv[7]) if len(sys.argv) > 7 else 0
GOP = "I" + "PBB" * 5
distinct = min(S, 64)
files = [synth.generate(640, 480, 15, GOP, 1, seed=5000 + i, profile=PROFILE) for i in range(distinct)]
parsed = [api.parse_file(f) for f in files]
bufs = [ctypes.create_string_buffer(f, len(f) + 8) for f in files]
bases = [ctypes.addressof(b) for b in bufs]
if os.environ.get('E2E_REGISTER'):
    for b in bufs:
        assert api.lib().HVQM4HostRegister(ctypes.addressof(b), len(b)) == 0
batch = api.Batch(S, 640, 480, 15, host_threads=T, gpu_entropy=bool(GPU_ENTROPY), host_share=SHARE)
ids = list(range(S))
steps = []
for k in range(len(parsed[0][1])):
    frs = [parsed[i % distinct][1][k] for i in range(S)]
    steps.append(api.Batch.prepare_step(ids, [f.frame_type for f in frs], [bases[i % distinct] + frs[i].offset for i in range(S)], [f.bytes for f in frs]))
ids_arr = (ctypes.c_int32 * S)(*ids)
pinned = api.lib().HVQM4HostAlloc(S * batch.frame_bytes)


def gop():
    for st in steps:
        batch.decode_prepared(st)
        if D2H:
            batch.read_frames_async(ids_arr, S, pinned, batch.frame_bytes)


gop()
batch.sync()
h0 = batch.stats()["host_ns"]
t0 = time.perf_counter()
for _ in range(GOPS):
    gop()
batch.sync()
t1 = time.perf_counter()
h1 = batch.stats()["host_ns"]
n = S * 16 * GOPS
print(f"gpu_entropy={GPU_ENTROPY} share={SHARE} S={S} threads={T or os.cpu_count()} d2h={D2H} profile={PROFILE}: {n / (t1 - t0):.0f} fps e2e; "
      f"host stage alone {(h1 - h0) / 1e9:.3f} s of {t1 - t0:.3f} s wall -> {n / ((h1 - h0) / 1e9):.0f} fps if host-only")
if GPU_ENTROPY:
    prof = (ctypes.c_uint64 * 8)()
    api.lib().HVQM4DevEntropyProfile(prof)
    names = ["hdr+trees", "pass1", "plan", "schedule", "copies", "flat decode", "fill records", "vector chain"]
    tot = sum(prof[:8]) or 1
    npic = S * 16 * (GOPS + 1)
    print("   GPU parser, per picture: " + ", ".join(f"{nm} {prof[i] / npic / 1.9e3:.0f} us ({100 * prof[i] / tot:.0f}%)" for i, nm in enumerate(names)))
batch.close()

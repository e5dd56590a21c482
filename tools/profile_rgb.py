"""End to end with the reference's RGB output: GPU entropy stage + reconstruction + yuv2rgb kernel
+ read-back of every RGB frame (what the reference program writes as PPM, h4m:2126).
    python tools/profile_rgb.py [S] [profile] [gops]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hvqm4_b200 import api, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
PROFILE = int(sys.argv[2]) if len(sys.argv) > 2 else 0
GOPS = int(sys.argv[3]) if len(sys.argv) > 3 else 2
GOP = "I" + "PBB" * 5
distinct = min(S, 64)
files = [synth.generate(640, 480, 15, GOP, 1, seed=5000 + i, profile=PROFILE) for i in range(distinct)]
parsed = [api.parse_file(f) for f in files]
bufs = [ctypes.create_string_buffer(f, len(f) + 8) for f in files]
bases = [ctypes.addressof(b) for b in bufs]
batch = api.Batch(S, 640, 480, 15, gpu_entropy=True)
ids = list(range(S))
steps = []
for k in range(len(parsed[0][1])):
    frs = [parsed[i % distinct][1][k] for i in range(S)]
    steps.append(api.Batch.prepare_step(ids, [f.frame_type for f in frs], [bases[i % distinct] + frs[i].offset for i in range(S)], [f.bytes for f in frs]))
ids_arr = (ctypes.c_int32 * S)(*ids)
rgb_bytes = 640 * 480 * 3
pinned = api.lib().HVQM4HostAlloc(S * rgb_bytes)


def gop():
    for st in steps:
        batch.decode_prepared(st)
        batch.read_frames_rgb_async(ids_arr, S, pinned, rgb_bytes)


gop()
batch.sync()
t0 = time.perf_counter()
for _ in range(GOPS):
    gop()
batch.sync()
t1 = time.perf_counter()
n = S * 16 * GOPS
print(f"S={S} profile={PROFILE}: {n / (t1 - t0):.0f} fps end to end with RGB read-back "
      f"({n * rgb_bytes / (t1 - t0) / 1e9:.1f} GB/s D2H)")
batch.close()

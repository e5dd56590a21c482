"""Small fixed workload for ncu: records one GOP of S streams, then replays it R times.
Launch order: 16 launches while recording (with host stage + uploads), then 16*R replays.
    python tools/profile_recon.py [S] [R] [profile]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hvqm4_b200 import api, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
R = int(sys.argv[2]) if len(sys.argv) > 2 else 2
PROFILE = int(sys.argv[3]) if len(sys.argv) > 3 else 0
GOP = "I" + "PBB" * 5
distinct = min(S, 32)
files = [synth.generate(640, 480, 15, GOP, 1, seed=5000 + i, profile=PROFILE) for i in range(distinct)]
parsed = [api.parse_file(f) for f in files]
bufs = [ctypes.create_string_buffer(f, len(f) + 8) for f in files]
bases = [ctypes.addressof(b) for b in bufs]
batch = api.Batch(S, 640, 480, 15)
batch.record(True)
for k in range(len(parsed[0][1])):
    frs = [parsed[i % distinct][1][k] for i in range(S)]
    batch.decode(list(range(S)), [f.frame_type for f in frs], [bases[i % distinct] + frs[i].offset for i in range(S)], [f.bytes for f in frs])
batch.sync()
batch.record(False)
ms = batch.replay(R)
st = batch.stats()
print(f"S={S} R={R} profile={PROFILE} replay {ms:.3f} ms -> {S * 16 * R / ms * 1e3:.0f} fps, "
      f"{st['algorithmic_bytes'] / 16 / (ms / (16 * R)) / 1e6:.1f} GB/s algorithmic")
batch.close()

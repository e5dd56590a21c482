"""Single-stream latency path (BASELINE config 2) driven like a C caller: HVQM4PlayerNextFrame in a loop,
no Python copies of the frames.  HVQM4_SDK_TRACE=1 adds the per-phase split of the SDK calls.
    python tools/profile_sdk.py [profile] [gop]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hvqm4_b200 import api, synth  # noqa: E402

PROFILE = int(sys.argv[1]) if len(sys.argv) > 1 else 0
GOP = sys.argv[2] if len(sys.argv) > 2 else "I" + "P" * 14
data = synth.generate(640, 480, 15, GOP, 4, seed=102, profile=PROFILE)
lib = api.lib()
for rep in range(3):
    pl = api.FilePlayer(data)
    ptr, disp, ftype = ctypes.c_void_p(), ctypes.c_uint32(), ctypes.c_uint32()
    lib.HVQM4PlayerNextFrame(pl._h, ctypes.byref(ptr), ctypes.byref(disp), ctypes.byref(ftype))
    t0 = time.perf_counter()
    n = 0
    while lib.HVQM4PlayerNextFrame(pl._h, ctypes.byref(ptr), ctypes.byref(disp), ctypes.byref(ftype)) == 1:
        n += 1
    dt = time.perf_counter() - t0
    print(f"profile={PROFILE} gop={GOP}: {n} frames, {n / dt:.0f} frames/s, {1e6 * dt / n:.0f} us per frame (C player loop)")
    pl.close()

"""examples/h4m_player.c -- the reference program's main loop written in C against the library
(both variants: HVQM4Player* and the seven SDK entry points driven like decode_video()).
CPU: it compiles with gcc against include/hvqm4.h and links with libhvqm4_b200.so.
GPU: both binaries decode a generated file; their per-frame output must equal the oracle's."""
import os
import subprocess

import pytest

from hvqm4_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fnv1a(b: bytes) -> str:
    import numpy as np
    h = 0xcbf29ce484222325
    for x in np.frombuffer(b, np.uint8).tolist():
        h = ((h ^ x) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def build_example(native_lib, tmpdir, sdk_calls: bool) -> str:
    out = os.path.join(str(tmpdir), "h4m_player_sdk" if sdk_calls else "h4m_player")
    libdir = os.path.dirname(native_lib.LIB)
    cmd = ["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-std=c11", "-I", os.path.join(ROOT, "include")]
    if sdk_calls:
        cmd.append("-DUSE_SDK_CALLS")
    cmd += [os.path.join(ROOT, "examples", "h4m_player.c"), "-L", libdir, "-lhvqm4_b200", f"-Wl,-rpath,{libdir}", "-o", out]
    subprocess.check_call(cmd)
    return out


@pytest.mark.parametrize("sdk_calls", [False, True])
def test_c_example_compiles_and_links(native_lib, tmp_path, sdk_calls):
    exe = build_example(native_lib, tmp_path, sdk_calls)
    assert os.path.exists(exe)
    # no file argument: usage message, exit code 2, and nothing has touched CUDA yet
    assert subprocess.run([exe], capture_output=True).returncode == 2


@pytest.mark.gpu
@pytest.mark.parametrize("sdk_calls", [False, True])
def test_c_example_output_equals_oracle(native_lib, oracle, tmp_path, sdk_calls):
    exe = build_example(native_lib, tmp_path, sdk_calls)
    checker = oracle.RefDecoder if oracle.have_ref() else oracle.PortDecoder
    for version, gop in ((15, "IPBBPB"), (13, "IPB")):
        data = synth.generate(320, 240, version, gop, 2, seed=40 + version, profile=0)
        path = os.path.join(str(tmp_path), f"v{version}.h4m")
        with open(path, "wb") as f:
            f.write(data)
        want = [f"{'?IPB'[t >> 4]} {disp} {fnv1a(yuv)} {fnv1a(checker.yuv_to_rgb(yuv, 320, 240))}"
                for t, _, disp, yuv in checker(data).frames()]
        run = subprocess.run([exe, path], capture_output=True, text=True)
        assert run.returncode == 0, run.stderr
        assert run.stdout.strip().split("\n") == want

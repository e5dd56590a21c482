"""The drop-in boundary: libhvqm4_b200.so loads without a GPU and exports every symbol that
include/hvqm4.h declares; the SDK structs have the reference's layout.  No compute calls."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "hvqm4.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(HVQM4\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native_lib):
    lib = ctypes.CDLL(native_lib.LIB)
    names = declared_functions()
    assert {"HVQM4InitDecoder", "HVQM4InitSeqObj", "HVQM4BuffSize", "HVQM4SetBuffer",
            "HVQM4DecodeIpic", "HVQM4DecodePpic", "HVQM4DecodeBpic"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hvqm4.h but not exported"
    assert set(native_lib.SIGNATURES) == set(names)


def test_sdk_struct_layout_matches_reference(native_lib):
    # SeqObj{VideoState*; u16 width,height; u8 h_samp,v_samp} h4m:516-523 ; VideoInfo h4m:533-540
    assert native_lib.SeqObj.width.offset == ctypes.sizeof(ctypes.c_void_p)
    assert native_lib.SeqObj.height.offset == native_lib.SeqObj.width.offset + 2
    assert native_lib.SeqObj.h_samp.offset == native_lib.SeqObj.width.offset + 4
    assert native_lib.VideoInfo.video_mode.offset == 6


def test_container_walker_matches_python_demux(native_lib):
    from hvqm4_b200 import synth
    from tests.h4m_util import demux
    data = synth.generate(320, 240, 13, "IPBBP", 2, seed=3, profile=1)
    info, frames = native_lib.parse_file(data)
    version, w, h, recs = demux(data)
    assert (info.version, info.width, info.height, info.n_gops, info.n_video_frames) == (version, w, h, 2, 10)
    assert [(f.frame_type, f.disp_id, f.bytes) for f in frames] == [(t, d, len(p)) for t, d, p in recs]
    assert all(data[f.offset:f.offset + f.bytes] == recs[i][2] for i, f in enumerate(frames))


def test_no_device_means_loud_failure_not_fallback(native_lib):
    """On a machine without a GPU the batch runtime refuses to start; nothing decodes on the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(native_lib.HVQM4Error):
        native_lib.Batch(1, 320, 240)

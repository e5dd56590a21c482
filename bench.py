#!/usr/bin/env python
"""bench.py -- HVQM4 picture-decode throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json config 5 -- 1024 independent synthetic 640x480 HVQM4
1.5 I/P/B streams (GOP = I + 5 x PBB, dense profile; 128 distinct bitstreams per GPU, reused
cyclically with independent decoder state), one picture per stream per step, batched per launch.
Multi-GPU: every rank owns its own 1024 streams (weak scaling, --streams-per-gpu fixed as N
grows), no collective -- streams are independent.  One STEP = one GOP (16 pictures) of every
stream of the rank = 16 launches of the fused band kernel (dense content; sparse content runs a
map kernel + record kernel pair per picture step).

Printed JSON (one line, rank 0):
  value      reconstruction-only frames/s: symbol buffers already resident in HBM, the 16
             kernels of a GOP replayed K times, timed with CUDA events on the launching
             stream inside the library (HVQM4BatchReplay), max over ranks.
  e2e        same metric through the C ABI with HOST buffers: bitstreams in host memory ->
             H2D of the raw pictures -> entropy stage on the GPU (entropy.c compiled as device
             code, one warp per picture) -> reconstruction kernels -> D2H of every decoded frame
             into pinned host memory, all inside the timed region.
  e2e_host_entropy  the same with the entropy stage on host threads (north_star's layout):
             host threads -> pinned symbol arena -> H2D -> kernels -> D2H.
  roofline   recon kernel: algorithmic bytes per launch (frame bytes written + reference
             bytes of inter macroblocks + symbol bytes read) / measured launch time,
             against the measured HBM copy peak of MEASURED_PEAKS.json.
  cpu_baseline  the reference decoder (oracle/_ref; else the oracle port) on the host cores,
             bounded sample, N=1 only.
`--impl reference` times that CPU decoder as its own line instead.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "640x480 decoded frames/sec (HVQM4 1.5 I/P/B, multi-stream)"
UNIT = "frames/s"
W, H = 640, 480
GOP = "I" + "PBB" * 5
BASE_SEED = 5000            # stream s uses seed BASE_SEED + s (tests/golden pins s = 0, 1, 511, 1023)
DISTINCT_STREAMS = 128      # distinct bitstreams generated per rank; larger batches reuse them cyclically


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def gen_streams(rank: int, world: int, per_gpu: int, count: int, profile: int):
    """The first `count` distinct bitstreams of this rank's slab of streams (hvqm4_b200/shard.py)."""
    from hvqm4_b200 import shard, synth
    ids = shard.rank_streams(rank, world, per_gpu)[:count]
    return [synth.generate(W, H, 15, GOP, 1, seed=shard.stream_seed(BASE_SEED, s), profile=profile) for s in ids]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, windows):
        rows = [r for t, r in self.rows if any(a <= t <= b for a, b in windows)] or [r for _, r in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


KERNEL_SOURCES = ("recon.cu", "recon_core.h", "recon_dev.cuh", "symbuf.h", "entropy.c")


def kernel_sources_sha() -> str:
    """Hash of the sources the reconstruction kernels and the symbol buffer they read are built from: a committed ncu
    capture is only quoted next to a bench number if it was taken from the same sources."""
    import hashlib
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "hvqm4_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def measured_traffic(profile: int, streams: int):
    """(DRAM bytes per step, note): dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one step,
    averaged over a GOP, from the committed ncu capture of this workload (tools/ncu_traffic.py) -- None when there is
    none for these kernel sources (the capture is stamped with kernel_sources_sha())."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if t.get("kernel_sources_sha") != kernel_sources_sha():
            return None, "profiles/r02_traffic.json was captured from other kernel sources (%s)" % t.get("kernel_sources_sha")
        key = f"{'dense' if profile == 0 else 'realistic'}_{streams}"
        return t["dram_bytes_per_step"].get(key), "ncu capture of these kernel sources (profiles/r02_traffic.json)"
    except (OSError, KeyError, ValueError):
        return None, "no ncu capture committed"


def cpu_reference_run(streams, seconds_target: float, nproc: int):
    """Times the reference CPU decoder (all host cores, one process per core -- it keeps state in
    globals, h4m:604-605) on a bounded sample: every process decodes one stream's GOP `reps` times."""
    from oracle import bindings
    bindings.build(ref=os.path.exists("/root/reference/h4m_audio_decode.c"), port=True)
    if bindings.have_ref():
        dec, kind = bindings.RefDecoder, "reference"
    else:
        dec, kind = bindings.PortDecoder, "port"
    t1, n1 = dec.bench(streams[0], 1)                      # calibration: one GOP on one core
    per_gop = max(t1, 1e-3)
    reps = max(1, int(seconds_target / per_gop))
    wall, cpu_s, frames = dec.bench_mp_streams(streams, nproc, reps)
    return {
        "value": frames / wall, "unit": UNIT, "cores": nproc, "kind": kind,
        "sample": (f"{nproc} processes x {reps} GOPs of 16 pictures each, cycling the same {len(streams)} distinct 640x480 I/P/B bitstreams "
                   f"as the GPU arm (seeds {BASE_SEED}+), wall clock around fork..exit"),
        "single_core_fps": n1 / t1, "wall_s": wall,
    }, reps


def reference_last_frames(files, which):
    """MD5 of the last decoded picture (decode order) of the bitstreams `which`, from the checker -- the unmodified
    reference (oracle/_ref) when it was built, else the oracle port.  Used only AFTER the clock has stopped."""
    import hashlib
    from oracle import bindings
    dec = bindings.RefDecoder if bindings.have_ref() else bindings.PortDecoder
    out = {}
    for k in which:
        last = None
        for _, _, _, yuv in dec(files[k]).frames():
            last = yuv
        out[k] = hashlib.md5(last).hexdigest()
    return out, dec.__name__


def pcie_ceiling(torch, frame_bytes, n_frames, h2d_bytes, barrier):
    """What this box's PCIe link gives this rank while every rank does the same: one step's read-back (n_frames frames,
    pinned host memory) next to one step's upload, on two streams, best of three; GB/s per direction."""
    import hashlib  # noqa: F401
    d2h_n = n_frames * frame_bytes
    try:
        d_src = torch.empty(d2h_n, dtype=torch.uint8, device="cuda")
        h_dst = torch.empty(d2h_n, dtype=torch.uint8, pin_memory=True)
        h_src = torch.empty(max(h2d_bytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        d_dst = torch.empty_like(h_src, device="cuda")
    except RuntimeError:
        return None
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = None
    for _ in range(3):
        barrier()
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        with torch.cuda.stream(s1):
            e0.record()
            h_dst.copy_(d_src, non_blocking=True)
            e1.record()
        with torch.cuda.stream(s2):
            e2.record()
            d_dst.copy_(h_src, non_blocking=True)
            e3.record()
        torch.cuda.synchronize()
        r = (d2h_n / (e0.elapsed_time(e1) * 1e-3) / 1e9, h_src.numel() / (e2.elapsed_time(e3) * 1e-3) / 1e9)
        best = r if best is None or r[0] > best[0] else best
    return {"d2h_gbs_per_gpu": best[0], "h2d_gbs_per_gpu": best[1]}


def make_config(args, world):
    """The workload both arms are measured on (identical keys and values: the driver compares them)."""
    S = streams_per_rank(args, world)
    return {"workload": workload_name(args), "gop": GOP, "profile": "dense" if args.profile == 0 else "realistic",
            "streams_per_gpu": S, "distinct_bitstreams": min(S, DISTINCT_STREAMS), "pictures_per_step": S * len(GOP),
            "scaling": args.scaling}


def streams_per_rank(args, world):
    """weak: --streams-per-gpu on every rank; strong: that many streams in all, split over the ranks (BASELINE config 5 as
    written: 1024 streams sharded across the GPUs)."""
    return args.streams_per_gpu if args.scaling == "weak" else max(1, args.streams_per_gpu // world)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    S = streams_per_rank(args, world)
    streams = gen_streams(0, 1, S, min(S, DISTINCT_STREAMS), args.profile)
    nproc = os.cpu_count() or 1
    t0 = time.perf_counter()
    total_frames, total_wall, base = 0, 0.0, None
    # every step is a bounded sample; the whole arm is kept near 75 s whatever K and W are
    per_step = min(args.ref_seconds, 75.0 / max(1, args.warmup + args.steps))
    for i in range(args.warmup + args.steps):
        base, _ = cpu_reference_run(streams, per_step, nproc)
        if i >= args.warmup:
            total_frames += base["value"] * base["wall_s"]
            total_wall += base["wall_s"]
    value = total_frames / total_wall
    base["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_wall / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": make_config(args, world),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    per = "per GPU" if args.scaling == "weak" else "in all, sharded across the GPUs"
    return (f"BASELINE config 5: {args.streams_per_gpu} independent synthetic 640x480 HVQM4 1.5 I/P/B streams {per} "
            f"(GOP {GOP}, seeds {BASE_SEED}+), one picture per stream per step")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams-per-gpu", type=int, default=1024, help="streams per GPU (weak scaling) or in all (strong scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--gather", default="auto", choices=["auto", "on", "off"],
                    help="end to end: bitstreams fetched by the GPU from registered host memory (auto: from 2 ranks per box on)")
    ap.add_argument("--profile", type=int, default=0, help="0 dense (headline), 1 realistic")
    ap.add_argument("--host-threads", type=int, default=0)
    ap.add_argument("--host-share", type=int, default=-1,
                    help="end to end, GPU entropy stage: streams per GPU whose pictures the host threads parse next to the parse kernel "
                         "(HVQM4BatchSetHostShare); -1 = one stream in eight per 16 host threads (at most one in eight) at one rank, none beyond")
    ap.add_argument("--ref-seconds", type=float, default=12.0, help="CPU seconds per process for the reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-host", action="store_true", help="skip the host-entropy end-to-end leg (experiments)")
    ap.add_argument("--no-realistic", action="store_true", help="skip the secondary realistic-profile measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; hvqm4_b200 has no CPU path (use --impl reference for the CPU decoder)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its banner ("NCCL version ...", when NCCL_DEBUG is
        # set) to file descriptor 1 while the communicator is created, so that happens with fd 1 pointing at stderr
        import torch.distributed as dist
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    # every rank keeps its host threads (bitstream copies, host entropy stage) and its pinned arenas on its own share of
    # the cores: without this the ranks' threads wander over all cores of the box and share caches with each other
    cores = sorted(os.sched_getaffinity(0))
    if world > 1 and len(cores) >= world:
        mine = cores[rank * len(cores) // world:(rank + 1) * len(cores) // world]
        try:
            os.sched_setaffinity(0, mine)
        except OSError:
            mine = cores
    else:
        mine = cores

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if dist:
        dist.barrier()
    from hvqm4_b200 import api

    from hvqm4_b200 import shard
    S = streams_per_rank(args, world)
    distinct = min(S, DISTINCT_STREAMS)
    files = gen_streams(rank, world, S, distinct, args.profile)
    parsed = [api.parse_file(f) for f in files]
    bufs = [ctypes.create_string_buffer(f, len(f) + 8) for f in files]
    bases = [ctypes.addressof(b) for b in bufs]
    # --gather: the bitstreams are page-locked and mapped (HVQM4HostRegister), the GPU entropy stage fetches the pictures
    # itself instead of the host threads copying them into the pinned staging arena (include/hvqm4.h)
    # measured (profiles/r02_n2_ab.txt, r02_share_gather_ab.txt): one rank with 16 cores 105 k frames/s with host copies, 98-102 k
    # with the gather; two ranks with 12 cores each 180 k vs 198 k; eight ranks with 4 cores each 151 k vs 185 k
    gather = args.gather == "on" or (args.gather == "auto" and world >= 2)
    if gather:
        for b, f in zip(bufs, files):
            if api.lib().HVQM4HostRegister(ctypes.addressof(b), len(f)) != 0:
                gather = False
                break
    n_pics = len(parsed[0][1])
    threads = args.host_threads or max(1, len(mine))
    batch = api.Batch(S, W, H, 15, device=local, host_threads=threads)
    ids = list(range(S))
    steps = []
    for k in range(n_pics):
        frs = [parsed[i % distinct][1][k] for i in range(S)]
        steps.append(api.Batch.prepare_step(ids, [f.frame_type for f in frs],
                                            [bases[i % distinct] + frs[i].offset for i in range(S)], [f.bytes for f in frs]))
    ids_arr = (ctypes.c_int32 * S)(*ids)
    frame_bytes = batch.frame_bytes
    pinned = api.lib().HVQM4HostAlloc(S * frame_bytes)
    if not pinned:
        raise SystemExit("HVQM4HostAlloc failed")

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return shard.max_over_ranks(x, dist, "cuda")

    sampler = ClockSampler(local)
    sampler.start()
    windows = []

    # ---- record one GOP: symbol buffers of all 16 steps stay resident in HBM
    batch.record(True)
    for st in steps:
        batch.decode_prepared(st)
    batch.sync()
    batch.record(False)
    stats0 = batch.stats()
    alg_bytes_per_gop = stats0["algorithmic_bytes"]
    sym_bytes_per_gop = stats0["symbol_bytes"]
    inter_frac = stats0["inter_mcbs"] / max(1, stats0["total_mcbs"])

    # ---- reconstruction only (value)
    batch.replay(args.warmup)
    barrier()
    launches0 = api.kernel_launches()
    band0 = batch.stats()["band_launches"]
    t_a = time.perf_counter()
    ms = batch.replay(args.steps)
    t_b = time.perf_counter()
    windows.append((t_a, t_b))
    barrier()
    ms = max_over_ranks(ms)
    recon_launches = api.kernel_launches() - launches0
    fused = batch.stats()["band_launches"] - band0 == recon_launches
    kernel_name = "recon_band_kernel (map work + records fused per band)" if fused else "recon_map_kernel + recon_record_kernel"
    frames_per_step = S * n_pics
    value = shard.job_throughput(frames_per_step * args.steps, world, ms * 1e-3)
    launch_ms = ms / (args.steps * n_pics)
    achieved_gbs = alg_bytes_per_gop / n_pics / (launch_ms * 1e-3) / 1e9
    # B_out + B_ref only (no symbol bytes): every output byte once + 96 reference bytes per inter macroblock
    pixel_gbs = (stats0["pictures"] * frame_bytes + stats0["inter_mcbs"] * 96) / n_pics / (launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_gbs()

    # ---- the clock has stopped: what the timed replay left in the surfaces is checked against the reference decoder
    import hashlib
    check_ids = sorted({0, 1, S // 2, S - 1, (S // 3) | 1, distinct - 1} & set(range(S)))
    want_md5, checker_name = reference_last_frames(files, sorted({i % distinct for i in check_ids})) if rank == 0 else ({}, "")
    parity = {"checker": checker_name, "replay_frames": 0, "e2e_frames": 0}
    if rank == 0:
        for i in check_ids:
            got = hashlib.md5(batch.read_frame(i)).hexdigest()
            if got != want_md5[i % distinct]:
                raise SystemExit(f"bench.py: stream {i}: the picture left by the timed replay differs from {checker_name}")
            parity["replay_frames"] += 1

    # ---- the other scaling curve, reconstruction only: BASELINE config 5 as written shards 1024 streams across the GPUs
    # (strong scaling: 1024 / N streams per rank); the headline line keeps --streams-per-gpu on every rank (weak)
    strong = None
    if world > 1 and args.scaling == "weak":
        S2 = max(1, args.streams_per_gpu // world)
        b2 = api.Batch(S2, W, H, 15, device=local, host_threads=threads)
        ids2 = list(range(S2))
        b2.record(True)
        for k in range(n_pics):
            frs = [parsed[i % distinct][1][k] for i in range(S2)]
            b2.decode(ids2, [f.frame_type for f in frs], [bases[i % distinct] + frs[i].offset for i in range(S2)], [f.bytes for f in frs])
        b2.sync()
        b2.record(False)
        b2.replay(args.warmup)
        barrier()
        l0 = api.kernel_launches()
        ms2 = max_over_ranks(b2.replay(args.steps))
        recon_launches += api.kernel_launches() - l0
        barrier()
        strong = {"value": shard.job_throughput(S2 * n_pics * args.steps, world, ms2 * 1e-3), "unit": UNIT, "scaling": "strong",
                  "streams_per_gpu": S2, "streams_total": S2 * world, "ms_per_step": ms2 / args.steps}
        b2.close()

    # ---- end to end through the C ABI with host buffers, once per entropy-stage placement
    e2e = None
    e2e_host = None
    e2e_launches = 0
    if not args.no_e2e:
        def measure_e2e(b, seconds, what):
            """K GOPs through HVQM4BatchDecode + HVQM4BatchReadFramesAsync: bitstreams in host memory in,
            every decoded frame in pinned host memory out, wall clock, max over ranks."""
            nonlocal e2e_launches

            def one_gop():
                for st in steps:
                    b.decode_prepared(st)
                    b.read_frames_async(ids_arr, S, pinned, frame_bytes)
            up0 = b.stats()["symbol_bytes"]
            tw0 = time.perf_counter()
            one_gop()
            b.sync()
            est = max_over_ranks(time.perf_counter() - tw0)
            h2d = b.stats()["symbol_bytes"] - up0
            k = max(2, min(args.steps, int(seconds / max(est, 1e-3))))
            for _ in range(min(args.warmup, 3) - 1):
                one_gop()
            b.sync()
            barrier()
            l0 = api.kernel_launches()
            t0 = time.perf_counter()
            for _ in range(k):
                one_gop()
            b.sync()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            windows.append((t0, t1))
            sec = max_over_ranks(t1 - t0)
            barrier()
            e2e_launches += api.kernel_launches() - l0
            return {"value": world * frames_per_step * k / sec, "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": frames_per_step * frame_bytes,
                    "ms_per_step": 1e3 * sec / k, "steps": k, "entropy_stage": what}

        # (a) the layout BASELINE.json's north_star describes: entropy stage on host threads
        if not args.no_e2e_host:
            e2e_host = measure_e2e(batch, 12.0, "host threads")
            e2e_host["host_threads_per_gpu"] = threads
            e2e_host["note"] = ("wall clock around K GOPs: host entropy threads -> pinned symbol arena -> H2D -> kernels -> "
                                "D2H of every frame to pinned host memory")
        # (b) the same C source compiled as device code (entropy_dev.cu): raw pictures up, frames down
        # the host cores are idle in this mode: they parse a share of the streams next to the parse kernel (measured on one
        # B200 + 16 cores, profiles/r02_host_share_ab.txt: 106 k frames/s with no share, 110-111 k with 64..160 of 1 024 streams,
        # 104 k with 192, 87 k with 256).  With several ranks per box the host cores are fewer per GPU and the bitstreams are
        # fetched by the GPU (--gather): two ranks with 12 cores each 198 k with no share, 197 k with 96 streams: no share.
        share = args.host_share if args.host_share >= 0 else (min(S * threads // 128, S // 8) if world == 1 else 0)
        share = max(0, min(S, share))
        gb = api.Batch(S, W, H, 15, device=local, host_threads=threads, gpu_entropy=True, host_share=share)
        e2e = measure_e2e(gb, 12.0, "gpu (one warp per picture)" + (f" + host threads for {share} of {S} streams" if share else ""))
        e2e["host_share_streams"] = share
        e2e["note"] = ("wall clock around K GOPs: raw picture bytes H2D -> entropy stage on the GPU -> reconstruction kernels -> "
                       "D2H of every frame to pinned host memory; HVQM4BatchSetEntropyMode(1)" +
                       (f", HVQM4BatchSetHostShare({share})" if share else ""))
        if rank == 0:
            # the pinned buffer holds the last picture of every stream, as read back inside the timed region
            for i in check_ids:
                got = hashlib.md5(ctypes.string_at(pinned + i * frame_bytes, frame_bytes)).hexdigest()
                if got != want_md5[i % distinct]:
                    raise SystemExit(f"bench.py: stream {i}: the frame read back by the timed end-to-end run differs from {checker_name}")
                parity["e2e_frames"] += 1
        gb.close()
        # the link's own ceiling on this box, with every rank copying at once
        ceil = pcie_ceiling(torch, frame_bytes, S, e2e["h2d_bytes_per_step"] // n_pics, barrier)
        if ceil:
            ceil["frames_per_s_ceiling"] = world * ceil["d2h_gbs_per_gpu"] * 1e9 / frame_bytes
            e2e["pcie_ceiling"] = ceil
            e2e["frac_of_ceiling"] = e2e["value"] / ceil["frames_per_s_ceiling"]
            e2e["d2h_gbs_per_gpu"] = e2e["value"] / world * frame_bytes / 1e9
    sampler.stop()
    clocks = sampler.summary(windows)

    # ---- secondary: the same measurement on the realistic content profile (reported, not the headline)
    realistic = None
    if not args.no_realistic and args.profile == 0:
        api.lib().HVQM4HostFree(pinned)
        pinned = None
        batch.close()
        batch = None
        rfiles = gen_streams(rank, world, S, min(S, 32), 1)
        rparsed = [api.parse_file(f) for f in rfiles]
        rbufs = [ctypes.create_string_buffer(f, len(f) + 8) for f in rfiles]
        rbases = [ctypes.addressof(b) for b in rbufs]
        rb = api.Batch(S, W, H, 15, device=local, host_threads=threads)
        rb.record(True)
        for k in range(n_pics):
            frs = [rparsed[i % len(rfiles)][1][k] for i in range(S)]
            rb.decode(ids, [f.frame_type for f in frs], [rbases[i % len(rfiles)] + frs[i].offset for i in range(S)], [f.bytes for f in frs])
        rb.sync()
        rb.record(False)
        rst = rb.stats()
        rb.replay(args.warmup)
        barrier()
        rsteps = max(3, args.steps // 2)
        l0 = api.kernel_launches()
        rms = max_over_ranks(rb.replay(rsteps))
        realistic_launches = api.kernel_launches() - l0
        barrier()
        r_gbs = rst["algorithmic_bytes"] / n_pics / (rms / (rsteps * n_pics) * 1e-3) / 1e9
        realistic = {"value": world * frames_per_step * rsteps / (rms * 1e-3), "unit": UNIT, "achieved": r_gbs, "frac": r_gbs / peak,
                     "inter_mcb_fraction": round(rst["inter_mcbs"] / max(1, rst["total_mcbs"]), 4),
                     "traffic": measured_traffic(1, S)[0], "steps": rsteps}
        recon_launches += realistic_launches
        rb.close()
        if not args.no_e2e:
            # end to end on the same content (GPU entropy stage, every frame read back): PCIe bound
            rsteps_e2e = [api.Batch.prepare_step(ids, [rparsed[i % len(rfiles)][1][k].frame_type for i in range(S)],
                                                 [rbases[i % len(rfiles)] + rparsed[i % len(rfiles)][1][k].offset for i in range(S)],
                                                 [rparsed[i % len(rfiles)][1][k].bytes for i in range(S)]) for k in range(n_pics)]
            rpinned = api.lib().HVQM4HostAlloc(S * frame_bytes)
            rg = api.Batch(S, W, H, 15, device=local, host_threads=threads, gpu_entropy=True)

            def r_gop():
                for st in rsteps_e2e:
                    rg.decode_prepared(st)
                    rg.read_frames_async(ids_arr, S, rpinned, frame_bytes)
            r_gop()
            rg.sync()
            barrier()
            l0 = api.kernel_launches()
            kr = max(2, args.steps // 4)
            t0 = time.perf_counter()
            for _ in range(kr):
                r_gop()
            rg.sync()
            sec = max_over_ranks(time.perf_counter() - t0)
            barrier()
            e2e_launches += api.kernel_launches() - l0
            realistic["e2e"] = {"value": world * frames_per_step * kr / sec, "unit": UNIT, "d2h_bytes_per_step": frames_per_step * frame_bytes,
                                "d2h_gbs_per_gpu": frames_per_step * frame_bytes * kr / sec / 1e9, "steps": kr, "entropy_stage": "gpu (one warp per picture)"}
            rg.close()
            api.lib().HVQM4HostFree(rpinned)

    # ---- BASELINE config 2: one 640x480 I/P GOP-15 stream through the SDK entry points (latency bound)
    single = None
    if rank == 0 and not args.no_e2e:
        from hvqm4_b200 import synth
        one = synth.generate(W, H, 15, "I" + "P" * 14, 2, seed=102, profile=args.profile)
        ptr, disp, ftype = ctypes.c_void_p(), ctypes.c_uint32(), ctypes.c_uint32()
        for _ in range(2):                              # first pass warms allocations and the driver
            # the reference program's loop in C (player.cpp: demux, rotation, HVQM4Decode?pic into host frame
            # buffers); the frame pointer is handed back, nothing is copied on the Python side
            pl = api.FilePlayer(one)
            nxt = lambda: api.lib().HVQM4PlayerNextFrame(pl._h, ctypes.byref(ptr), ctypes.byref(disp), ctypes.byref(ftype))
            assert nxt() == 1                           # the first picture allocates the device twins
            t0 = time.perf_counter()
            nfr = 0
            while nxt() == 1:
                nfr += 1
            t_sdk = time.perf_counter() - t0
            assert pl.errors() == 0
            pl.close()
        single = {"workload": "BASELINE config 2: one synthetic 640x480 HVQM4 1.5 I/P GOP-15 stream, HVQM4Decode?pic with host buffers",
                  "sdk_fps": nfr / t_sdk, "frames": nfr}
        if not args.no_cpu_baseline:
            from oracle import bindings
            dec = bindings.RefDecoder if bindings.have_ref() else bindings.PortDecoder
            tr, nr = dec.bench(one, 2)
            single["reference_cpu_fps_1core"] = nr / tr

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_reference_run(files, args.ref_seconds, os.cpu_count() or 1)

    traffic, traffic_note = measured_traffic(args.profile, S)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u8/int32", "data": "synthetic",
            "config": make_config(args, world),
            "details": {"launches_per_step": int(recon_launches // max(1, args.steps)), "inter_mcb_fraction": round(inter_frac, 4),
                        "bitstreams_fetched_by_gpu": gather, "host_threads_per_gpu": threads, "host_cores_of_rank0": f"{mine[0]}-{mine[-1]}", "symbol_bytes_per_picture": round(sym_bytes_per_gop / frames_per_step),
                        "l2": "inputs larger than L2: one step touches %.0f MB of symbols + %.0f MB of surfaces per GPU"
                              % (sym_bytes_per_gop / 1e6, 4 * S * frame_bytes / 1e6)},
            "parity_checked": parity,
            "mpixel_per_s": value * W * H / 1e6,
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peak, "unit": "GB/s", "frac": achieved_gbs / peak,
                         "frac_pixel_only": pixel_gbs / peak, "achieved_pixel_only": pixel_gbs,
                         "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src, "kernel": kernel_name, "launch": "one step = one picture of every stream",
                         "algorithmic_bytes_per_launch": alg_bytes_per_gop / n_pics, "launch_ms": launch_ms,
                         "frac_of_nominal_8TBs": achieved_gbs / 8000.0},
            "e2e": e2e, "gpu_launches": int(recon_launches + e2e_launches), "clocks": clocks,
        }
        if strong:
            line["strong_scaling"] = strong
        if e2e_host:
            line["e2e_host_entropy"] = e2e_host
        if realistic:
            line["realistic_profile"] = realistic
        if single:
            line["single_stream"] = single
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)

    if pinned:
        api.lib().HVQM4HostFree(pinned)
    if batch:
        batch.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

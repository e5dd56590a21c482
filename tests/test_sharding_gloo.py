"""N>1 bookkeeping of bench.py on CPU: two gloo ranks shard the streams with no data-path
collective; only the max-over-ranks of the timed interval crosses ranks."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hvqm4_b200 import shard, synth
    from oracle import bindings
    per_gpu = 2
    mine = shard.rank_streams(rank, world, per_gpu)
    # every rank decodes only its own streams (here with the CPU checker -- the product has no CPU path)
    md5s = {}
    for s in mine:
        data = synth.generate(320, 240, 15, "IPB", 1, seed=shard.stream_seed(5000, s), profile=1)
        md5s[s] = [m for _, _, m in bindings.PortDecoder.md5s(data)]
    elapsed = 1.0 + rank            # rank 1 is "slower"
    worst = shard.max_over_ranks(elapsed, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, md5s))      # test-only gather to check the partition
    if rank == 0:
        out.put((worst, gathered, shard.job_throughput(per_gpu * 3, world, worst)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo(oracle):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    worst, gathered, fps = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert worst == 2.0                                  # max over ranks, not mean
    ids = [s for mine, _ in gathered for s in mine]
    assert sorted(ids) == [0, 1, 2, 3] and len(set(ids)) == 4       # disjoint, complete
    assert fps == 2 * 3 * 2 / 2.0
    # distinct seeds -> distinct streams -> distinct pictures; same seed -> same pictures on any rank
    all_md5 = {s: m for _, d in gathered for s, m in d.items()}
    assert len({tuple(m) for m in all_md5.values()}) == 4


def test_rank_streams_and_seeds():
    sys.path.insert(0, ROOT)
    from hvqm4_b200 import shard
    assert shard.rank_streams(3, 8, 128) == list(range(384, 512))
    assert shard.stream_seed(5000, 1023) == 6023
    assert shard.stream_seed(5000, 130, distinct=128) == 5002

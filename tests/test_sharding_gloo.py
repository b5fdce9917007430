"""The N>1 host path on CPU: two gloo ranks shard a clip the way main_func does, agree on coverage,
and reduce their timings with max-over-ranks (what bench.py does under torchrun)."""
import os
import socket

import pytest

from vr_video_generator_b200 import shard, tables


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r = shard.Ranks(backend="gloo")
    mine = shard.shard_for_rank(0, 10 ** 14, 18000, r.world, r.rank)     # config 5: 10 min of 1080p30
    r.barrier()
    everyone = r.gather(mine)
    slowest = r.max(1.0 + rank)                                        # rank 1 is "slower"
    frames = sum(e - b for b, e in everyone)
    q.put((rank, mine, everyone, slowest, frames / slowest))
    r.close()


def test_two_ranks_shard_a_clip():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(i, 2, port, q)) for i in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == (0, 9000) and got[1][1] == (9000, 18000)
    for rank, mine, everyone, slowest, fps in got:
        assert everyone == [(0, 9000), (9000, 18000)]
        assert slowest == 2.0 and fps == 9000.0                           # whole-job frames / max time


def test_shard_and_flush_bookkeeping():
    assert shard.shard_for_rank(0, 10 ** 14, 100, 8, 7) == (91, 104)
    assert shard.shard_for_rank(0, 10 ** 14, 3, 8, 5) is None
    # nibba_woka: flush every Max_Frame_Count frames and at the end; names are {first}_{last}.mp4
    # names and frame counts exactly as nibba_woka's loop produces them (PredictAndGenerate.py:221-250): frame i-1 is
    # appended at iteration i, so "0_15" holds the 15 frames 0..14 and the next clip is named from 16
    assert shard.flush_ranges(0, 25, 100, 15) == [(0, 15, 15), (16, 24, 10)]
    assert shard.flush_ranges(91, 104, 100, 15) == [(91, 99, 9)]
    assert shard.subclip_name(16, 24) == "16_24.mp4"
    # same names as the worker loop writes, every frame of every range exactly once
    from vr_video_generator_b200 import worker
    clips = [c for b, e in tables.clip_ranges(0, 10 ** 14, 100, 4) for c in shard.flush_ranges(b, e, 100, 15)]
    assert sum(c[2] for c in clips) == 100 and clips[0][0] == 0 and clips[-1][1] == 99
    assert worker.order_subclips([shard.subclip_name(a, b) for a, b, _ in reversed(clips)]) == [shard.subclip_name(a, b) for a, b, _ in clips]

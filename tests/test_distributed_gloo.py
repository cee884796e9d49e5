"""world_size-2 gloo run of the multi-GPU host logic on CPU: shard, fill own rows, in-place all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from video_analytics_b200.distributed import gather_video_rows, shard_bounds, trim_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_videos, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi, per = shard_bounds(n_videos, rank, world)
    bufs = {"video_scores": torch.zeros(world * per, 101), "score_pred": torch.full((world * per,), -1, dtype=torch.int32)}
    for v in range(lo, hi):                      # "evaluate" the local shard: row content is a function of the video id
        bufs["video_scores"][rank * per + (v - lo)] = torch.arange(101, dtype=torch.float32) + 1000.0 * v
        bufs["score_pred"][rank * per + (v - lo)] = v % 101
    gather_video_rows(bufs, rank, world, per)
    out = trim_rows(bufs, n_videos)
    ok = all(float(out["video_scores"][v, 0]) == 1000.0 * v and int(out["score_pred"][v]) == v % 101 for v in range(n_videos))
    q.put((rank, ok, out["video_scores"].shape[0]))
    dist.destroy_process_group()


def test_two_rank_gather_cpu():
    world, n_videos = 2, 7                        # odd count -> the last rank carries a padding row
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_videos, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True, n_videos), (1, True, n_videos)]

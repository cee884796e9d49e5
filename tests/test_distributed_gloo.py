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


def test_bench_reference_arm_contract_under_torchrun():
    """`bench.py --impl reference` launched the driver's way with 2 ranks: rank 0 alone prints ONE JSON line carrying the
    contract keys, the other rank exits 0 without work (both workloads; tiny samples)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra, metric in (([], "two-stream snippets/sec"), (["--workload", "train"], "two-stream training snippets/sec")):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29655", os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
               "--warmup", "0", "--ref-snippets", "1"] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        assert len(lines) == 1, r.stdout[-2000:]
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["metric"].startswith(metric) and d["value"] > 0
        for key in ("unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config",
                    "cpu_baseline", "e2e"):
            assert key in d, key
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def _sampler_worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from video_analytics_b200.utils import RankShardSampler
    torch.manual_seed(100 + rank)                      # ranks are seeded DIFFERENTLY (own crops / dropout masks)
    s = RankShardSampler(n_items, True, rank, world)
    epochs = [list(iter(s)) for _ in range(2)]
    plain = list(iter(RankShardSampler(n_items, False, rank, world)))
    q.put((rank, epochs, plain, len(s)))
    dist.destroy_process_group()


def test_rank_shard_sampler_partitions_every_epoch():
    """Data-parallel loaders (utils.getDataLoader under a process group): the ranks' index lists partition each epoch's
    permutation -- no repeated batches across ranks, even with different per-rank seeds -- and have equal lengths."""
    world, n_items = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sampler_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, e0, p0, l0), (_, e1, p1, l1) = results
    assert l0 == l1 == 6
    for a, b in zip(e0, e1):
        assert len(a) == len(b) == 6
        assert sorted(set(a + b)) == list(range(n_items))          # union covers the data set (one wrapped duplicate)
        assert len(set(a) & set(b)) <= 1
    assert e0[0] != e0[1] or e1[0] != e1[1]                          # a new permutation per epoch
    assert p0 == [0, 2, 4, 6, 8, 10] and p1 == [1, 3, 5, 7, 9, 0]

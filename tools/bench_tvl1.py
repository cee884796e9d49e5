"""Throughput of va_tvl1_flow on synthetic 340x256 clips: frame pairs per second, device-timed with
CUDA events, plus the inner-iteration counts (the work is data dependent: the solver stops when the update falls under
epsilon).  `--oracle` also times oracle/tvl1.py (numpy, 1 core) on one pair as the CPU baseline beside it."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analytics_b200 import flow  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=9)
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--h", type=int, default=256)
    ap.add_argument("--w", type=int, default=340)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--epsilon", type=float, default=0.01)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--cycles", action="store_true", help="per (level, warp) cycle counters of cluster 0's first pair")
    a = ap.parse_args()
    clip = flow.synthetic_clip(a.frames, a.h, a.w, seed=7)
    dev = torch.device("cuda")
    frames = torch.from_numpy(clip).to(dev)
    n = a.pairs
    base = torch.arange(n, dtype=torch.int32, device=dev) % (a.frames - 1)
    k = torch.arange(n, dtype=torch.int32, device=dev)
    table = torch.stack([base, base + 1, k, k + n], dim=1).contiguous()
    out = torch.empty((2 * n, a.h, a.w), dtype=torch.uint8, device=dev)
    p = flow.TVL1Params(epsilon=a.epsilon)
    res = flow.tvl1(frames, (a.h, a.w, 3), table, out, params=p, return_iterations=True)
    torch.cuda.synchronize()
    its = res["iterations"].cpu().numpy()
    times = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flow.tvl1(frames, (a.h, a.w, 3), table, out, params=p)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    levels = p.levels(a.h, a.w)
    sizes = [(a.h, a.w)]
    for _ in range(1, levels):
        sizes.append((int(round(sizes[-1][0] * 0.8)), int(round(sizes[-1][1] * 0.8))))
    px = np.array([s[0] * s[1] for s in sizes[::-1]], np.float64)                   # processing order: coarsest first
    pix_iters = float((its.reshape(n, levels, p.warps).sum(2) * px[None, :]).sum())
    line = {"what": "va_tvl1_flow", "image": [a.h, a.w], "pairs": n, "ms": ms, "pairs_per_s": n / ms * 1e3,
            "inner_iterations_per_pair": float(its.sum(1).mean()), "pixel_iterations_per_s": pix_iters / ms * 1e3,
            "epsilon": a.epsilon, "times_ms": times}
    if a.cycles:
        from video_analytics_b200 import _lib
        dbg = torch.zeros(2 * levels * p.warps, dtype=torch.int64, device=dev)
        _lib.check(_lib.load().va_tvl1_debug_cycles(_lib.ptr(dbg)))
        flow.tvl1(frames, (a.h, a.w, 3), table, out, params=p)
        torch.cuda.synchronize()
        _lib.check(_lib.load().va_tvl1_debug_cycles(None))
        d = dbg.cpu().numpy().reshape(levels, p.warps, 2)
        it0 = its[0].reshape(levels, p.warps)
        line["cycles_first_pair"] = {"level_sizes": sizes[::-1], "iterations": it0.tolist(), "warp_phase_cycles": d[:, :, 0].tolist(),
                                     "inner_cycles": d[:, :, 1].tolist(),
                                     "cycles_per_iteration": (d[:, :, 1] / np.maximum(it0, 1)).round(0).tolist()}
    if a.oracle:
        from oracle import tvl1 as otv
        t = time.time()
        otv.tvl1_flow(otv.gray_from_rgb(clip[0]), otv.gray_from_rgb(clip[1]))
        line["oracle_numpy_s_per_pair"] = time.time() - t
    print(json.dumps(line))


if __name__ == "__main__":
    main()

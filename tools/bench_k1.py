"""K1 / K4 alone on the GPU: ms per launch and bytes moved (the aux_rooflines figures of bench.py without the rest of
the bench).  `python tools/bench_k1.py [--videos 4] [--once]`; --once = one launch per kernel (for ncu)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analytics_b200 import ops                                        # noqa: E402
from video_analytics_b200.evaluate import spatial_table, temporal_table    # noqa: E402
from video_analytics_b200.store import DeviceStore, make_layout             # noqa: E402
from video_analytics_b200.parameters import NORM_MEANS_TF, NORM_STDS_TF          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=4)
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    lay = make_layout(a.videos)
    store = DeviceStore(lay)
    ts = torch.cat([torch.from_numpy(spatial_table(m, lay.rgb_shape)) for m in lay.videos]).cuda()
    tt = torch.cat([torch.from_numpy(temporal_table(m, lay.flow_shape)) for m in lay.videos]).cuda()
    n = ts.shape[0]
    out_s = torch.empty((n, 224, 224, 16), dtype=torch.bfloat16, device="cuda")
    out_t = torch.empty((n, 224, 224, 32), dtype=torch.bfloat16, device="cuda")
    f_s = lambda: ops.preprocess(store.rgb, lay.rgb_shape, ts, list(NORM_MEANS_TF), list(NORM_STDS_TF), c_pad=16, out=out_s)
    f_t = lambda: ops.preprocess(store.flow, lay.flow_shape, tt, [0.485] * 20, [0.229] * 20, c_pad=32, out=out_t)
    V, D, C = 512, 256, 101
    g = torch.Generator(device="cuda").manual_seed(1)
    fd = [torch.rand((V * 250, D), device="cuda", generator=g) for _ in range(2)]
    fs = [torch.rand((V * 250, C), device="cuda", generator=g) for _ in range(2)]
    offs = torch.arange(0, (V + 1) * 250, 250, dtype=torch.int32, device="cuda")
    res = {}
    f_f = lambda: ops.fuse(fd[0], fd[1], fs[0], fs[1], offs, out=res)
    px = 224 * 224
    for name, fn, moved in (("K1 rgb", f_s, n * (px * 3 + px * 32)), ("K1 flow", f_t, n * (px * 20 + px * 64)),
                            ("K4 fuse", f_f, V * (714_000 + (2 * D + C) * 4 + 4))):
        fn(); torch.cuda.synchronize()
        if a.once:
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name}: {ms:.4f} ms per launch, {moved / ms / 1e6:.0f} GB/s moved ({n} snippets / {V} videos)")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-role stall breakdown of the layer kernel (CTA 0) for selected layers: which of producer / MMA / epilogue is
the bottleneck.  python tools/diag_roles.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_analytics_b200 import _lib, ops

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 125
lib = _lib.load()
cnt = torch.zeros(16, dtype=torch.int64, device="cuda")
cfg = [("conv1_1 spatial", 224, 16, 3, 64, 0, 0, 0), ("conv1_1 temporal", 224, 32, 20, 64, 0, 0, 0),
       ("conv1_1 spatial r=1", 224, 16, 3, 64, 0, 0, 1),
       ("conv1_2", 224, 64, 64, 64, 1, 0, 0), ("conv1_2 r=1", 224, 64, 64, 64, 1, 0, 1),
       ("conv2_1", 112, 64, 64, 128, 0, 0, 0), ("conv2_2", 112, 128, 128, 128, 1, 0, 0),
       ("conv3_2", 56, 256, 256, 256, 0, 0, 0), ("conv4_2", 28, 512, 512, 512, 0, 0, 0), ("conv5_1", 14, 512, 512, 512, 0, 0, 0),
       ("conv3_2 1-CTA", 56, 256, 256, 256, 0, 0, 1), ("conv4_2 1-CTA", 28, 512, 512, 512, 0, 0, 1),
       ("conv5_1 1-CTA", 14, 512, 512, 512, 0, 0, 1), ("conv3_1", 56, 128, 128, 256, 0, 0, 0), ("conv3_1 1-CTA", 56, 128, 128, 256, 0, 0, 1)]
for (name, H, cin_pad, cin, cout, pool, bn, r) in cfg:
    x = torch.randn(batch, H, H, cin_pad, device="cuda").bfloat16()
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    b = torch.zeros(cout, device="cuda")
    ops.conv2d_nhwc(x, w, b, pool=bool(pool), force_bn=bn, force_r=r)
    lib.va_debug_conv_counters(_lib.ptr(cnt))
    cnt.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv2d_nhwc(x, w, b, pool=bool(pool), force_bn=bn, force_r=r)
    e1.record()
    torch.cuda.synchronize()
    lib.va_debug_conv_counters(None)
    c = cnt.cpu().tolist()
    tiles = max(1, c[11])
    print(f"{name:22s} {e0.elapsed_time(e1):7.3f} ms tiles/CTA {tiles:5d} clk/tile {c[2]/tiles:8.0f} | producer wait-empty {100*c[1]/max(1,c[0]):5.1f}% | "
          f"MMA wait-full {100*c[3]/max(1,c[2]):5.1f}% wait-tempty {100*c[4]/max(1,c[2]):5.1f}% | "
          f"epi0 wait-tfull {100*c[6]/max(1,c[5]):5.1f}% wait-staging {100*c[7]/max(1,c[5]):5.1f}% | "
          f"epi1 wait-tfull {100*c[9]/max(1,c[8]):5.1f}% wait-staging {100*c[10]/max(1,c[8]):5.1f}%")

#!/usr/bin/env python
"""Runs the HBM-bound kernels of the path a few times (for ncu captures): preprocess (RGB, flow), head, fuse."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_analytics_b200 import ops
from video_analytics_b200.evaluate import spatial_table, temporal_table
from video_analytics_b200.store import DeviceStore, make_layout

lay = make_layout(4)
store = DeviceStore(lay)
ts = torch.cat([torch.from_numpy(spatial_table(m, lay.rgb_shape)) for m in lay.videos]).cuda()
tt = torch.cat([torch.from_numpy(temporal_table(m, lay.flow_shape)) for m in lay.videos]).cuda()
for _ in range(3):
    xs = ops.preprocess(store.rgb, lay.rgb_shape, ts, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], c_pad=16)
    xt = ops.preprocess(store.flow, lay.flow_shape, tt, [0.485] * 20, [0.229] * 20, c_pad=32)
V, D, C = 512, 256, 101
g = torch.Generator(device="cuda").manual_seed(0)
ds, dt = torch.rand(V * 250, D, device="cuda", generator=g), torch.rand(V * 250, D, device="cuda", generator=g)
ss, st = torch.rand(V * 250, C, device="cuda", generator=g), torch.rand(V * 250, C, device="cuda", generator=g)
offs = torch.arange(0, (V + 1) * 250, 250, dtype=torch.int32, device="cuda")
W = torch.randn(C, 2 * D, dtype=torch.float64, device="cuda")
b = torch.randn(C, dtype=torch.float64, device="cuda")
for _ in range(3):
    ops.fuse(ds, dt, ss, st, offs, svm_w=W, svm_b=b)
torch.cuda.synchronize()
print("ok")

#!/usr/bin/env python
"""GPU diagnostic: full-stream forward (va_forward) vs torchvision VGG16 fp32 on the same random-init weights,
plus per-layer timings of the tcgen05 kernel at a given batch.  python tools/diag_forward.py [batch]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from video_analytics_b200 import ops


def build_ref(cin, desc_dim=256, n_classes=101, seed=0):
    import torchvision.models as models
    torch.manual_seed(seed)
    m = models.vgg16(weights=None)
    if cin != 3:
        m.features[0] = nn.Conv2d(cin, 64, kernel_size=3, padding=1)
    m.classifier = nn.Sequential(nn.Linear(25088, 4096), nn.ReLU(True), nn.Dropout(), nn.Linear(4096, 4096), nn.ReLU(True),
                                 nn.Dropout(), nn.Linear(4096, desc_dim), nn.ReLU(True), nn.Dropout(),
                                 nn.Linear(desc_dim, n_classes))
    return m.eval()


def check_stream(cin, n=3):
    m = build_ref(cin).cuda()
    net = ops.StreamNet(0 if cin == 3 else 1, cin, max_batch=2)   # max_batch 2 -> exercises chunking with n=3
    net.load_state_dict(m.state_dict())
    g = torch.Generator().manual_seed(7)
    xc = torch.randn(n, cin, 224, 224, generator=g).cuda().bfloat16()
    x = torch.zeros(n, 224, 224, net.c_pad, dtype=torch.bfloat16, device="cuda")
    x[..., :cin] = xc.permute(0, 2, 3, 1)
    desc, logits, probs, pred = net.forward(x)
    torch.cuda.synchronize()
    with torch.no_grad():
        f = m.features(xc.float()).flatten(1)
        d_ref = m.classifier[:9](f)
        l_ref = m.classifier[9:](d_ref)
        p_ref = torch.softmax(l_ref, 1)
    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
    print(f"[stream cin={cin}] desc rel {rel(desc, d_ref):.3e}  logits rel {rel(logits, l_ref):.3e} "
          f"probs maxrel {float(((probs - p_ref).abs() / p_ref).max()):.3e}  pred agree "
          f"{int((pred.long() == l_ref.argmax(1)).sum())}/{n}  desc_absmean {float(d_ref.abs().mean()):.3e}")
    net.close()


def time_layers(batch):
    """Per-layer device time of the layer kernel at one chunk size, default (auto) variant first."""
    cfg = [(224, 16, 3, 64, 0), (224, 32, 20, 64, 0), (224, 64, 64, 64, 1), (112, 64, 64, 128, 0), (112, 128, 128, 128, 1),
           (56, 128, 128, 256, 0), (56, 256, 256, 256, 0), (56, 256, 256, 256, 1), (28, 256, 256, 512, 0), (28, 512, 512, 512, 0),
           (28, 512, 512, 512, 1), (14, 512, 512, 512, 0), (14, 512, 512, 512, 1)]
    tot = 0.0
    for (H, cin_pad, cin, cout, pool) in cfg:
        x = torch.randn(batch, H, H, cin_pad, device="cuda").bfloat16()
        w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
        b = torch.zeros(cout, device="cuda")
        variants = [(0, 0)]
        if H >= 112:
            variants += [(0, 1), (0, 3)] if cin_pad < 64 else [(0, 1)]
        if cout >= 256:
            variants += [(128, 0)] if H > 14 else [(128, 0)]
        for (bn, r) in variants:
            for _ in range(2):
                ops.conv2d_nhwc(x, w, b, pool=bool(pool), force_bn=bn, force_r=r)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                ops.conv2d_nhwc(x, w, b, pool=bool(pool), force_bn=bn, force_r=r)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 4
            fl = 2.0 * batch * H * H * cout * 9 * cin
            print(f"  conv H={H} {cin}->{cout} pool={pool} bn={bn} r={r}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s (algorithmic)")
            if (bn, r) == (0, 0) and cin != 20:
                tot += ms * (2 if (H, cin, pool) in ((56, 256, 0), (28, 512, 0), (14, 512, 0)) else 1)
    print(f"  spatial-stream conv total, default variants: {tot:.3f} ms for batch {batch} = {tot/batch*1e3:.1f} us/snippet")
    # FC layers
    for (fin, fout, f32) in ((25088, 4096, False), (4096, 4096, False), (4096, 256, True)):
        x = torch.randn(batch, fin, device="cuda").bfloat16()
        w = torch.randn(fout, fin, device="cuda") * 0.01
        b = torch.zeros(fout, device="cuda")
        for _ in range(2):
            ops.linear(x, w, b, out_f32=f32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            ops.linear(x, w, b, out_f32=f32)
        e1.record()
        torch.cuda.synchronize()
        print(f"  fc {fin}->{fout}: {e0.elapsed_time(e1)/4:.3f} ms (includes the fp32->bf16 weight pack of this test entry)")


if __name__ == "__main__":
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    print(ops.device_info())
    check_stream(3)
    check_stream(20)
    time_layers(batch)

"""One weight-gradient GEMM per VGG16 layer class at batch 64 (for ncu captures of wgrad_tc_kernel)."""
import sys

import torch

sys.path.insert(0, ".")
from video_analytics_b200 import train_ops as T  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
n = 64
for (H, cin, cin_pad, cout) in ((56, 256, 256, 256), (112, 128, 128, 128), (224, 64, 64, 64), (224, 3, 16, 64)):
    x = torch.randn(n, H, H, cin_pad, device="cuda", generator=g).bfloat16()
    if cin_pad != cin:
        x[..., cin:] = 0
    dz = torch.randn(n, H, H, cout, device="cuda", generator=g).bfloat16()
    for _ in range(2):
        dw = T.conv2d_wgrad(dz, x, cin)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dw = T.conv2d_wgrad(dz, x, cin)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * n * H * H * cout * cin * 9
    print(f"wgrad n={n} {H}x{H} {cin}->{cout}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s algorithmic")
    del x, dz, dw

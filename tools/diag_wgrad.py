"""Diagnostic: one weight-gradient GEMM case per process (a device fault poisons the context for later cases)."""
import sys
import torch
import torch.nn.functional as F
from video_analytics_b200 import train_ops as T

n, H, W, cin, cin_pad, cout = map(int, sys.argv[1:7])
g = torch.Generator().manual_seed(5)
xc = torch.randn(n, cin, H, W, generator=g).cuda().bfloat16().float()
dz = torch.randn(n, cout, H, W, generator=g).cuda().bfloat16().float()
x = torch.zeros(n, H, W, cin_pad, dtype=torch.bfloat16, device="cuda")
x[..., :cin] = xc.permute(0, 2, 3, 1)
w = torch.zeros(cout, cin, 3, 3, device="cuda", requires_grad=True)
F.conv2d(xc, w, None, padding=1).backward(dz)
dw = T.conv2d_wgrad(dz.permute(0, 2, 3, 1).contiguous().bfloat16(), x, cin)
torch.cuda.synchronize()
rel = float((dw - w.grad).norm() / w.grad.norm())
print(sys.argv[1:7], "rel", rel, "max", float((dw - w.grad).abs().max()))

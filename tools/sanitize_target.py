#!/usr/bin/env python
"""Small cases of every tcgen05 / TMA kernel variant, for compute-sanitizer (one --tool per gpurun call):
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_target.py
Each case is checked against torch fp32 so that a sanitizer-clean run is also a correct run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from video_analytics_b200 import ops
from video_analytics_b200 import train_ops as T

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(1)
bad = 0


def check(name, y, ref, tol=2e-2):
    global bad
    err = float((y.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-6))
    ok = err < tol
    bad += 0 if ok else 1
    print(f"{name:46s} rel err {err:.2e} {'ok' if ok else 'MISMATCH'}", flush=True)


def conv_case(name, n, H, cin, cin_pad, cout, pool, force_r, force_bn=0):
    xc = torch.randn(n, cin, H, H, generator=g).cuda().bfloat16()
    x = torch.zeros(n, H, H, cin_pad, dtype=torch.bfloat16, device="cuda")
    x[..., :cin] = xc.permute(0, 2, 3, 1)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
    b = (torch.randn(cout, generator=g) * 0.1).cuda()
    y = ops.conv2d_nhwc(x, w, b, relu=True, pool=pool, force_bn=force_bn, force_r=force_r)
    ref = torch.relu(torch.nn.functional.conv2d(xc.float(), w.bfloat16().float(), b, padding=1))
    if pool:
        ref = torch.nn.functional.max_pool2d(ref, 2, 2)
    check(name, y, ref.permute(0, 2, 3, 1))


conv_case("conv_tc_kernel R=1 BN=64", 2, 16, 64, 64, 64, False, 1)
conv_case("conv_tc_kernel R=3 WRES BN=64 pool", 2, 16, 64, 64, 64, True, 3)
conv_case("conv_tc_kernel HALO", 2, 16, 64, 64, 64, True, 10)
conv_case("conv_tc2h_kernel<64> (pair HALO, pool)", 3, 16, 64, 64, 64, True, 11)
conv_case("conv_tc2h_kernel<128> 64->128", 2, 16, 64, 64, 128, False, 11)
conv_case("conv_tc2h_kernel<128> 128->128 pool", 2, 16, 128, 128, 128, True, 11)
conv_case("conv_tc_kernel S=3 first layer (3ch)", 2, 16, 3, 16, 64, False, 0)
conv_case("conv_tc_kernel S=3 first layer (20ch)", 2, 16, 20, 32, 64, False, 0)
conv_case("conv_tc2_kernel (pair, BN=256)", 3, 8, 128, 128, 256, False, 0)
conv_case("conv_tc_kernel BN=256 1-CTA", 3, 8, 128, 128, 256, True, 1)
conv_case("conv_tc2_kernel 14x14-like tile (4x4)", 5, 4, 256, 256, 512, True, 0)

x = torch.randn(70, 512, generator=g).cuda().bfloat16()
w = (torch.randn(256, 512, generator=g) / 512 ** 0.5).cuda()
b = torch.randn(256, generator=g).cuda() * 0.1
check("linear bf16 out", ops.linear(x, w, b, relu=True), torch.relu(x.float() @ w.bfloat16().float().t() + b))
check("linear fp32 out", ops.linear(x, w, b, relu=True, out_f32=True), torch.relu(x.float() @ w.bfloat16().float().t() + b))

# fused gather + conv1_1 on a tiny store
from video_analytics_b200.store import DeviceStore, make_layout
lay = make_layout(1)
store = DeviceStore(lay)
for (images, shape, planes, mean, std) in ((store.rgb, lay.rgb_shape, 1, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
                                          (store.flow, lay.flow_shape, 20, [0.485] * 20, [0.229] * 20)):
    table = torch.tensor([[[p % 5, 3 + p, 7 + 2 * p, p & 1] for p in range(planes)]], dtype=torch.int32).cuda()
    cin = planes * shape[2]
    w = (torch.randn(64, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
    b = (torch.randn(64, generator=g) * 0.1).cuda()
    y = ops.conv1_fused(images, shape, table, mean, std, w, b)
    xr = ops.preprocess(images, shape, table, mean, std, reference_layout=True)
    ref = torch.relu(torch.nn.functional.conv2d(xr.bfloat16().float(), w.bfloat16().float(), b, padding=1)).permute(0, 2, 3, 1)
    check(f"conv1_fused_kernel ({cin} channels)", y, ref)

# weight-gradient GEMM (wgrad_tc / wgrad_tc2)
for (n, H, cin, cout) in ((2, 16, 64, 64), (3, 8, 256, 256)):
    xc = torch.randn(n, cin, H, H, generator=g).cuda().bfloat16()
    dz = torch.randn(n, cout, H, H, generator=g).cuda().bfloat16()
    xn, dzn = xc.permute(0, 2, 3, 1).contiguous(), dz.permute(0, 2, 3, 1).contiguous()
    dw = T.conv2d_wgrad(dzn, xn, cin)
    wref = torch.zeros(cout, cin, 3, 3, device="cuda", requires_grad=True)
    torch.nn.functional.conv2d(xc.float(), wref, padding=1).backward(dz.float())
    check(f"wgrad_tc {cin}->{cout} at {H}x{H}", dw, wref.grad)
torch.cuda.synchronize()
print("SANITIZE TARGET DONE bad=%d" % bad, flush=True)
sys.exit(1 if bad else 0)

"""K2/K3: the tcgen05 implicit-GEMM layer kernel vs torch fp32 on bf16-rounded operands (per layer, every kernel
variant, every tile shape of the VGG16 path, partial batches)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# bf16 output rounding is 2^-8 relative; operands are identical, accumulation is fp32 on both sides
RTOL = 2.0 ** -7


def _check(ours, ref):
    ours = ours.float()
    tol = RTOL * ref.abs() + 2e-2 * ref.abs().mean().clamp_min(1e-6)
    bad = (ours - ref).abs() > tol
    assert not torch.isnan(ours).any()
    assert not bad.any(), f"{int(bad.sum())} of {bad.numel()} outside tolerance; max abs err {float((ours - ref).abs().max())}"


def _conv_case(n, H, cin, cin_pad, cout, pool, bn, r, seed=0, W=None):
    from video_analytics_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    W = W or H
    g = torch.Generator().manual_seed(seed)
    xc = torch.randn(n, cin, H, W, generator=g).cuda().bfloat16()
    x = torch.zeros(n, H, W, cin_pad, dtype=torch.bfloat16, device="cuda")
    x[..., :cin] = xc.permute(0, 2, 3, 1)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
    b = (torch.randn(cout, generator=g) * 0.1).cuda()
    y = ops.conv2d_nhwc(x, w, b, relu=True, pool=bool(pool), force_bn=bn, force_r=r)
    ref = torch.relu(torch.nn.functional.conv2d(xc.float(), w.bfloat16().float(), b, padding=1))
    if pool:
        ref = torch.nn.functional.max_pool2d(ref, 2, 2)
    _check(y, ref.permute(0, 2, 3, 1).contiguous())
    return y


CONV_CASES = [
    # n, H, cin, cin_pad, cout, pool, bn, r        -- tile shape exercised
    (1, 16, 64, 64, 64, 0, 64, 1),                 # 8x16x1
    (2, 16, 64, 64, 64, 1, 64, 1),
    (2, 32, 128, 128, 128, 1, 128, 1),
    (2, 32, 128, 128, 256, 0, 256, 1),
    (3, 56, 128, 128, 256, 1, 0, 1),               # 8x8x2, odd batch -> partial tile in n
    (5, 28, 256, 256, 512, 1, 0, 1),               # 4x4x8
    (33, 14, 512, 512, 512, 1, 0, 1),              # 2x2x32
    (2, 32, 3, 16, 64, 0, 64, 1),                  # conv1_1 spatial: 16-channel k-blocks, 32B swizzle
    (2, 32, 20, 32, 64, 0, 64, 1),                 # conv1_1 temporal: 32-channel k-blocks, 64B swizzle
    (2, 32, 3, 16, 64, 0, 64, 3),                  # first-layer variants: vertical reuse only ...
    (2, 32, 20, 32, 64, 0, 64, 3),
    (2, 32, 3, 16, 64, 0, 64, 9),                  # ... and whole-filter stages (S=3)
    (3, 32, 20, 32, 64, 1, 64, 9),
    (2, 224, 3, 16, 64, 0, 0, 0),                  # full-size conv1_1, auto variant (S=3)
    (2, 224, 20, 32, 64, 0, 0, 0),
    (2, 32, 64, 64, 64, 1, 64, 1),                 # forced one-tap-per-stage on an R=3-capable shape
    (2, 32, 64, 64, 64, 1, 64, 3),                 # vertical tap reuse
    (2, 32, 128, 128, 128, 0, 128, 3),
    (2, 32, 128, 128, 256, 0, 256, 0),             # CTA-pair kernel (cta_group::2): auto variant for Cout >= 256
    (1, 56, 128, 128, 256, 1, 0, 0),               # ... odd number of M tiles (49): the pair's tail tile is out of bounds
    (3, 56, 256, 256, 256, 0, 0, 0),
    (5, 28, 256, 256, 512, 1, 0, 0),               # ... two N tiles
    (33, 14, 512, 512, 512, 1, 0, 0),
    (70, 14, 512, 512, 512, 0, 0, 0),              # ... more work units than clusters
    (2, 224, 64, 64, 64, 1, 64, 0),                # full-size conv1_2, auto variant
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "n%d_H%d_c%d_%d_p%d_bn%d_r%d" % (c[0], c[1], c[2], c[4], c[5], c[6], c[7]))
def test_conv_vs_torch(case):
    _conv_case(*case)


def test_conv_no_relu_and_1x1():
    from video_analytics_b200 import ops
    g = torch.Generator().manual_seed(5)
    xc = torch.randn(2, 64, 16, 16, generator=g).cuda().bfloat16()
    x = xc.permute(0, 2, 3, 1).contiguous()
    w = (torch.randn(128, 64, 1, 1, generator=g) / 8).cuda()
    b = torch.randn(128, generator=g).cuda()
    y = ops.conv2d_nhwc(x, w, b, relu=False, pool=False)
    ref = torch.nn.functional.conv2d(xc.float(), w.bfloat16().float(), b)
    _check(y, ref.permute(0, 2, 3, 1).contiguous())
    assert float(y.float().min()) < 0.0          # ReLU really off


@pytest.mark.parametrize("n,fin,fout,bn,f32", [(128, 64, 64, 64, False), (200, 512, 256, 128, False), (200, 512, 256, 256, False),
                                              (70, 1024, 256, 64, True), (1, 4096, 256, 0, True), (250, 25088, 4096, 0, False)])
def test_linear_vs_torch(n, fin, fout, bn, f32):
    from video_analytics_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, fin, generator=g).cuda().bfloat16()
    w = (torch.randn(fout, fin, generator=g) / fin ** 0.5).cuda()
    b = (torch.randn(fout, generator=g) * 0.1).cuda()
    y = ops.linear(x, w, b, relu=True, out_f32=f32, force_bn=bn)
    ref = torch.relu(x.float() @ w.bfloat16().float().t() + b)
    if f32:
        assert y.dtype == torch.float32
        assert torch.allclose(y, ref, rtol=1e-4, atol=1e-4)
    else:
        _check(y, ref)


def test_conv_linearity_and_determinism():
    """Size-independent properties at full layer size: conv(2x) == 2 conv(x) exactly for power-of-two scaling
    (bias 0, no ReLU clipping issue since scaling is positive), and repeated launches are bit-identical."""
    from video_analytics_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4, 56, 56, 256, generator=g).cuda().bfloat16()
    w = (torch.randn(256, 256, 3, 3, generator=g) / 48).cuda()
    b = torch.zeros(256).cuda()
    y1 = ops.conv2d_nhwc(x, w, b, relu=True, pool=True)
    y2 = ops.conv2d_nhwc(x * 2, w, b, relu=True, pool=True)
    y3 = ops.conv2d_nhwc(x, w, b, relu=True, pool=True)
    assert torch.equal(y1, y3)
    assert torch.equal(y2.float(), y1.float() * 2)


def test_rejects_bad_shapes():
    from video_analytics_b200 import ops
    from video_analytics_b200._lib import VAError
    x = torch.zeros(1, 7, 7, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(VAError):
        ops.conv2d_nhwc(x, torch.zeros(64, 64, 3, 3).cuda(), torch.zeros(64).cuda())       # odd spatial size
    x = torch.zeros(1, 16, 16, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(VAError):
        ops.conv2d_nhwc(x, torch.zeros(48, 64, 3, 3).cuda(), torch.zeros(48).cuda())       # Cout not a multiple of 64

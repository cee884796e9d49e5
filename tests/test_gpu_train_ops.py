"""K5 primitives vs torch autograd (fp32 math on the same bf16-rounded operands): pool / ReLU / dropout backward,
bias gradients, the tcgen05 data-gradient and weight-gradient GEMMs, CE + logit layer, SGD-momentum.
The reference computes these with loss.backward() + optimizer.step() (Sheet03/spatialModel.py:178-181)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _nhwc(t):   # NCHW fp32 -> NHWC bf16
    return t.permute(0, 2, 3, 1).contiguous().bfloat16()


def _rel(ours, ref):
    return float((ours.float() - ref).norm() / ref.norm().clamp_min(1e-30))


def test_maxpool_fwd_bit_exact():
    from video_analytics_b200 import train_ops as T
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 64, 12, 20, generator=g).cuda().bfloat16()
    y = T.maxpool2x2(_nhwc(x.float()))
    ref = F.max_pool2d(x.float(), 2, 2)
    assert torch.equal(y.float(), ref.permute(0, 2, 3, 1))


@pytest.mark.parametrize("pooled", [True, False])
def test_relu_pool_bwd_bit_exact(pooled):
    """Routing (first maximum of a window in scan order, zero where the activation is zero) is index work: exact."""
    from video_analytics_b200 import train_ops as T
    g = torch.Generator().manual_seed(1)
    z = torch.randn(2, 32, 8, 12, generator=g).cuda().bfloat16().float()
    z[0, :, :2, :2] = 0.5            # ties inside windows: the first element must take the gradient
    z[1, :, 2:4, 2:4] = -1.0         # a window that ReLU zeroes entirely
    z.requires_grad_(True)
    y = torch.relu(z)
    out = F.max_pool2d(y, 2, 2) if pooled else y
    dout = torch.randn(out.shape, generator=g).cuda().bfloat16().float()
    out.backward(dout)
    dz = T.relu_pool_bwd(_nhwc(dout), _nhwc(y.detach()), pooled=pooled)
    assert torch.equal(dz.float(), z.grad.permute(0, 2, 3, 1))
    # fused bias gradient (8-channel path: C = 32 -> 4 channel groups) gives the same dz and the column sums
    db = torch.empty(32, dtype=torch.float32, device="cuda")
    dz2 = T.relu_pool_bwd(_nhwc(dout), _nhwc(y.detach()), pooled=pooled, bias_grad_out=db)
    assert torch.equal(dz2, dz)
    assert torch.allclose(db, z.grad.sum((0, 2, 3)), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("C", [64, 128, 512])
def test_pool_codes_route_like_the_activation(C):
    """maxpool codes + pool_bwd_codes == relu_pool_bwd on the activation itself (bit-exact), incl. ties and dead windows."""
    from video_analytics_b200 import train_ops as T
    g = torch.Generator().manual_seed(21)
    y = torch.relu(torch.randn(3, 12, 20, C, generator=g)).bfloat16()
    y[0, :4, :4] = 0.75                       # ties: the first element of each window must win
    y[1, 2:6, 2:8] = 0.0                      # windows with no positive element: no gradient
    y = y.cuda()
    p, codes = T.maxpool2x2(y, with_codes=True)
    assert torch.equal(p, T.maxpool2x2(y))
    dout = torch.randn(p.shape, generator=g).bfloat16().cuda()
    db_ref = torch.empty(C, dtype=torch.float32, device="cuda")
    db = torch.full((C,), 3.0, dtype=torch.float32, device="cuda")
    ref = T.relu_pool_bwd(dout, y, pooled=True, bias_grad_out=db_ref)
    got = T.pool_bwd_codes(dout, codes, bias_grad_out=db)
    assert torch.equal(got, ref)
    assert torch.allclose(db, db_ref, rtol=1e-5, atol=1e-5)
    assert torch.equal(T.pool_bwd_codes(dout, codes), ref)           # without the bias gradient


@pytest.mark.parametrize("shape", [(3, 14, 14, 512), (2, 28, 28, 192), (5, 1, 1, 4096)], ids=str)
@pytest.mark.parametrize("pooled", [True, False])
def test_relu_pool_bwd_fused_bias_shapes(shape, pooled):
    """Fused path (C/8 divides 256), the two-pass fallback (C = 192) and the FC shape (C = 4096)."""
    from video_analytics_b200 import train_ops as T
    n, H, W, C = shape
    if pooled and H == 1:
        pytest.skip("no pooling on FC activations")
    g = torch.Generator().manual_seed(12)
    y = torch.relu(torch.randn(n, H, W, C, generator=g)).cuda().bfloat16()
    dout = torch.randn((n, H // 2, W // 2, C) if pooled else (n, H, W, C), generator=g).cuda().bfloat16()
    dz = T.relu_pool_bwd(dout, y, pooled=pooled)
    db = torch.full((C,), 7.0, dtype=torch.float32, device="cuda")          # must be overwritten, not accumulated into
    dz2 = T.relu_pool_bwd(dout, y, pooled=pooled, bias_grad_out=db)
    assert torch.equal(dz, dz2)
    ref = dz.float().sum((0, 1, 2))
    assert torch.allclose(db, ref, rtol=1e-4, atol=1e-3 * float(ref.abs().max()))


def test_bias_grad():
    from video_analytics_b200 import train_ops as T
    g = torch.Generator().manual_seed(2)
    dz = torch.randn(5, 14, 14, 192, generator=g).cuda().bfloat16()
    db = T.bias_grad(dz)
    ref = dz.float().sum(dim=(0, 1, 2))
    assert torch.allclose(db, ref, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_dropout(dtype):
    from video_analytics_b200 import train_ops as T
    g = torch.Generator().manual_seed(3)
    x = torch.randn(7, 4096, generator=g).cuda().to(dtype)
    mask = (torch.rand(7, 4096, generator=g) < 0.5).to(torch.uint8).cuda()
    y = T.dropout(x, mask, 0.5)
    ref = (x.float() * mask.float() * 2.0).to(dtype)
    assert torch.equal(y, ref)


DGRAD_CASES = [
    # n, H, W, cin, cout
    (2, 16, 16, 64, 64),
    (2, 32, 32, 64, 128),
    (3, 28, 28, 256, 512),
    (5, 14, 14, 512, 512),
    (1, 56, 56, 128, 256),
]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_dgrad(case):
    from video_analytics_b200 import train_ops as T
    n, H, W, cin, cout = case
    g = torch.Generator().manual_seed(4)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
    dz = torch.randn(n, cout, H, W, generator=g).cuda().bfloat16().float()
    x = torch.zeros(n, cin, H, W, device="cuda", requires_grad=True)
    F.conv2d(x, w.bfloat16().float(), None, padding=1).backward(dz)
    dx = T.conv2d_dgrad(_nhwc(dz), w)
    ref = x.grad.permute(0, 2, 3, 1)
    assert _rel(dx, ref) < 4e-3, _rel(dx, ref)          # bf16 rounding of the output: 2^-9 rms
    assert float((dx.float() - ref).abs().max()) < 2.0 ** -7 * float(ref.abs().max()) + 1e-6


WGRAD_CASES = [
    # n, H, W, cin, cin_pad, cout
    (1, 4, 4, 64, 64, 64),             # tiny: single-row K splits, units whose every row is padding
    (2, 16, 16, 64, 64, 64),
    (2, 32, 32, 3, 16, 64),            # conv1_1 spatial: 3 real input channels inside a 16-channel NHWC tensor
    (2, 32, 32, 20, 32, 64),           # conv1_1 temporal
    (2, 224, 224, 3, 16, 64),          # full-size rows: 4 K chunks of 64 pixels, the last one half out of bounds
    (2, 112, 112, 64, 64, 128),
    (3, 56, 56, 128, 128, 256),
    (3, 28, 28, 256, 256, 512),        # 28-pixel rows padded to 32, CKP=32
    (5, 14, 14, 512, 512, 512),        # 14-pixel rows padded to 16, CKP=16
    (2, 12, 20, 64, 64, 192),          # H != W, Cout not a multiple of 128
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_wgrad(case):
    from video_analytics_b200 import train_ops as T
    n, H, W, cin, cin_pad, cout = case
    g = torch.Generator().manual_seed(5)
    xc = torch.randn(n, cin, H, W, generator=g).cuda().bfloat16().float()
    dz = torch.randn(n, cout, H, W, generator=g).cuda().bfloat16().float()
    x = torch.zeros(n, H, W, cin_pad, dtype=torch.bfloat16, device="cuda")
    x[..., :cin] = xc.permute(0, 2, 3, 1)
    w = torch.zeros(cout, cin, 3, 3, device="cuda", requires_grad=True)
    torch.backends.cudnn.allow_tf32 = False
    F.conv2d(xc, w, None, padding=1).backward(dz)
    dw = T.conv2d_wgrad(_nhwc(dz), x, cin)
    assert dw.shape == w.grad.shape
    # exact bf16 products, fp32 accumulation on both sides (order differs: split-K atomics)
    assert _rel(dw, w.grad) < 2e-5, _rel(dw, w.grad)


@pytest.mark.parametrize("case", [(8, 256, 128), (64, 4096, 256), (33, 512, 4096), (125, 1024, 25088)],
                         ids=lambda c: "x".join(map(str, c)))
def test_linear_wgrad_dgrad(case):
    from video_analytics_b200 import train_ops as T
    n, fin, fout = case
    g = torch.Generator().manual_seed(6)
    x = torch.randn(n, fin, generator=g).cuda().bfloat16()
    dy = torch.randn(n, fout, generator=g).cuda().bfloat16()
    w = (torch.randn(fout, fin, generator=g) / fin ** 0.5).cuda()
    dw = T.linear_wgrad(dy, x)
    ref_dw = dy.float().t() @ x.float()
    assert _rel(dw, ref_dw) < 2e-5, _rel(dw, ref_dw)
    dx = T.linear_dgrad(dy, w)
    ref_dx = dy.float() @ w.bfloat16().float()
    assert _rel(dx, ref_dx) < 4e-3, _rel(dx, ref_dx)


def test_transpose_roundtrip():
    from video_analytics_b200 import train_ops as T
    g = torch.Generator().manual_seed(7)
    x = torch.randn(5, 49, 512, generator=g).cuda().bfloat16()
    y = T.transpose_bf16(x)
    assert torch.equal(y, x.permute(0, 2, 1).contiguous())
    assert torch.equal(T.transpose_bf16(y), x)


def test_ce_train_vs_autograd():
    from video_analytics_b200 import train_ops as T
    g = torch.Generator().manual_seed(8)
    n, D, Cn = 37, 256, 101
    x = torch.randn(n, D, generator=g).cuda().requires_grad_(True)
    w4 = (torch.randn(Cn, D, generator=g) / D ** 0.5).cuda().requires_grad_(True)
    b4 = (torch.randn(Cn, generator=g) * 0.1).cuda().requires_grad_(True)
    labels = torch.randint(1, Cn, (n,), generator=g).cuda()        # the reference feeds its 1-based ids as they are
    torch.backends.cuda.matmul.allow_tf32 = False
    logits = F.linear(x, w4, b4)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    out = T.ce_train(x.detach(), w4.detach(), b4.detach(), labels)
    assert torch.allclose(out["logits"], logits.detach(), rtol=1e-5, atol=1e-5)
    assert abs(float(out["loss"]) - float(loss.detach())) < 1e-5 * max(1.0, abs(float(loss.detach())))
    assert torch.allclose(out["dx"], x.grad, rtol=1e-4, atol=1e-7)
    assert torch.allclose(out["dw4"], w4.grad, rtol=1e-4, atol=1e-7)
    assert torch.allclose(out["db4"], b4.grad, rtol=1e-4, atol=1e-7)


def test_sgd_momentum_matches_torch():
    from video_analytics_b200 import train_ops as T
    g = torch.Generator().manual_seed(9)
    p0 = torch.randn(10007, generator=g).cuda()
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([ref_p], lr=0.1, momentum=0.9)
    p = p0.clone()
    buf = torch.zeros_like(p)
    for step in range(3):
        grad = torch.randn(10007, generator=g).cuda()
        ref_p.grad = grad.clone()
        opt.step()
        T.sgd_momentum_(p, grad, buf, lr=0.1, momentum=0.9, first_step=(step == 0))
        assert torch.allclose(p, ref_p.detach(), rtol=1e-6, atol=1e-7), step


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_own_allreduce_peer_form_on_one_gpu(world):
    """va_allreduce_bf16, peer-pointer form: the `world` symmetric buffers are emulated by `world` buffers on this GPU (peer
    pointers are plain device pointers); each rank's launch reduces its slice in fp32 in rank order and writes it to every
    buffer.  After all ranks ran, every buffer holds round_bf16(sum of the original bf16 buffers) -- bit for bit."""
    import torch
    from video_analytics_b200 import train_ops as T
    g = torch.Generator().manual_seed(world)
    n = 8 * 1024 * 37 + 8 * 5                      # not a multiple of world * threads: ragged last slice
    bufs = [(torch.randn(n, generator=g) * (1 + r)).bfloat16().cuda() for r in range(world)]
    want = torch.zeros(n, dtype=torch.float32, device="cuda")
    for b in bufs:
        want += b.float()
    want = want.bfloat16()
    ptrs = [b.data_ptr() for b in bufs]
    for r in range(world):
        T.allreduce_bf16_(ptrs, 0, world, r, n, n_ctas=3)
    torch.cuda.synchronize()
    for b in bufs:
        assert torch.equal(b, want)


def test_own_allreduce_rejects_bad_arguments():
    import torch
    from video_analytics_b200 import train_ops as T
    from video_analytics_b200._lib import VAError
    x = torch.zeros(64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(VAError):
        T.allreduce_bf16_([x.data_ptr()], 0, 1, 0, 60)            # not a multiple of 8 elements
    with pytest.raises(VAError):
        T.allreduce_bf16_([x.data_ptr()] * 3, 0, 3, 0, 64)        # peer form: 1, 2, 4 or 8 ranks
    with pytest.raises(VAError):
        T.allreduce_bf16_([x.data_ptr(), 0], 0, 2, 0, 64)         # NULL peer

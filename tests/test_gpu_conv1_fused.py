"""Fused front end (csrc/va_conv1_fused.cu): the snippet transform gathered straight into conv1_1's tensor-core operand.

Parity chain: the oracle's transform (reference utils.py:137-151 semantics, bit-exact fp32) -> bf16 rounding -> fp32
`F.conv2d` + ReLU is the reference result; the fused kernel must match it like the unfused K1 + layer kernel does, for
protocol tables, for the reference's per-plane random crops/flips (temporalModel.py:86 quirk) and at the crop borders
(zero padding of the NORMALISED tensor, not of the image)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    from oracle import synth
    from video_analytics_b200.store import DeviceStore, make_layout
    lay = make_layout(3)
    store = DeviceStore(lay)
    rgb, flow = synth.build_store_numpy(lay)
    return lay, store, rgb, flow


def _weights(cin, seed):
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(64, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
    b = (torch.randn(64, generator=g) * 0.1).cuda()
    return w, b


def _reference(images, shape, table, mean, std, w, b):
    """K1's reference-layout fp32 output (bit-exact with the oracle, test_gpu_preprocess.py) -> bf16 -> fp32 conv."""
    from video_analytics_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = ops.preprocess(images, shape, table, mean, std, reference_layout=True)          # fp32 NCHW
    ref = torch.relu(torch.nn.functional.conv2d(x.bfloat16().float(), w.bfloat16().float(), b, padding=1))
    return ref.permute(0, 2, 3, 1).contiguous()


def _check(y, ref):
    diff = (y.float() - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 1e-2 * ref.abs().mean()
    assert not torch.isnan(y.float()).any()
    assert bool((diff <= tol).all()), (float(diff.max()), float(ref.abs().mean()), int((diff > tol).sum()))


def test_spatial_random_crops_and_borders(world):
    from video_analytics_b200 import ops
    lay, store, _, _ = world
    g = torch.Generator().manual_seed(5)
    n = 6
    ids = torch.randint(0, lay.n_rgb_images, (n, 1), generator=g)
    ci = torch.randint(0, 240 - 224 + 1, (n, 1), generator=g)
    cj = torch.randint(0, 320 - 224 + 1, (n, 1), generator=g)
    fl = torch.randint(0, 2, (n, 1), generator=g)
    ci[0], cj[0], fl[0] = 0, 0, 0            # crop at the image origin
    ci[1], cj[1], fl[1] = 16, 96, 1          # crop at the far corner, flipped
    ids[2] = lay.n_rgb_images - 1            # last image of the store
    table = torch.stack([ids, ci, cj, fl], dim=-1).to(torch.int32).cuda()
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    w, b = _weights(3, 1)
    y = ops.conv1_fused(store.rgb, lay.rgb_shape, table, mean, std, w, b)
    _check(y, _reference(store.rgb, lay.rgb_shape, table, mean, std, w, b))


def test_temporal_per_plane_crops(world):
    """20 independent (image, crop, flip) rows per stack: the reference's per-image transform quirk."""
    from video_analytics_b200 import ops
    lay, store, _, _ = world
    g = torch.Generator().manual_seed(6)
    n, planes = 4, 20
    ids = torch.randint(0, lay.n_flow_images, (n, planes), generator=g)
    ci = torch.randint(0, 256 - 224 + 1, (n, planes), generator=g)
    cj = torch.randint(0, 340 - 224 + 1, (n, planes), generator=g)
    fl = torch.randint(0, 2, (n, planes), generator=g)
    ci[0, :5], cj[0, :5] = 0, 0
    ci[1, :5], cj[1, :5], fl[1, :5] = 32, 116, 1
    table = torch.stack([ids, ci, cj, fl], dim=-1).to(torch.int32).cuda()
    mean, std = [0.485] * planes, [0.229] * planes
    w, b = _weights(planes, 2)
    y = ops.conv1_fused(store.flow, lay.flow_shape, table, mean, std, w, b)
    _check(y, _reference(store.flow, lay.flow_shape, table, mean, std, w, b))


@pytest.mark.parametrize("planes", [6, 16, 17])
def test_other_stack_depths_and_per_channel_normalisation(world, planes):
    """Even / odd chunk counts of the dense-K schedule (6 -> 1 chunk, 16 -> 2, 17 -> 3) and distinct (mean, std)."""
    from video_analytics_b200 import ops
    lay, store, _, _ = world
    g = torch.Generator().manual_seed(planes)
    n = 2
    ids = torch.randint(0, lay.n_flow_images, (n, planes), generator=g)
    ci = torch.randint(0, 33, (n, planes), generator=g)
    cj = torch.randint(0, 117, (n, planes), generator=g)
    fl = torch.randint(0, 2, (n, planes), generator=g)
    table = torch.stack([ids, ci, cj, fl], dim=-1).to(torch.int32).cuda()
    mean = [0.485 if p % 2 == 0 else 0.4 for p in range(planes)]
    std = [0.229 if p % 2 == 0 else 0.25 for p in range(planes)]
    w, b = _weights(planes, 3)
    y = ops.conv1_fused(store.flow, lay.flow_shape, table, mean, std, w, b)
    _check(y, _reference(store.flow, lay.flow_shape, table, mean, std, w, b))


def test_matches_unfused_layer_kernel(world):
    """Against K1 (bf16 NHWC) + the tcgen05 layer kernel: identical operands, only the fp32 summation order inside the
    tensor core differs -> at most one bf16 ulp apart, and mostly bit-equal."""
    from video_analytics_b200 import ops
    from video_analytics_b200.evaluate import spatial_table, temporal_table
    lay, store, _, _ = world
    m = lay.videos[1]
    sel = [0, 7, 131, 249]
    for (images, shape, tab, mean, std, c_pad) in (
            (store.rgb, lay.rgb_shape, spatial_table(m, lay.rgb_shape)[sel], [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], 16),
            (store.flow, lay.flow_shape, temporal_table(m, lay.flow_shape)[sel], [0.485] * 20, [0.229] * 20, 32)):
        table = torch.from_numpy(tab).cuda()
        cin = table.shape[1] * shape[2]
        w, b = _weights(cin, 4)
        y = ops.conv1_fused(images, shape, table, mean, std, w, b).float()
        x = ops.preprocess(images, shape, table, mean, std, c_pad=c_pad)
        y2 = ops.conv2d_nhwc(x, w, b, relu=True, pool=False).float()
        ulp = 2.0 ** -7 * y2.abs() + 1e-6
        assert bool(((y - y2).abs() <= ulp).all()), float((y - y2).abs().max())
        assert float((y != y2).float().mean()) < 0.02


def test_forward_store_equals_preprocess_plus_forward(world):
    """Whole stream: va_forward_store == va_preprocess + va_forward within the bf16 class-score budget."""
    from oracle import two_stream as ts
    from video_analytics_b200 import ops
    from video_analytics_b200.evaluate import spatial_table, temporal_table
    lay, store, _, _ = world
    m = lay.videos[0]
    sel = list(range(0, 250, 25)) + [249]
    for kind, cin, images, shape, tab, mean, std in (
            (0, 3, store.rgb, lay.rgb_shape, spatial_table(m, lay.rgb_shape)[sel], ts.NORM_MEANS_TF, ts.NORM_STDS_TF),
            (1, 20, store.flow, lay.flow_shape, temporal_table(m, lay.flow_shape)[sel], [0.485] * 20, [0.229] * 20)):
        model = ts.build_spatial_model(seed=0) if kind == 0 else ts.build_temporal_model(seed=0)
        net = ops.StreamNet(kind, cin, max_batch=8)
        net.load_state_dict(model.state_dict())
        table = torch.from_numpy(tab).cuda()
        d1, l1, p1, k1 = net.forward_store(images, shape, table, list(mean), list(std))
        x = ops.preprocess(images, shape, table, list(mean), list(std), c_pad=net.c_pad)
        d2, l2, p2, k2 = net.forward(x)
        assert float((p1 - p2).abs().max() / p2.abs().max()) < 1e-3
        assert float((d1 - d2).abs().max()) <= 2e-2 * float(d2.abs().max()) + 1e-6
        net.close()

"""K1 (fused snippet preprocess) and the synthetic store on the GPU vs the CPU oracle: bytes and fp32 pixels
bit-exact, bf16 output == round-to-nearest of the oracle's fp32."""
import hashlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def sha(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


@pytest.fixture(scope="module")
def world():
    from oracle import synth, two_stream as ts
    from video_analytics_b200.store import DeviceStore, make_layout
    lay = make_layout(3)
    store = DeviceStore(lay)
    rgb, flow = synth.build_store_numpy(lay)
    return lay, store, rgb, flow, ts.OracleStore(lay, rgb, flow)


def test_device_store_equals_oracle_store(world):
    lay, store, rgb, flow, _ = world
    assert np.array_equal(store.rgb.cpu().numpy()[:rgb.size].reshape(rgb.shape), rgb)
    assert np.array_equal(store.flow.cpu().numpy()[:flow.size].reshape(flow.shape), flow)


def test_golden_transform_cases_bit_exact(golden):
    """The kernel reproduces the sha256 the REFERENCE's getTransforms() output had (tests/golden/transform_cases.json)."""
    from oracle import synth
    from video_analytics_b200 import ops
    for c in golden("transform_cases.json"):
        shape = tuple(c["shape"])
        img = synth.synth_image(c["store_seed"], c["image_id"], shape)
        dev_img = torch.from_numpy(img).reshape(-1).cuda()
        table = torch.tensor([[[0, c["i"], c["j"], c["flip"]]]], dtype=torch.int32, device="cuda")
        mean, std = ([0.485], [0.229]) if c["flow"] else ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
        out = ops.preprocess(dev_img, shape, table, mean, std, reference_layout=True)
        assert out.shape == (1, shape[2], 224, 224)
        assert sha(out[0]) == c["sha256"], c


def test_protocol_snippets_bit_exact_and_bf16_rounding(world):
    from oracle import two_stream as ts
    from video_analytics_b200 import ops
    from video_analytics_b200.evaluate import spatial_table, temporal_table
    lay, store, _, _, ost = world
    m = lay.videos[1]
    sel = [0, 9, 10, 55, 128, 249]
    snips_s, _ = ts.video_snippets_spatial(ost, m.name)
    snips_t, _ = ts.video_snippets_temporal(ost, m.name)
    tab_s = torch.from_numpy(spatial_table(m, lay.rgb_shape)[sel]).cuda()
    tab_t = torch.from_numpy(temporal_table(m, lay.flow_shape)[sel]).cuda()
    ref_s = ops.preprocess(store.rgb, lay.rgb_shape, tab_s, ts.NORM_MEANS_TF, ts.NORM_STDS_TF, reference_layout=True)
    assert torch.equal(ref_s.cpu(), snips_s[sel])
    ref_t = ops.preprocess(store.flow, lay.flow_shape, tab_t, [0.485] * 20, [0.229] * 20, reference_layout=True)
    assert torch.equal(ref_t.cpu(), snips_t[sel])
    # network layout: bf16 NHWC, zero padded channels
    x_s = ops.preprocess(store.rgb, lay.rgb_shape, tab_s, ts.NORM_MEANS_TF, ts.NORM_STDS_TF, c_pad=16).cpu()
    assert torch.equal(x_s[..., :3], snips_s[sel].permute(0, 2, 3, 1).bfloat16())
    assert float(x_s[..., 3:].abs().max()) == 0.0
    x_t = ops.preprocess(store.flow, lay.flow_shape, tab_t, [0.485] * 20, [0.229] * 20, c_pad=32).cpu()
    assert torch.equal(x_t[..., :20], snips_t[sel].permute(0, 2, 3, 1).bfloat16())
    assert float(x_t[..., 20:].abs().max()) == 0.0


def test_per_image_crop_quirk_and_generic_shapes(world):
    """20 independent crops/flips per stack (reference temporalModel.py:86) and a non-standard plane count."""
    from oracle import two_stream as ts
    from video_analytics_b200 import ops
    lay, store, _, flow, _ = world
    g = torch.Generator().manual_seed(3)
    for planes in (20, 6):
        n = 3
        ids = torch.randint(0, lay.n_flow_images, (n, planes), generator=g)
        ci = torch.randint(0, 256 - 224 + 1, (n, planes), generator=g)
        cj = torch.randint(0, 340 - 224 + 1, (n, planes), generator=g)
        fl = torch.randint(0, 2, (n, planes), generator=g)
        table = torch.stack([ids, ci, cj, fl], dim=-1).to(torch.int32).cuda()
        out = ops.preprocess(store.flow, lay.flow_shape, table, [0.485] * planes, [0.229] * planes, reference_layout=True).cpu()
        for a in range(n):
            for p in range(planes):
                ref = ts.apply_transform(flow[int(ids[a, p])], int(ci[a, p]), int(cj[a, p]), int(fl[a, p]), [0.485], [0.229])
                assert torch.equal(out[a, p], ref[0])


def test_network_layout_random_per_plane_crops(world):
    """bf16 NHWC network input (the row-staged kernel: aligned word staging, per-plane byte shift and flip) for tables
    with 20 independent crops/flips per stack and every crop_j alignment, against the oracle transform, bit-exact; more
    snippets than one persistent block's share so the double-buffered staging wraps."""
    from oracle import two_stream as ts
    from video_analytics_b200 import ops
    lay, store, rgb, flow, _ = world
    g = torch.Generator().manual_seed(11)
    # flow stacks
    n, planes = 5, 20
    ids = torch.randint(0, lay.n_flow_images, (n, planes), generator=g)
    ci = torch.randint(0, 256 - 224 + 1, (n, planes), generator=g)
    cj = torch.randint(0, 340 - 224 + 1, (n, planes), generator=g)
    fl = torch.randint(0, 2, (n, planes), generator=g)
    ids[0, 0], ci[0, 0], cj[0, 0], fl[0, 0] = lay.n_flow_images - 1, 32, 116, 0     # last bytes of the store
    ids[0, 1], ci[0, 1], cj[0, 1], fl[0, 1] = lay.n_flow_images - 1, 32, 116, 1
    table = torch.stack([ids, ci, cj, fl], dim=-1).to(torch.int32).cuda()
    out = ops.preprocess(store.flow, lay.flow_shape, table, [0.485] * planes, [0.229] * planes, c_pad=32).cpu()
    assert out.shape == (n, 224, 224, 32) and float(out[..., 20:].abs().max()) == 0.0
    for a in range(n):
        for p in range(planes):
            ref = ts.apply_transform(flow[int(ids[a, p])], int(ci[a, p]), int(cj[a, p]), int(fl[a, p]), [0.485], [0.229])
            assert torch.equal(out[a, :, :, p], ref[0].bfloat16()), (a, p)
    # RGB frames, every column alignment (crop_j * 3 mod 4) and both flips
    n = 12
    ids = torch.randint(0, lay.n_rgb_images, (n, 1), generator=g)
    ci = torch.randint(0, 240 - 224 + 1, (n, 1), generator=g)
    cj = (torch.arange(n).reshape(n, 1) * 7 + 1) % (320 - 224 + 1)
    fl = (torch.arange(n).reshape(n, 1) // 4) % 2
    ids[0, 0], ci[0, 0], cj[0, 0] = lay.n_rgb_images - 1, 16, 96
    table = torch.stack([ids, ci, cj, fl], dim=-1).to(torch.int32).cuda()
    out = ops.preprocess(store.rgb, lay.rgb_shape, table, ts.NORM_MEANS_TF, ts.NORM_STDS_TF, c_pad=16).cpu()
    assert float(out[..., 3:].abs().max()) == 0.0
    for a in range(n):
        ref = ts.apply_transform(rgb[int(ids[a, 0])], int(ci[a, 0]), int(cj[a, 0]), int(fl[a, 0]), ts.NORM_MEANS_TF, ts.NORM_STDS_TF)
        assert torch.equal(out[a, :, :, :3], ref.permute(1, 2, 0).bfloat16()), a


def test_network_layout_many_snippets_matches_reference_layout(world):
    """A launch with more items than persistent blocks (every block loops and double-buffers): the bf16 NHWC result is
    the round-to-nearest of the (oracle-checked) fp32 reference-layout result for all 250 protocol snippets."""
    from video_analytics_b200 import ops
    from video_analytics_b200.evaluate import spatial_table, temporal_table
    from oracle import two_stream as ts
    lay, store, _, _, _ = world
    m = lay.videos[2]
    tab_s = torch.from_numpy(spatial_table(m, lay.rgb_shape)).cuda()
    tab_t = torch.from_numpy(temporal_table(m, lay.flow_shape)).cuda()
    for _ in range(2):      # second launch reuses the cached LUT and function attributes
        ref_s = ops.preprocess(store.rgb, lay.rgb_shape, tab_s, ts.NORM_MEANS_TF, ts.NORM_STDS_TF, reference_layout=True)
        x_s = ops.preprocess(store.rgb, lay.rgb_shape, tab_s, ts.NORM_MEANS_TF, ts.NORM_STDS_TF, c_pad=16)
        assert torch.equal(x_s[..., :3], ref_s.permute(0, 2, 3, 1).bfloat16())
        assert float(x_s[..., 3:].abs().max()) == 0.0
        del ref_s, x_s
        ref_t = ops.preprocess(store.flow, lay.flow_shape, tab_t, [0.485] * 20, [0.229] * 20, reference_layout=True)
        x_t = ops.preprocess(store.flow, lay.flow_shape, tab_t, [0.485] * 20, [0.229] * 20, c_pad=32)
        assert torch.equal(x_t[..., :20], ref_t.permute(0, 2, 3, 1).bfloat16())
        assert float(x_t[..., 20:].abs().max()) == 0.0
        del ref_t, x_t


def test_empty_and_invalid(world):
    from video_analytics_b200 import ops
    from video_analytics_b200._lib import VAError
    lay, store, _, _, _ = world
    empty = torch.zeros((0, 1, 4), dtype=torch.int32, device="cuda")
    out = ops.preprocess(store.rgb, lay.rgb_shape, empty, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], c_pad=16)
    assert out.shape == (0, 224, 224, 16)
    with pytest.raises(VAError):
        ops.preprocess(store.rgb, (100, 320, 3), torch.zeros((1, 1, 4), dtype=torch.int32, device="cuda"),
                       [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])          # crop larger than the image

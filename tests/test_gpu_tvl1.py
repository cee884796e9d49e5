"""GPU parity of the TV-L1 flow producer (va_tvl1_flow) against oracle/tvl1.py: BIT-exact u8 images,
fp32 flow and iteration counts (every fp32 operation of the kernel is the oracle's separately rounded IEEE operation)."""
import numpy as np
import pytest
import torch

from oracle import tvl1 as otv
from video_analytics_b200 import flow
from video_analytics_b200._lib import VAError

pytestmark = pytest.mark.gpu


def _oracle_pair(f0, f1, p):
    g0 = otv.gray_from_rgb(f0) if f0.shape[-1] == 3 else f0[..., 0]
    g1 = otv.gray_from_rgb(f1) if f1.shape[-1] == 3 else f1[..., 0]
    op = otv.TVL1Params(p.tau, p.lambda_, p.theta, p.nscales, p.warps, p.epsilon, p.iterations, p.scale_step, p.bound)
    u1, u2, st = otv.tvl1_flow(g0, g1, op, return_stats=True)
    return u1, u2, st, otv.flow_to_u8(u1, p.bound), otv.flow_to_u8(u2, p.bound)


def _run(clip, p, pairs=None):
    dev = torch.device("cuda")
    n, h, w, c = clip.shape
    frames = torch.from_numpy(clip).to(dev)
    if pairs is None:
        pairs = [(k, k + 1) for k in range(n - 1)]
    m = len(pairs)
    table = torch.tensor([[a, b, k, k + m] for k, (a, b) in enumerate(pairs)], dtype=torch.int32, device=dev)
    out = torch.full((2 * m, h, w), 7, dtype=torch.uint8, device=dev)
    res = flow.tvl1(frames, (h, w, c), table, out, params=p, return_flow=True, return_iterations=True)
    torch.cuda.synchronize()
    return out.cpu().numpy(), res["flow"].cpu().numpy(), res["iterations"].cpu().numpy()


@pytest.mark.parametrize("shape,channels,seed", [((48, 64), 3, 11), ((77, 100), 1, 12), ((33, 41), 3, 13), ((96, 128), 3, 14)])
def test_bit_exact_vs_oracle_small(shape, channels, seed):
    p = flow.TVL1Params()
    clip = flow.synthetic_clip(3, shape[0], shape[1], seed=seed, channels=channels, velocity=(1.2, -0.6), object_velocity=(-1.5, 0.8))
    out, fl, its = _run(clip, p)
    for k in range(2):
        u1, u2, st, qx, qy = _oracle_pair(clip[k], clip[k + 1], p)
        assert list(its[k][:len(st)]) == st, (k, list(its[k]), st)
        assert np.array_equal(fl[k, 0], u1) and np.array_equal(fl[k, 1], u2), float(np.abs(fl[k, 0] - u1).max())
        assert np.array_equal(out[k], qx) and np.array_equal(out[2 + k], qy)


def test_golden_fixture(golden):
    g = golden("tvl1_small.npz")
    p = flow.TVL1Params()
    out, fl, its = _run(g["clip_a"], p)
    for k in range(2):
        assert np.array_equal(fl[k, 0], g[f"a{k}_u1"]) and np.array_equal(fl[k, 1], g[f"a{k}_u2"])
        assert np.array_equal(out[k], g[f"a{k}_x"]) and np.array_equal(out[2 + k], g[f"a{k}_y"])
        assert list(its[k]) == list(g[f"a{k}_iters"])
    out, fl, its = _run(g["clip_b"], p)
    assert np.array_equal(fl[0, 0], g["b0_u1"]) and np.array_equal(out[0], g["b0_x"]) and np.array_equal(out[1], g["b0_y"])


def test_full_size_tsn_images_and_many_pairs():
    """340 x 256 (the TSN flow-image size, the band capacity limit) with more pairs than clusters; two pairs are checked
    against the oracle, all of them for determinism (same pair computed by different clusters)."""
    p = flow.TVL1Params()
    clip = flow.synthetic_clip(3, 256, 340, seed=21, channels=3)
    pairs = [(0, 1), (1, 2)] * 9 + [(0, 1)]
    out, fl, its = _run(clip, p, pairs)
    m = len(pairs)
    for k in (0, 1):
        u1, u2, st, qx, qy = _oracle_pair(clip[pairs[k][0]], clip[pairs[k][1]], p)
        assert list(its[k]) == st
        assert np.array_equal(fl[k, 0], u1) and np.array_equal(fl[k, 1], u2)
        assert np.array_equal(out[k], qx) and np.array_equal(out[m + k], qy)
    for k in range(2, m):
        assert np.array_equal(out[k], out[k % 2]) and np.array_equal(out[m + k], out[m + k % 2])
        assert np.array_equal(fl[k], fl[k % 2])


def test_frame_resize_to_the_tsn_flow_size():
    """UCF101 frames (240 x 320) resized to 340 x 256 first, as dense_flow does (cv::resize pinned to cv2 in the oracle's
    tests): u8 flow images and fp32 flow bit-equal to the oracle, written into the DEFAULT store layout's 256 x 340 images."""
    from video_analytics_b200.store import DeviceStore, make_layout
    p = flow.TVL1Params(new_size=(340, 256))
    clip = flow.synthetic_clip(3, 240, 320, seed=51)
    fr = torch.from_numpy(clip).cuda()
    fx, fy, fl = flow.flow_images(fr, params=p, return_flow=True)
    assert tuple(fx.shape) == (2, 256, 340)
    for k in range(2):
        ox, oy = otv.flow_images(clip[k], clip[k + 1], otv.TVL1Params(), new_size=(340, 256))
        assert np.array_equal(fx[k].cpu().numpy(), ox) and np.array_equal(fy[k].cpu().numpy(), oy)
    g0 = otv.gray_from_rgb(otv.resize_linear_u8(clip[0], 256, 340))
    g1 = otv.gray_from_rgb(otv.resize_linear_u8(clip[1], 256, 340))
    u1, u2 = otv.tvl1_flow(g0, g1)
    assert np.array_equal(fl[0, 0].cpu().numpy(), u1) and np.array_equal(fl[0, 1].cpu().numpy(), u2)
    lay = make_layout(1, min_frames=3, frame_span=1, flows_per_frame=1)            # default shapes: 240x320x3 frames, 256x340 flow
    store = DeviceStore(lay)
    m = lay.videos[0]
    cnt = flow.fill_flow_store(store, 0, fr, params=p)
    imgs = store.flow.view(-1, 256, 340)
    assert cnt == 2 and torch.equal(imgs[m.flowx_first:m.flowx_first + 2], fx) and torch.equal(imgs[m.flowy_first:m.flowy_first + 2], fy)


def test_video_to_flow_tree(tmp_path):
    """convertVideosToFlow: list line -> <root>/<Category>/<video>.avi -> flow_x_%04d.jpg / flow_y_%04d.jpg, the directory
    TemporalDataset walks (temporalModel.py:76-81); the pixels are the oracle's on the frames cv2.VideoCapture decodes."""
    cv2 = pytest.importorskip("cv2")
    import os
    clip = flow.synthetic_clip(5, 240, 320, seed=61)
    root, save = tmp_path / "videos", tmp_path / "flow"
    (root / "Archery").mkdir(parents=True)
    path = str(root / "Archery" / "v_Archery_g01_c01.avi")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (320, 240))
    if not vw.isOpened():
        pytest.skip("no video writer in this OpenCV build")
    for f in clip:
        vw.write(np.ascontiguousarray(f[..., ::-1]))
    vw.release()
    lst = tmp_path / "list.txt"
    lst.write_text("Archery/v_Archery_g01_c01.avi 1\n")
    assert flow.convertVideosToFlow(str(root), str(save), str(lst), mode="train") == 1
    out = save / "Archery" / "v_Archery_g01_c01"
    names = sorted(os.listdir(out))
    assert names == ["flow_x_%04d.jpg" % k for k in range(1, 5)] + ["flow_y_%04d.jpg" % k for k in range(1, 5)]
    frames = flow.read_video_frames(path)
    assert frames.shape == (5, 240, 320, 3)
    ox, oy = otv.flow_images(frames[0], frames[1], otv.TVL1Params(), new_size=(340, 256))
    want = cv2.imdecode(cv2.imencode(".jpg", ox, [cv2.IMWRITE_JPEG_QUALITY, 95])[1], cv2.IMREAD_GRAYSCALE)
    got = cv2.imread(str(out / "flow_x_0001.jpg"), cv2.IMREAD_GRAYSCALE)
    assert got.shape == (256, 340) and np.array_equal(got, want)


def test_store_straight_from_videos(tmp_path):
    """DeviceStore.from_videos: video files -> (every 10th frame, all TV-L1 flow pairs) in HBM, the counts and pixels the
    reference's two offline steps would have left on disk; the temporal protocol tables address it like any other store."""
    cv2 = pytest.importorskip("cv2")
    from video_analytics_b200.evaluate import temporal_table
    from video_analytics_b200.store import DeviceStore
    root = tmp_path / "videos"
    lines = []
    for v, (cat, nfr) in enumerate((("Archery", 23), ("Bowling", 31))):
        (root / cat).mkdir(parents=True)
        name = f"v_{cat}_g01_c0{v + 1}"
        vw = cv2.VideoWriter(str(root / cat / (name + ".avi")), cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (320, 240))
        if not vw.isOpened():
            pytest.skip("no video writer in this OpenCV build")
        for f in flow.synthetic_clip(nfr, 240, 320, seed=70 + v):
            vw.write(np.ascontiguousarray(f[..., ::-1]))
        vw.release()
        lines.append(f"{cat}/{name}.avi {v + 1}\n")
    store = DeviceStore.from_videos(lines, "train", str(root))
    lay = store.layout
    assert [(m.n_frames, m.n_flows, m.label) for m in lay.videos] == [(3, 22, 1), (4, 30, 2)]
    assert lay.rgb_shape == (240, 320, 3) and lay.flow_shape == (256, 340, 1)
    frames = flow.read_video_frames(str(root / "Bowling" / "v_Bowling_g01_c02.avi"))
    m = lay.videos[1]
    assert np.array_equal(store.rgb.view(-1, 240, 320, 3)[m.rgb_first + 2].cpu().numpy(), frames[20])       # frame "2.jpg" = video frame 20
    ox, oy = otv.flow_images(frames[7], frames[8], otv.TVL1Params(), new_size=(340, 256))
    fl = store.flow.view(-1, 256, 340)
    assert np.array_equal(fl[m.flowx_first + 7].cpu().numpy(), ox) and np.array_equal(fl[m.flowy_first + 7].cpu().numpy(), oy)
    tab = temporal_table(m, lay.flow_shape)                  # 250 snippets x 20 planes over this video's 30 flow pairs
    assert tab.shape == (250, 20, 4) and tab[:, :, 0].min() >= m.flowx_first and tab[:, :, 0].max() < m.flowy_first + m.n_flows


def test_more_pairs_than_scratch_slots():
    """One call with more pairs than a launch batch holds (63 scratch slots): the batches are processed in sequence and a
    pair's result does not depend on its batch or slot."""
    p = flow.TVL1Params()
    clip = flow.synthetic_clip(3, 40, 56, seed=81, channels=1)
    pairs = [(0, 1), (1, 2), (0, 2)] * 45                     # 135 pairs: three batches
    out, fl, its = _run(clip, p, pairs)
    m = len(pairs)
    u1, u2, st, qx, qy = _oracle_pair(clip[0], clip[2], p)
    assert np.array_equal(out[2], qx) and np.array_equal(out[m + 2], qy) and list(its[2][:len(st)]) == st
    for k in range(3, m):
        assert np.array_equal(out[k], out[k % 3]) and np.array_equal(out[m + k], out[m + k % 3]), k
        assert np.array_equal(fl[k], fl[k % 3]) and np.array_equal(its[k], its[k % 3]), k


def test_parameters_and_saturation():
    """Non-default parameters (fewer levels/warps, no early stop, small bound so that the 8-bit mapping saturates)."""
    p = flow.TVL1Params(tau=0.2, lambda_=0.1, theta=0.25, nscales=3, warps=2, epsilon=0.0, iterations=40, scale_step=0.7, bound=1.0)
    clip = flow.synthetic_clip(2, 60, 80, seed=31, channels=1, velocity=(2.5, -1.5))
    out, fl, its = _run(clip, p)
    u1, u2, st, qx, qy = _oracle_pair(clip[0], clip[1], p)
    assert st == [40] * len(st) and list(its[0][:len(st)]) == st
    assert np.array_equal(fl[0, 0], u1) and np.array_equal(fl[0, 1], u2)
    assert np.array_equal(out[0], qx) and np.array_equal(out[1], qy)
    assert (qx == 255).mean() > 0.5 and (qy == 0).mean() > 0.5


def test_identical_frames():
    clip = flow.synthetic_clip(1, 64, 64, seed=5, channels=3)
    out, fl, its = _run(np.concatenate([clip, clip]), flow.TVL1Params())
    assert float(np.abs(fl).max()) == 0.0 and np.all(out == 128) and np.all(its[0] == 2)


def test_flow_store_feeds_the_temporal_stream():
    """fill_flow_store writes the images where the temporal index tables read them (x images then y images of a video)."""
    from video_analytics_b200.store import DeviceStore, make_layout
    lay = make_layout(1, min_frames=3, frame_span=1, flows_per_frame=1, rgb_shape=(48, 64, 3), flow_shape=(48, 64, 1))
    store = DeviceStore(lay)
    m = lay.videos[0]
    clip = flow.synthetic_clip(m.n_flows + 1, 48, 64, seed=41)
    cnt = flow.fill_flow_store(store, 0, torch.from_numpy(clip).cuda())
    assert cnt == m.n_flows
    fx, fy = flow.flow_images(torch.from_numpy(clip).cuda())
    imgs = store.flow.view(-1, 48, 64)
    assert torch.equal(imgs[m.flowx_first:m.flowx_first + cnt], fx) and torch.equal(imgs[m.flowy_first:m.flowy_first + cnt], fy)
    _, _, _, qx, qy = _oracle_pair(clip[0], clip[1], flow.TVL1Params())
    assert np.array_equal(fx[0].cpu().numpy(), qx) and np.array_equal(fy[0].cpu().numpy(), qy)


def test_frames_beyond_the_on_chip_capacity():
    """288 x 400 frames: the finest pyramid level does not fit in the shared memory of a 16-CTA cluster and runs the
    global-memory fallback kernel, the four coarser ones the on-chip kernel -- same formulas, results bit-equal to the oracle."""
    p = flow.TVL1Params(iterations=60)                       # bounded oracle time at this size
    clip = flow.synthetic_clip(2, 288, 400, seed=91, channels=1)
    out, fl, its = _run(clip, p, [(0, 1), (1, 0), (0, 1)])
    u1, u2, st, qx, qy = _oracle_pair(clip[0], clip[1], p)
    assert list(its[0]) == st
    assert np.array_equal(fl[0, 0], u1) and np.array_equal(fl[0, 1], u2)
    assert np.array_equal(out[0], qx) and np.array_equal(out[3], qy)
    assert np.array_equal(out[2], out[0]) and np.array_equal(fl[2], fl[0])


def test_rejects_bad_arguments():
    dev = torch.device("cuda")
    frames = torch.zeros((2, 4200, 64, 1), dtype=torch.uint8, device=dev)
    with pytest.raises(VAError):
        flow.flow_images(frames)                             # beyond 4096 pixels
    with pytest.raises(VAError):
        flow.tvl1(torch.zeros(10, dtype=torch.uint8), (1, 1, 1), torch.zeros((1, 4), dtype=torch.int32), torch.zeros(10, dtype=torch.uint8))

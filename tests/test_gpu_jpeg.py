"""the CUDA baseline-JPEG decoder vs Pillow (the reference's own loader, spatialModel.py:76-79) and the
oracle restatement -- identical bytes, on the files cv2.imwrite writes (utils.py:116-120) and on the edge cases."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _synth(rng, h, w, c):
    yy, xx = np.mgrid[0:h, 0:w]
    chans = [(xx * 0.7 + yy * 0.3) % 256, (xx * 0.2 + yy * 0.9) % 256, 128 + 60 * np.sin(xx / 17) + 40 * np.cos(yy / 11)][:c]
    img = (np.stack(chans, -1) + rng.integers(-25, 25, (h, w, c))).clip(0, 255).astype(np.uint8)
    return img if c == 3 else img[..., 0]


def _files():
    import cv2
    from PIL import Image
    rng = np.random.default_rng(0)
    files = []
    for (h, w, c) in [(240, 320, 3), (256, 340, 1), (37, 53, 3), (100, 17, 1), (33, 2, 3), (1, 1, 3), (8, 8, 1), (16, 16, 3),
                      (17, 33, 3), (250, 7, 1)]:
        img = _synth(rng, h, w, c)
        files.append(cv2.imencode(".jpg", img)[1].tobytes())
        for q, ss in ((35, 0), (100, 2)):
            b = io.BytesIO()
            Image.fromarray(img).save(b, "JPEG", quality=q, subsampling=ss)
            files.append(b.getvalue())
    files.append(cv2.imencode(".jpg", _synth(rng, 64, 83, 3), [cv2.IMWRITE_JPEG_RST_INTERVAL, 3])[1].tobytes())
    files.append(cv2.imencode(".jpg", _synth(rng, 64, 83, 1), [cv2.IMWRITE_JPEG_RST_INTERVAL, 5, cv2.IMWRITE_JPEG_OPTIMIZE, 1])[1].tobytes())
    files.append(cv2.imencode(".jpg", rng.integers(0, 256, (48, 64, 3)).astype(np.uint8))[1].tobytes())      # noise: long codes
    files.append(cv2.imencode(".jpg", np.zeros((40, 40), np.uint8))[1].tobytes())                            # all-EOB blocks
    return files


def test_decode_bit_exact_vs_pillow_and_oracle():
    from PIL import Image
    from oracle import jpeg_baseline as J
    from video_analytics_b200 import jpeg
    files = _files()
    outs = jpeg.decode(files)
    torch.cuda.synchronize()
    for k, (f, o) in enumerate(zip(files, outs)):
        ref = np.asarray(Image.open(io.BytesIO(f)))
        got = o.cpu().numpy()
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        assert np.array_equal(got, ref), (k, ref.shape, int(np.abs(got.astype(int) - ref.astype(int)).max()))
    small = files[6]
    assert np.array_equal(jpeg.decode([small])[0].cpu().numpy(), J.decode(small))


def test_decode_store_frames_feed_preprocess():
    """Frames written the reference's way (cv2.imwrite), decoded into a store, cropped by K1: equals the reference
    pipeline Image.open -> crop -> ToTensor -> Normalize on the same files (bit-exact fp32)."""
    import cv2
    from PIL import Image
    from video_analytics_b200 import jpeg, ops
    rng = np.random.default_rng(5)
    h, w = 240, 320
    frames_bgr = [_synth(rng, h, w, 3) for _ in range(6)]
    files = [cv2.imencode(".jpg", f)[1].tobytes() for f in frames_bgr]
    store = torch.empty(len(files) * h * w * 3, dtype=torch.uint8, device="cuda")
    jpeg.decode_into(files, store, [k * h * w * 3 for k in range(len(files))])
    table = torch.tensor([[[k, 3 + k, 11 * k, k & 1]] for k in range(len(files))], dtype=torch.int32, device="cuda")
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    x = ops.preprocess(store, (h, w, 3), table, mean, std, reference_layout=True)
    torch.cuda.synchronize()
    for k, f in enumerate(files):
        img = np.asarray(Image.open(io.BytesIO(f)))                              # RGB, as the reference sees it
        crop = img[3 + k:3 + k + 224, 11 * k:11 * k + 224]
        if k & 1:
            crop = crop[:, ::-1]
        t = torch.from_numpy(crop.copy()).permute(2, 0, 1).float().div(255)
        t = (t - torch.tensor(mean).view(3, 1, 1)) / torch.tensor(std).view(3, 1, 1)
        assert torch.equal(x[k].cpu(), t)


def test_batch_of_flow_images_and_errors():
    import cv2
    from PIL import Image
    from video_analytics_b200 import jpeg
    from video_analytics_b200._lib import VAError
    rng = np.random.default_rng(9)
    imgs = [_synth(rng, 256, 340, 1) for _ in range(40)]
    files = [cv2.imencode(".jpg", im)[1].tobytes() for im in imgs]
    outs = jpeg.decode(files)
    for f, o in zip(files, outs):
        assert np.array_equal(o.cpu().numpy(), np.asarray(Image.open(io.BytesIO(f))))
    assert jpeg.decode([]) == []
    b = io.BytesIO()
    Image.fromarray(_synth(rng, 32, 32, 3)).save(b, "JPEG", progressive=True)
    with pytest.raises(jpeg.JpegFormatError):
        jpeg.decode([b.getvalue()])
    with pytest.raises(jpeg.JpegFormatError):
        jpeg.decode([b"not a jpeg at all"])
    with pytest.raises(VAError):
        jpeg.decode_into(files[:1], torch.empty(10, dtype=torch.uint8, device="cuda"), [0])     # does not fit
    with pytest.raises(VAError):
        jpeg.decode_into(files[:1], torch.empty(256 * 340, dtype=torch.uint8), [0])              # CPU tensor: no fallback


def test_many_files_with_their_own_huffman_tables():
    """Optimised-Huffman files carry private tables; decode_into splits the batch (<= 8 tables per kernel call)."""
    import cv2
    from PIL import Image
    from video_analytics_b200 import jpeg
    rng = np.random.default_rng(4)
    files = [cv2.imencode(".jpg", _synth(rng, 40 + 3 * k, 56 + k, 3 if k % 2 else 1), [cv2.IMWRITE_JPEG_OPTIMIZE, 1,
                                                                                      cv2.IMWRITE_JPEG_QUALITY, 50 + 4 * k])[1].tobytes()
             for k in range(9)]
    for f, o in zip(files, jpeg.decode(files)):
        assert np.array_equal(o.cpu().numpy(), np.asarray(Image.open(io.BytesIO(f))))


def test_store_from_jpeg_files_runs_the_protocol():
    """Frames and flow images written the reference's way -> DeviceStore.from_jpeg_files -> the 25x10 evaluation: the
    store holds Pillow's pixels, so the index tables / K1 / networks see what the reference's loader would feed."""
    import cv2
    from PIL import Image
    from oracle import synth
    from video_analytics_b200.store import DeviceStore, make_layout
    lay = make_layout(1)
    rgb, flow = synth.build_store_numpy(lay)
    rgb = rgb.reshape(-1, *lay.rgb_shape)
    flow = flow.reshape(-1, lay.flow_shape[0], lay.flow_shape[1])
    rgb_files = [cv2.imencode(".jpg", f[..., ::-1])[1].tobytes() for f in rgb]          # cv2 takes BGR
    flow_files = [cv2.imencode(".jpg", f)[1].tobytes() for f in flow]
    store = DeviceStore.from_jpeg_files(lay, rgb_files, flow_files)
    torch.cuda.synchronize()
    got_rgb = store.rgb.cpu().numpy().reshape(rgb.shape)
    got_flow = store.flow.cpu().numpy().reshape(flow.shape)
    for k in (0, len(rgb_files) // 2, len(rgb_files) - 1):
        assert np.array_equal(got_rgb[k], np.asarray(Image.open(io.BytesIO(rgb_files[k]))))
    for k in (0, len(flow_files) // 3, len(flow_files) - 1):
        assert np.array_equal(got_flow[k], np.asarray(Image.open(io.BytesIO(flow_files[k]))))
    with pytest.raises(ValueError):
        DeviceStore.from_jpeg_files(lay, flow_files[:1], [])                            # wrong size for the RGB store


@pytest.mark.parametrize("mode", ["serial", "parallel"])
def test_both_entropy_decoders_on_all_files(mode, monkeypatch):
    """The one-thread-per-image decoder and the block-per-image (self-synchronising chunks) decoder give Pillow's bytes."""
    from PIL import Image
    from video_analytics_b200 import jpeg
    monkeypatch.setenv("VA_JPEG_PARALLEL", "0" if mode == "serial" else "1")
    files = _files()
    for f, o in zip(files, jpeg.decode(files)):
        assert np.array_equal(o.cpu().numpy(), np.asarray(Image.open(io.BytesIO(f))))

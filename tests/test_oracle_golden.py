"""The CPU oracle replays the fixtures that oracle/make_golden.py recorded while checking it against the reference's
own code (utils.py / parameters.py imported, SpatialDataset and combineDescriptors exec'd from source)."""
import hashlib
import io
import random

import numpy as np
import pytest
import torch

from oracle import synth, two_stream as ts
from video_analytics_b200.store import make_layout


def sha(t):
    a = t.detach().cpu().contiguous().numpy() if isinstance(t, torch.Tensor) else np.ascontiguousarray(t)
    return hashlib.sha256(a.tobytes()).hexdigest()


def test_videoinfo_samples(golden):
    g = golden("videoinfo_samples.json")
    assert g["n_lines"] == 2409 + 951
    for mode in ("train", "test"):
        assert len(g["samples"][mode]) >= 10
        for s in g["samples"][mode]:
            assert list(ts.videoInfo(s["line"], mode)) == s["expect"]
    labels = [int(s["expect"][2]) for s in g["samples"]["train"]]
    assert min(labels) >= 1 and max(labels) <= 25          # 1-based mini-UCF101 labels


def test_transform_cases_bit_exact(golden):
    for c in golden("transform_cases.json"):
        img = synth.synth_image(c["store_seed"], c["image_id"], tuple(c["shape"]))
        torch.manual_seed(c["seed"])
        i, j, flip = ts.draw_transform_params(c["shape"][0], c["shape"][1])
        assert (i, j, flip) == (c["i"], c["j"], c["flip"])
        mean, std = (ts.flow_norm_constants(1) if c["flow"] else (ts.NORM_MEANS_TF, ts.NORM_STDS_TF))
        out = ts.apply_transform(img, i, j, flip, mean, std)
        assert sha(out) == c["sha256"]
        assert [float(x) for x in out.flatten()[:8]] == c["first8"]
    exact = [c for c in golden("transform_cases.json") if c["shape"][0] == 224]
    assert exact and all(c["i"] == 0 and c["j"] == 0 for c in exact)      # no crop draw for crop-sized images


def test_average_meter(golden):
    g = golden("average_meter.json")
    torch.manual_seed(g["seed"])
    m = ts.AverageMeter()
    for _ in range(g["n"]):
        m.update(torch.rand(g["dim"]))
    assert sha(m.avg) == g["sha256_avg"] and m.count == g["n"]


def test_spatial_dataset_draws(golden):
    """Frame index + crop/flip draws recorded from the REFERENCE class under fixed seeds (the draws depend only on
    the RNG streams and the folder sizes, so the synthetic store replays them)."""
    lay = make_layout(3)
    store = ts.OracleStore(lay, np.zeros((lay.n_rgb_images,) + tuple(lay.rgb_shape), np.uint8),
                           np.zeros((0,) + tuple(lay.flow_shape), np.uint8))
    lines = [lay.list_line(v, "train") for v in range(3)]
    ds = ts.SpatialDataset(lines, store, actionLabelDict={m.category: m.label for m in lay.videos})
    for rec in golden("spatial_dataset_draws.json"):
        random.seed(rec["seed"]); torch.manual_seed(rec["seed"])
        _, label, name = ds[rec["index"]]
        assert ds.last_indices["frame"] == rec["frame"]
        assert [list(c) for c in ds.last_indices["crops"]] == rec["crops"]
        assert label == lay.videos[rec["index"]].label and name == lay.videos[rec["index"]].name


def test_combine_descriptors(golden, tmp_path):
    g = golden("combine_descriptors.json")
    ps, pt = tmp_path / "s.csv", tmp_path / "t.csv"
    ps.write_text(g["csv_spatial"], newline="")
    pt.write_text(g["csv_temporal"], newline="")
    X, y = ts.combineDescriptors(str(ps), str(pt))
    assert list(X.shape) == g["X_shape"] and [int(v) for v in y] == g["labels"]
    assert sha(X) == g["sha256_X"]
    assert X.shape[1] == 512 and len(y) == 5               # video 7 exists only in the temporal file -> inner join drops it


def test_svm_decision_matches_sklearn():
    from sklearn import svm
    rng = np.random.RandomState(0)
    X = rng.rand(60, 512)
    y = rng.randint(1, 6, size=60)
    clf = svm.LinearSVC().fit(X, y)
    scores, idx = ts.svm_decision(X, clf.coef_, clf.intercept_)
    assert np.array_equal(clf.classes_[idx], clf.predict(X))
    assert np.allclose(scores, clf.decision_function(X), rtol=0, atol=1e-12)


def test_protocol_indices():
    for n in (12, 13, 25, 30, 181):
        f = ts.test_frame_indices(n)
        assert len(f) == 25 and f[0] == 0 and f[-1] == n - 1 and f == sorted(f)
    for nfl in (34, 35, 70):
        s = ts.test_flow_starts(nfl)
        assert len(s) == 25 and s[0] == 1 and s[-1] == nfl - 10 and s == sorted(s)
    crops = ts.ten_crop_params(240, 320)
    assert crops[:5] == [(0, 0, 0), (0, 96, 0), (16, 0, 0), (16, 96, 0), (8, 48, 0)]
    assert crops[5:] == [(0, 96, 1), (0, 0, 1), (16, 96, 1), (16, 0, 1), (8, 48, 1)]


def test_ten_crop_params_equal_torchvision():
    """Our (top, left, flip) table reproduces torchvision.transforms.functional.ten_crop pixel for pixel."""
    import torchvision.transforms.functional as F
    for shape in ((240, 320, 3), (256, 340, 1), (225, 231, 1)):
        img = synth.synth_image(1, 3, shape)
        t = torch.from_numpy(img).permute(2, 0, 1)
        ref = F.ten_crop(t, [224, 224])
        for (i, j, flip), r in zip(ts.ten_crop_params(shape[0], shape[1]), ref):
            patch = img[i:i + 224, j:j + 224]
            if flip:
                patch = patch[:, ::-1]
            assert np.array_equal(np.ascontiguousarray(patch).transpose(2, 0, 1), r.numpy())


def test_oracle_forward_vectors(golden):
    """The restated model reproduces its committed forward vectors (spatial only here: one VGG16 CPU forward of 4
    snippets; the temporal vectors are replayed by the GPU suite)."""
    g = golden("oracle_forward_spatial.npz")
    lay = make_layout(2)
    m = lay.videos[0]
    sel = [int(s) for s in g["sel"]]
    frames = ts.test_frame_indices(m.n_frames)
    crops = ts.ten_crop_params(*lay.rgb_shape[:2])
    xs = []
    for k, s in enumerate(sel):
        f, (i, j, fl) = frames[s // 10], crops[s % 10]
        assert [f, i, j, fl] == [int(v) for v in g["recs"][k]]
        x = ts.apply_transform(synth.synth_image(lay.seed, m.rgb_first + f, lay.rgb_shape), i, j, fl, ts.NORM_MEANS_TF, ts.NORM_STDS_TF)
        assert sha(x) == str(g["input_sha"][k])
        xs.append(x)
    fv, logits, pred = ts.forward_eval(ts.build_spatial_model(seed=0), torch.stack(xs))
    assert np.allclose(fv.numpy(), g["desc"], rtol=1e-4, atol=1e-6)
    assert np.allclose(logits.numpy(), g["logits"], rtol=1e-4, atol=1e-6)
    assert fv.shape == (4, 256) and logits.shape == (4, 101) and float(fv.min()) >= 0.0    # post-ReLU descriptors


def test_temporal_conv1_init_property():
    """Appendix A.10: at init every input channel of temporal conv1 holds the same (RGB-mean) kernel, so
    conv1(x) == conv(mean_kernel, sum_c x_c) + bias."""
    model = ts.build_temporal_model(seed=0)
    w = model.features[0].weight.data
    assert w.shape == (64, 20, 3, 3)
    assert all(torch.equal(w[:, 0], w[:, c]) for c in range(20))
    torch.manual_seed(1)
    x = torch.randn(1, 20, 16, 16)
    a = model.features[0](x)
    b = torch.nn.functional.conv2d(x.sum(1, keepdim=True), w[:, :1], model.features[0].bias, padding=1)
    assert torch.allclose(a, b, atol=1e-4)


def test_train_step_vs_golden(golden):
    """N5: the oracle's train_step reproduces the losses/weights recorded when make_golden.py ran it next to the
    reference's own loop-body lines (bit-equal there); across machines CPU conv kernels may differ in the last bits."""
    gold = golden("train_step.json")
    model = ts.build_spatial_model(seed=gold["model_seed"])
    opt = torch.optim.SGD(model.parameters(), gold["lr"], momentum=gold["momentum"])
    g = torch.Generator().manual_seed(gold["input_seed"])
    for it, st in enumerate(gold["steps"]):
        ip = torch.randn(gold["batch"], 3, 224, 224, generator=g)
        labels = torch.randint(1, 101, (gold["batch"],), generator=g)
        torch.manual_seed(gold["mask_seed_base"] + it)
        masks = ts.draw_dropout_masks([(gold["batch"], 4096), (gold["batch"], 4096), (gold["batch"], 256)])
        loss, fv, op = ts.train_step(model, opt, torch.nn.CrossEntropyLoss(), ip, labels, masks)
        assert abs(float(loss) - st["loss"]) < 1e-4 * abs(st["loss"]), (it, float(loss), st["loss"])
        assert abs(float(fv.abs().mean()) - st["fv_abs_mean"]) < 1e-3 * st["fv_abs_mean"]
        assert abs(float(model.features[0].weight.abs().mean()) - st["w0_abs_mean"]) < 1e-4 * st["w0_abs_mean"]
        assert abs(float(model.classifier[9].weight.abs().mean()) - st["w_last_abs_mean"]) < 1e-4 * st["w_last_abs_mean"]


def test_jpeg_restatement_vs_pillow():
    """(f)#2 oracle pin: oracle/jpeg_baseline.py restates Pillow's libjpeg-turbo decode path (what the reference's
    Image.open does to the files cv2.imwrite wrote, spatialModel.py:76-79, utils.py:116-120) -- identical bytes."""
    import io
    import cv2
    from PIL import Image
    from oracle import jpeg_baseline as J
    rng = np.random.default_rng(0)

    def synth(h, w, c):
        yy, xx = np.mgrid[0:h, 0:w]
        chans = [(xx * 0.7 + yy * 0.3) % 256, (xx * 0.2 + yy * 0.9) % 256, 128 + 60 * np.sin(xx / 17) + 40 * np.cos(yy / 11)][:c]
        img = (np.stack(chans, -1) + rng.integers(-25, 25, (h, w, c))).clip(0, 255).astype(np.uint8)
        return img if c == 3 else img[..., 0]

    files = []
    for (h, w, c) in [(240, 320, 3), (256, 340, 1), (37, 53, 3), (100, 17, 1), (33, 2, 3), (1, 1, 3)]:
        img = synth(h, w, c)
        files.append(cv2.imencode(".jpg", img)[1].tobytes())                      # the reference's writer
        b = io.BytesIO()
        Image.fromarray(img).save(b, "JPEG", quality=60, subsampling=0)           # 4:4:4 (colour) / other tables
        files.append(b.getvalue())
    files.append(cv2.imencode(".jpg", synth(64, 83, 3), [cv2.IMWRITE_JPEG_RST_INTERVAL, 3])[1].tobytes())
    files.append(cv2.imencode(".jpg", synth(64, 83, 1), [cv2.IMWRITE_JPEG_RST_INTERVAL, 5, cv2.IMWRITE_JPEG_OPTIMIZE, 1])[1].tobytes())
    for f in files:
        ref = np.asarray(Image.open(io.BytesIO(f)))
        mine = J.decode(f)
        assert mine.shape == ref.shape and np.array_equal(mine, ref)
    b = io.BytesIO()
    Image.fromarray(synth(32, 32, 3)).save(b, "JPEG", progressive=True)
    with pytest.raises(J.JpegUnsupported):
        J.decode(b.getvalue())


def test_model_surgery_vs_golden(golden):
    """State-dict hashes recorded while the reference's own __copyFirstLayer__ / __swapClassifier__ methods
    (temporalModel.py:149-181) were exec'd next to the restated builders (oracle/make_golden.py section (i))."""
    g = golden("model_surgery.json")
    for seed, h in g["seeds"].items():
        mt = ts.build_temporal_model(seed=int(seed)).state_dict()
        assert sha(mt["features.0.weight"]) == h["temporal_conv1_w_sha256"]
        assert sha(mt["features.0.bias"]) == h["temporal_conv1_b_sha256"]
        assert sha(mt["classifier.9.weight"]) == h["temporal_fc4_w_sha256"]
        ms = ts.build_spatial_model(seed=int(seed)).state_dict()
        assert sha(ms["classifier.0.weight"]) == h["spatial_fc1_w_sha256"]


def test_validate_body_vs_golden(golden):
    """validate() loop body (spatialModel.py:212-228) as recorded from the reference's own lines: summed loss, correct
    count and per-video AverageMeter means."""
    g = golden("validate_body.json")
    model = ts.build_spatial_model(seed=g["model_seed"])
    gen = torch.Generator().manual_seed(g["input_seed"])
    names_all = [("v_A_g01_c01", "v_B_g01_c01"), ("v_B_g01_c01", "v_C_g01_c02")]
    d, loss, correct = {}, 0, 0
    for it in range(2):
        ip = torch.randn(2, 3, 224, 224, generator=gen)
        labels = torch.randint(1, 101, (2,), generator=gen)
        fv, op, pred = ts.forward_eval(model, ip)
        loss = loss + torch.nn.CrossEntropyLoss()(op, labels)
        correct += int(pred.eq(labels).sum())
        ts.update_video_dict(d, names_all[it], labels, fv)
    assert float(loss) == g["loss"] and correct == g["correct"]
    assert {k: sha(v[0].avg) for k, v in sorted(d.items())} == g["avg_sha256"]

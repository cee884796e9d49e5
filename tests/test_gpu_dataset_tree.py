"""The reference's own constructor calls -- `SpatialDataset(videoListLoc, rootDir, transforms, actionLabelLoc=...)` and
`TemporalDataset(...)` over a TREE OF JPEG FOLDERS (spatialModel.py:286-298, temporalModel.py:315-324) -- work without
the extra `store=` argument: the folders are walked like the reference's __getitem__ does (spatialModel.py:72-77,
temporalModel.py:76-81), every JPEG is decoded once on the GPU (va_jpeg_decode) into the device store, and items equal
`Image.open(...)` + the reference transform bit for bit (the oracle replays the same RNG draws on Pillow's pixels)."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    cv2 = pytest.importorskip("cv2")
    from PIL import Image
    from oracle import synth
    root = tmp_path_factory.mktemp("ucf_tree")
    frames_root, flow_root = root / "frames", root / "flow"
    videos = [("ApplyEyeMakeup", "v_ApplyEyeMakeup_g01_c01", 1, 5, 13), ("Archery", "v_Archery_g02_c03", 2, 3, 12),
              ("Basketball", "v_Basketball_g01_c02", 8, 4, 11)]
    rgb, flow = {}, {}
    for vi, (cat, name, label, n_frames, n_flows) in enumerate(videos):
        d = frames_root / cat / name
        d.mkdir(parents=True)
        for i in range(n_frames):
            img = synth.synth_image(77, 100 * vi + i, (240, 320, 3))
            # low-pass the hash noise a little so the JPEGs look like frames, then write the reference's way (utils.py:116-120)
            img = cv2.GaussianBlur(img, (5, 5), 0)
            cv2.imwrite(str(d / f"{i}.jpg"), img[..., ::-1])
            rgb[(name, i)] = np.asarray(Image.open(d / f"{i}.jpg").convert("RGB"))
        d = flow_root / cat / name
        d.mkdir(parents=True)
        for i in range(1, n_flows + 1):
            for k, prefix in enumerate(("flow_x_", "flow_y_")):
                img = cv2.GaussianBlur(synth.synth_image(78 + k, 100 * vi + i, (256, 340, 1))[..., 0], (5, 5), 0)
                cv2.imwrite(str(d / f"{prefix}{i:04d}.jpg"), img)
                flow[(name, prefix, i)] = np.asarray(Image.open(d / f"{prefix}{i:04d}.jpg"))
    train_list = root / "train.txt"
    train_list.write_text("".join(f"{cat}/{name}.avi {label}\n" for cat, name, label, _, _ in videos))
    test_list = root / "test.txt"
    test_list.write_text("".join(f"{cat}/{name}.avi\n" for cat, name, _, _, _ in videos))
    class_ind = root / "classInd.txt"
    class_ind.write_text("".join(f"{label} {cat}\n" for cat, _, label, _, _ in videos))
    return dict(frames_root=str(frames_root), flow_root=str(flow_root), train_list=str(train_list), test_list=str(test_list),
                class_ind=str(class_ind), videos=videos, rgb=rgb, flow=flow)


def _oracle_store(ds_store, tree):
    """OracleStore over Pillow's decode of the same files, in the device store's image order."""
    from oracle import two_stream as ts
    lay = ds_store.layout
    rgb = np.zeros((max(1, lay.n_rgb_images),) + tuple(lay.rgb_shape), np.uint8)
    flow = np.zeros((max(1, lay.n_flow_images),) + tuple(lay.flow_shape[:2]) + (1,), np.uint8)
    for m in lay.videos:
        for i in range(m.n_frames):
            rgb[m.rgb_first + i] = tree["rgb"][(m.name, i)]
        for i in range(1, m.n_flows + 1):
            flow[m.flowx_first + i - 1, ..., 0] = tree["flow"][(m.name, "flow_x_", i)]
            flow[m.flowy_first + i - 1, ..., 0] = tree["flow"][(m.name, "flow_y_", i)]
    return ts.OracleStore(lay, rgb, flow)


@pytest.mark.parametrize("mode", ["train", "test"])
def test_spatial_dataset_over_jpeg_tree(tree, mode):
    from oracle import two_stream as ts
    from video_analytics_b200.spatialModel import SpatialDataset
    from video_analytics_b200.utils import getTransforms
    lst = tree["train_list"] if mode == "train" else tree["test_list"]
    ds = SpatialDataset(lst, tree["frames_root"], getTransforms(), mode=mode, actionLabelLoc=tree["class_ind"])   # reference call
    assert len(ds) == 3 and ds.store.layout.rgb_shape == (240, 320, 3)
    assert [m.n_frames for m in ds.store.layout.videos] == [5, 3, 4]
    # the decoded store IS Pillow's decode of every file
    ost = _oracle_store(ds.store, tree)
    assert np.array_equal(ds.store.rgb.cpu().numpy().reshape(ost.rgb.shape), ost.rgb)
    ods = ts.SpatialDataset(open(lst).readlines(), ost, mode=mode, actionLabelDict=ds.actionLabelDict)
    for rep in range(3):
        for idx in range(3):
            random.seed(31 * rep + idx); torch.manual_seed(31 * rep + idx)
            a = ds[idx]
            random.seed(31 * rep + idx); torch.manual_seed(31 * rep + idx)
            b = ods[idx]
            assert torch.equal(a[0].cpu(), b[0]) and a[1] == b[1] and a[2] == b[2], (rep, idx)
    assert [ds[i][1] for i in range(3)] == [1, 2, 8]


def test_temporal_dataset_over_jpeg_tree(tree):
    from oracle import two_stream as ts
    from video_analytics_b200.temporalModel import TemporalDataset
    from video_analytics_b200.utils import getTransforms
    ds = TemporalDataset(tree["train_list"], tree["flow_root"], getTransforms(), actionLabelLoc=tree["class_ind"])  # reference call
    assert [m.n_flows for m in ds.store.layout.videos] == [13, 12, 11]
    ost = _oracle_store(ds.store, tree)
    assert np.array_equal(ds.store.flow.cpu().numpy().reshape(ost.flow.shape), ost.flow)
    ods = ts.TemporalDataset(open(tree["train_list"]).readlines(), ost, actionLabelDict=ds.actionLabelDict)
    for idx in range(3):
        random.seed(5 + idx); torch.manual_seed(5 + idx)
        a = ds[idx]
        random.seed(5 + idx); torch.manual_seed(5 + idx)
        b = ods[idx]
        assert a[0].shape == (20, 224, 224)
        assert torch.equal(a[0].cpu(), b[0]) and a[1] == b[1] and a[2] == b[2], idx


def test_reference_main_flow_runs_on_the_tree(tree, tmp_path):
    """The body of the reference's main() (spatialModel.py:286-298) with its own argument lists: datasets from the
    tree, loaders, network, one validation pass; then the per-video descriptors are written in the reference's CSV."""
    from video_analytics_b200 import parameters as P
    from video_analytics_b200.spatialModel import SpatialDataset, SpatialNetwork
    from video_analytics_b200.utils import getDataLoader, getTransforms, saveVideoDescriptors
    imageTransforms = getTransforms()
    trainDataset = SpatialDataset(tree["train_list"], tree["frames_root"], imageTransforms, frameSampleSize=2,
                                  actionLabelLoc=tree["class_ind"])
    trainDataLoader = getDataLoader(trainDataset, batchSize=2)
    testDataset = SpatialDataset(tree["test_list"], tree["frames_root"], imageTransforms, mode="test", actionLabelLoc=tree["class_ind"])
    testDataLoader = getDataLoader(testDataset, batchSize=2)
    net = SpatialNetwork(P.NACTION_CLASSES, 1, P.INITIAL_LR, P.MOMENTUM_VAL, P.VIDEO_DESCRIPTOR_DIM, trainDataLoader,
                         testDataLoader, P.MILESTONES_LR, str(tmp_path / "ckp"), gpu=True, maxBatch=4)
    precision, loss = net.validate()
    assert 0.0 <= precision <= 1.0 and float(loss) > 0
    out = tmp_path / "spatial_test.csv"
    saveVideoDescriptors(net.testDict, str(out), True)
    rows = out.read_text().strip().splitlines()
    assert len(rows) == 3 and all(len(r.split(",")) == 2 + P.VIDEO_DESCRIPTOR_DIM for r in rows)

"""Host-side logic of the product package (no GPU): list parsing, RNG draw order, index tables, CSV wire format,
sharding.  The oracle (restated reference) is the checker."""
import io
import random
import types

import numpy as np
import pytest
import torch

from oracle import two_stream as ts
from video_analytics_b200 import utils as U
from video_analytics_b200.distributed import shard_bounds
from video_analytics_b200.evaluate import spatial_table, temporal_table
from video_analytics_b200.store import make_layout


def test_videoinfo_matches_reference(golden):
    g = golden("videoinfo_samples.json")
    for mode in ("train", "test"):
        for s in g["samples"][mode]:
            assert list(U.videoInfo(s["line"], mode)) == s["expect"]
    with pytest.raises(ValueError):
        U.videoInfo("NoSlashHere.avi 3", "train")            # the reference raises on malformed lines too


def test_transform_draw_order_matches_oracle():
    tr = U.getTransforms()
    for seed, (h, w) in enumerate([(240, 320), (256, 340), (224, 224), (224, 300)]):
        torch.manual_seed(seed)
        a = [tr.draw(h, w) for _ in range(5)]
        after_a = torch.rand(1)
        torch.manual_seed(seed)
        b = [ts.draw_transform_params(h, w) for _ in range(5)]
        after_b = torch.rand(1)
        assert a == b and torch.equal(after_a, after_b)       # same draws AND same RNG state afterwards
    with pytest.raises(ValueError):
        tr.draw(200, 320)
    with pytest.raises(NotImplementedError):
        U.getTransforms(jitter=[0.1, 0, 0, 0])


def test_norm_constants():
    tr = U.getTransforms()
    assert tr.norm_constants(3, 3) == ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    m, s = tr.norm_constants(20, 1)
    assert m == [0.485] * 20 and s == [0.229] * 20           # 2018 zip semantics for 1-channel flow images


class _StubStore:
    def __init__(self, layout):
        self.layout = layout


def _datasets(tmp_path, lay, mode="train"):
    from video_analytics_b200.spatialModel import SpatialDataset
    from video_analytics_b200.temporalModel import TemporalDataset
    lst, cls = tmp_path / "list.txt", tmp_path / "classInd.txt"
    lst.write_text("".join(lay.list_line(v, mode) for v in range(len(lay.videos))))
    cls.write_text("".join(f"{m.label} {m.category}\n" for m in lay.videos))
    tr = U.getTransforms()
    sd = SpatialDataset(str(lst), None, tr, mode=mode, actionLabelLoc=str(cls), store=_StubStore(lay))
    td = TemporalDataset(str(lst), None, tr, mode=mode, actionLabelLoc=str(cls), store=_StubStore(lay))
    return sd, td, lst.read_text().splitlines(keepends=True), {m.category: m.label for m in lay.videos}


@pytest.mark.parametrize("mode", ["train", "test"])
def test_dataset_indices_bit_exact_vs_oracle(tmp_path, mode):
    """Snippet, crop and flip indices are bit-exact with the restated reference datasets under the same seeds."""
    lay = make_layout(5)
    sd, td, lines, labels = _datasets(tmp_path, lay, mode)
    ost = ts.OracleStore(lay, np.zeros((lay.n_rgb_images,) + tuple(lay.rgb_shape), np.uint8),
                         np.zeros((lay.n_flow_images,) + tuple(lay.flow_shape), np.uint8))
    osd = ts.SpatialDataset(lines, ost, mode=mode, actionLabelDict=labels)
    otd = ts.TemporalDataset(lines, ost, mode=mode, actionLabelDict=labels)
    for seed in range(6):
        for idx in range(5):
            random.seed(seed * 7 + idx); torch.manual_seed(seed * 7 + idx)
            rows, label, name = sd.sample_indices(idx)
            random.seed(seed * 7 + idx); torch.manual_seed(seed * 7 + idx)
            _, olabel, oname = osd[idx]
            assert (label, name) == (olabel, oname)
            assert sd.last_indices == osd.last_indices
            m = lay.videos[idx]
            assert rows.tolist() == [[m.rgb_first + osd.last_indices["frame"], *osd.last_indices["crops"][0]]]
            random.seed(seed * 7 + idx); torch.manual_seed(seed * 7 + idx)
            rows, label, name = td.sample_indices(idx)
            random.seed(seed * 7 + idx); torch.manual_seed(seed * 7 + idx)
            _, olabel, oname = otd[idx]
            assert (label, name) == (olabel, oname) and td.last_indices == otd.last_indices
            s = otd.last_indices["start"]
            assert 1 <= s <= m.n_flows - 10
            assert rows.shape == (20, 4)
            assert rows[0::2, 0].tolist() == [m.flowx_first + s - 1 + l for l in range(10)]     # x_t, then
            assert rows[1::2, 0].tolist() == [m.flowy_first + s - 1 + l for l in range(10)]     # y_t interleaved
            assert len(set(map(tuple, rows[:, 1:].tolist()))) > 1                                # per-image crops (quirk)


def test_dataset_errors(tmp_path):
    from video_analytics_b200.spatialModel import SpatialDataset
    lay = make_layout(1)
    lst = tmp_path / "l.txt"
    lst.write_text(lay.list_line(0))
    with pytest.raises(ValueError, match="Action label dictionary required!"):
        SpatialDataset(str(lst), None, U.getTransforms(), store=_StubStore(lay))


def test_protocol_tables_match_oracle_records():
    lay = make_layout(3)
    ost = ts.OracleStore(lay, np.zeros((lay.n_rgb_images, 1, 1, 3), np.uint8), np.zeros((lay.n_flow_images, 1, 1, 1), np.uint8))
    for m in lay.videos:
        st = spatial_table(m, lay.rgb_shape)
        assert st.shape == (250, 1, 4) and st.dtype == np.int32
        frames = ts.test_frame_indices(m.n_frames)
        crops = ts.ten_crop_params(*lay.rgb_shape[:2])
        expect = [[m.rgb_first + f, *c] for f in frames for c in crops]
        assert st.reshape(250, 4).tolist() == expect
        tt = temporal_table(m, lay.flow_shape)
        assert tt.shape == (250, 20, 4)
        starts = ts.test_flow_starts(m.n_flows)
        fcrops = ts.ten_crop_params(*lay.flow_shape[:2])
        k = 0
        for s in starts:
            for c in fcrops:
                assert tt[k, 0::2, 0].tolist() == [m.flowx_first + s - 1 + l for l in range(10)]
                assert tt[k, 1::2, 0].tolist() == [m.flowy_first + s - 1 + l for l in range(10)]
                assert all(r == list(c) for r in tt[k, :, 1:].tolist())
                k += 1
        assert tt[:, :, 0].max() < lay.n_flow_images and st[:, :, 0].max() < lay.n_rgb_images


def test_ten_crop_and_indices_equal_oracle():
    for (h, w) in ((240, 320), (256, 340), (225, 231)):
        assert U.ten_crop_params(h, w) == ts.ten_crop_params(h, w)
    for n in (12, 17, 30):
        assert U.test_frame_indices(n) == ts.test_frame_indices(n)
        assert U.test_flow_starts(2 * n + 10) == ts.test_flow_starts(2 * n + 10)


def test_average_meter_and_csv_wire_format(golden, tmp_path):
    g = golden("combine_descriptors.json")
    # rebuild the dict the golden CSV was written from (same seed, same draw order as make_golden.py)
    torch.manual_seed(g["seed"])
    d = {}
    for v in range(5):
        m = U.AverageMeter()
        for _ in range(3):
            m.update(torch.rand(256))
        d[f"v_Class{v:03d}_g01_c01"] = (m, torch.tensor(1 + v))
    out = tmp_path / "s.csv"
    U.saveVideoDescriptors(d, str(out))
    assert out.read_bytes().replace(b"\r\n", b"\n") == g["csv_spatial"].encode().replace(b"\r\n", b"\n")
    # and the product join equals the reference's on the golden files
    from video_analytics_b200.combinedModel import combineDescriptors
    pt = tmp_path / "t.csv"
    pt.write_text(g["csv_temporal"], newline="")
    X, y = combineDescriptors(str(out), str(pt))
    Xo, yo = ts.combineDescriptors(str(out), str(pt))
    assert np.array_equal(X, Xo) and np.array_equal(y, yo) and X.shape == (5, 512)


def test_save_performance_and_dirs(tmp_path):
    p = tmp_path / "perf.csv"
    U.savePerformance(0.5, 1.25, str(p))
    U.savePerformance(0.75, 0.5, str(p))
    assert p.read_text() == "0.5,1.25\n0.75,0.5\n"
    a, b = tmp_path / "x" / "y", tmp_path / "z"
    assert U.checkAndMakeDirectories(str(a), str(b)) == [False, False]
    assert U.checkAndMakeDirectories(str(a), str(b)) == [True, True]
    assert float(U.getOneHot(3, 25)[0, 2]) == 1.0


def test_shard_bounds():
    assert shard_bounds(3783, 0, 8) == (0, 473, 473)
    assert shard_bounds(3783, 7, 8) == (3311, 3783, 473)       # last rank: 472 videos + 1 padding row
    covered = []
    for V, R in ((3783, 8), (10, 4), (3, 8), (16, 2), (1, 1)):
        rows = []
        for r in range(R):
            lo, hi, per = shard_bounds(V, r, R)
            assert 0 <= lo <= hi <= V and hi - lo <= per
            rows += list(range(lo, hi))
        assert rows == list(range(V))


def test_layout_is_deterministic_and_valid():
    a, b = make_layout(8), make_layout(8)
    assert a.videos == b.videos
    for m in a.videos:
        assert m.n_frames >= 25 and m.n_flows == 10 * m.n_frames and 1 <= m.label <= 25
        assert m.flowy_first == m.flowx_first + m.n_flows
    line = a.list_line(3, "train")
    assert U.videoInfo(line, "train")[1] == a.videos[3].name


def test_fastdiv_magic_numbers():
    """The magic-number division used for tile decode (csrc/va_conv_tc.cuh::FastDiv), restated: exact for every
    dividend < 2^31 and every divisor the planner can produce."""
    def make(d):
        if d == 1:
            return 0, 0
        l = 0
        while (1 << l) < d:
            l += 1
        p = 31 + l
        return (((1 << p) + d - 1) // d) & 0xFFFFFFFF, p - 32

    import random
    rng = random.Random(0)
    for d in list(range(1, 300)) + [392, 784, 3063, 12544, 49000, 98000]:
        mul, shr = make(d)
        assert ((1 << (31 + (shr if d > 1 else 0) + (1 if d > 1 else 0))) + d - 1) // d <= 0xFFFFFFFF or d == 1
        for x in [0, 1, d - 1, d, d + 1, 2 * d - 1, (1 << 31) - 1] + [rng.randrange(0, 1 << 31) for _ in range(200)]:
            q = x if d == 1 else ((x * mul) >> 32) >> shr
            assert q == x // d, (d, x)


def test_jpeg_header_parser_and_huffman_tables_match_oracle():
    """Host side of the JPEG row: the product's marker walk / canonical-table derivation vs the oracle's restatement of
    libjpeg (jdmarker.c / jpeg_make_d_derived_tbl) on files from cv2 and Pillow."""
    import io
    import cv2
    import numpy as np
    from PIL import Image
    from oracle import jpeg_baseline as J
    from video_analytics_b200 import jpeg
    rng = np.random.default_rng(2)
    files = [cv2.imencode(".jpg", rng.integers(0, 256, (40, 56, 3)).astype(np.uint8))[1].tobytes(),
             cv2.imencode(".jpg", rng.integers(0, 256, (33, 21)).astype(np.uint8), [cv2.IMWRITE_JPEG_OPTIMIZE, 1,
                                                                                      cv2.IMWRITE_JPEG_RST_INTERVAL, 2])[1].tobytes()]
    b = io.BytesIO()
    Image.fromarray(rng.integers(0, 256, (24, 24, 3)).astype(np.uint8)).save(b, "JPEG", quality=30, subsampling=0)
    files.append(b.getvalue())
    for f in files:
        mine, ref = jpeg.parse_header(f), J.parse_header(f)
        assert (mine.height, mine.width, mine.n_comp) == (ref.height, ref.width, len(ref.comps))
        assert mine.scan_offset == ref.scan_offset and mine.restart_interval == ref.restart_interval
        for c, comp in enumerate(ref.comps):
            assert np.array_equal(np.frombuffer(mine.qtables[c], dtype=np.uint16), ref.qt[comp.tq])
            for cls, tid, raw in ((0, comp.td, mine.htables[c][0]), (1, comp.ta, mine.htables[c][1])):
                t = jpeg._derive_huffman(raw[:16], raw[16:])
                maxcode, valoffset = J.derive_table(*ref.huff[(cls, tid)])
                assert np.array_equal(t["maxcode"][1:17], maxcode[1:17])
                used = [l for l in range(1, 17) if maxcode[l] >= 0]
                assert all(int(t["valoffset"][l]) == int(valoffset[l]) for l in used)
                assert np.array_equal(t["huffval"][:len(ref.huff[(cls, tid)][1])], ref.huff[(cls, tid)][1])
    batch = jpeg.JpegBatch(files, [0, 10000, 20000])
    assert batch.images["sampling"].tolist() == [2, 0, 1] and batch.qtables.shape[1] == 64
    assert int(batch.images["scan_offset"][1]) == len(files[0]) + jpeg.parse_header(files[1]).scan_offset
    with pytest.raises(jpeg.JpegFormatError):
        jpeg.parse_header(b"\x89PNG....")


def test_training_flop_accounting_matches_survey():
    """bench_train's algorithmic FLOPs: forward = SURVEY 8d figures minus the fp32 logit layer (2*256*101, not a
    tensor-core call); backward = data gradient (all layers but conv1_1) + weight gradient."""
    import bench_train as B
    fs, bs = B.stream_flops(3)
    ft, bt = B.stream_flops(20)
    assert int(fs) + 2 * 256 * 101 == 30_934_485_504
    assert int(ft) + 2 * 256 * 101 == 31_917_132_288
    first_s, first_t = 2.0 * 224 * 224 * 64 * 9 * 3, 2.0 * 224 * 224 * 64 * 9 * 20
    assert bs == 2 * fs - first_s and bt == 2 * ft - first_t


def test_flow_starts_reject_videos_shorter_than_a_stack():
    """reference temporalModel.py:79 `random.randint(1, nFlows - L)` raises ValueError for nFlows <= L; the protocol
    helper does too instead of returning start 0 (an image id before the video's first flow image)."""
    import pytest
    from video_analytics_b200.utils import test_flow_starts as flow_starts
    assert flow_starts(11, 10) == [1] * 25
    assert flow_starts(35, 10)[0] == 1 and flow_starts(35, 10)[-1] == 25
    for n in (10, 5, 0):
        with pytest.raises(ValueError):
            flow_starts(n, 10)


def test_frame_extraction_equals_the_reference(golden, tmp_path):
    """SURVEY 8f row 2 (second half): convertVideosToFrames / extractEveryNthFrame write what the REFERENCE's own functions
    wrote for the committed video (oracle/make_golden_frames.py ran Sheet03/utils.py:51-69,95-121 itself): same file names,
    same JPEG bytes, same kept frames."""
    import hashlib
    import os
    import shutil
    from conftest import GOLDEN
    cv2 = pytest.importorskip("cv2")
    man = golden("frames_manifest.json")
    video = os.path.join(GOLDEN, man["video"])
    assert hashlib.sha256(open(video, "rb").read()).hexdigest() == man["video_sha256"]
    if cv2.__version__ != man["cv2_version"]:
        pytest.skip("JPEG bytes are pinned for OpenCV %s" % man["cv2_version"])
    root, save = tmp_path / "videos", tmp_path / "frames"
    (root / "Archery").mkdir(parents=True)
    shutil.copy(video, root / "Archery" / "v_Archery_g01_c01.avi")
    lst = tmp_path / "list.txt"
    lst.write_text("Archery/v_Archery_g01_c01.avi 1\n")
    U.convertVideosToFrames(str(root), str(save), str(lst), man["sample_rate"], "train")
    out = save / "Archery" / "v_Archery_g01_c01"
    got = {n: hashlib.sha256(open(out / n, "rb").read()).hexdigest() for n in sorted(os.listdir(out))}
    assert got == man["reference_files"]
    kept = U.extractEveryNthFrame(str(root / "Archery" / "v_Archery_g01_c01.avi"), 7)
    assert len(kept) == man["every_7th"]["count"]
    assert [hashlib.sha256(np.ascontiguousarray(k).tobytes()).hexdigest() for k in kept] == man["every_7th"]["sha256"]
    with pytest.raises(ValueError):
        U.extractEveryNthFrame(str(root / "missing.avi"), 10)


def test_index_rows_are_validated_before_upload(tmp_path):
    """va_preprocess addresses images + id * image_bytes unchecked: rows are checked on the host against the store."""
    lay = make_layout(3)
    sd, td, _, _ = _datasets(tmp_path, lay)
    import random
    random.seed(1)
    torch.manual_seed(1)
    rows, _, _ = sd.sample_indices(0)
    sd.check_rows(rows)
    trows, _, _ = td.sample_indices(1)
    td.check_rows(trows)
    bad = rows.copy(); bad[0, 0] = lay.n_rgb_images
    with pytest.raises(IndexError):
        sd.check_rows(bad)
    bad = rows.copy(); bad[0, 0] = -1
    with pytest.raises(IndexError):
        sd.check_rows(bad)
    bad = trows.copy(); bad[3, 1] = lay.flow_shape[0] - 223          # a 224-row window starting there leaves the image
    with pytest.raises(ValueError):
        td.check_rows(bad)
    bad = trows.copy(); bad[0, 2] = -1
    with pytest.raises(ValueError):
        td.check_rows(bad)
    bad = rows.copy(); bad[0, 3] = 2
    with pytest.raises(ValueError):
        sd.check_rows(bad)
    U.check_index_rows(np.zeros((0, 1, 4), np.int32), 0, lay.rgb_shape)        # empty tables pass


def test_combined_model_exports_a_scikit_learn_estimator(monkeypatch):
    """combinedModel.main dumps what the reference dumps (combinedModel.py:36): a LinearSVC that predicts like the model."""
    from sklearn import svm
    from video_analytics_b200.combinedModel import CombinedModel
    monkeypatch.setattr(torch, "from_numpy", lambda a, _f=torch.from_numpy: _Host(_f(a)))
    rng = np.random.default_rng(0)
    X = rng.standard_normal((40, 6))
    for n_cls, labels in ((3, np.array([4, 9, 17])), (2, np.array([5, 8]))):
        y = labels[rng.integers(0, n_cls, 40)]
        ref = svm.LinearSVC().fit(X, y)
        m = CombinedModel().set_svm(ref.coef_, ref.intercept_, ref.classes_)
        est = m.as_sklearn()
        assert isinstance(est, svm.LinearSVC)
        assert np.array_equal(est.coef_, ref.coef_) and np.array_equal(est.intercept_, ref.intercept_)
        assert np.array_equal(est.classes_, ref.classes_)
        assert np.array_equal(est.predict(X), ref.predict(X))
        import io
        import joblib
        buf = io.BytesIO()
        joblib.dump(est, buf)
        buf.seek(0)
        assert np.array_equal(joblib.load(buf).predict(X), ref.predict(X))


class _Host:
    """torch tensor stand-in whose .cuda() stays on the host (set_svm uploads its weights; no GPU in the CPU suite)."""

    def __init__(self, t):
        self.t = t

    def cuda(self):
        return self.t

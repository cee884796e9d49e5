"""The reference-facing Python API on the GPU: datasets, loader, SpatialNetwork/TemporalNetwork.validate(), the
running consensus dict, CSV dump -> combineDescriptors -> CombinedModel, checkpoint save/resume."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    from oracle import synth, two_stream as ts
    from video_analytics_b200 import utils as U
    from video_analytics_b200.store import DeviceStore, make_layout
    tmp = tmp_path_factory.mktemp("api")
    lay = make_layout(6)
    store = DeviceStore(lay)
    rgb, flow = synth.build_store_numpy(lay)
    ost = ts.OracleStore(lay, rgb, flow)
    lst, cls = tmp / "list.txt", tmp / "classInd.txt"
    lst.write_text("".join(lay.list_line(v, "train") for v in range(6)))
    cls.write_text("".join(f"{m.label} {m.category}\n" for m in lay.videos))
    return dict(tmp=tmp, lay=lay, store=store, ost=ost, lst=str(lst), cls=str(cls), lines=lst.read_text().splitlines(keepends=True),
                labels={m.category: m.label for m in lay.videos}, tr=U.getTransforms())


def test_getitem_bit_exact_vs_oracle_datasets(setup):
    from oracle import two_stream as ts
    from video_analytics_b200.spatialModel import SpatialDataset
    from video_analytics_b200.temporalModel import TemporalDataset
    s = setup
    sd = SpatialDataset(s["lst"], None, s["tr"], actionLabelLoc=s["cls"], store=s["store"])
    td = TemporalDataset(s["lst"], None, s["tr"], actionLabelLoc=s["cls"], store=s["store"])
    osd = ts.SpatialDataset(s["lines"], s["ost"], actionLabelDict=s["labels"])
    otd = ts.TemporalDataset(s["lines"], s["ost"], actionLabelDict=s["labels"])
    assert len(sd) == len(osd) == 6
    for idx in (0, 3, 5):
        for ours, theirs in ((sd, osd), (td, otd)):
            random.seed(idx); torch.manual_seed(idx)
            a = ours[idx]
            random.seed(idx); torch.manual_seed(idx)
            b = theirs[idx]
            assert a[0].dtype == torch.float32 and tuple(a[0].shape) == tuple(b[0].shape)
            assert torch.equal(a[0].cpu(), b[0]) and a[1] == b[1] and a[2] == b[2]


def test_loader_order_and_batches_vs_oracle_dataloader(setup):
    """Same sampler permutation, same per-item draws, same pixels as torch's DataLoader over the oracle dataset."""
    from torch.utils.data import DataLoader
    from oracle import two_stream as ts
    from video_analytics_b200 import utils as U
    from video_analytics_b200.spatialModel import SpatialDataset
    s = setup
    sd = SpatialDataset(s["lst"], None, s["tr"], actionLabelLoc=s["cls"], store=s["store"])

    class _O(torch.utils.data.Dataset):
        def __init__(self):
            self.d = ts.SpatialDataset(s["lines"], s["ost"], actionLabelDict=s["labels"])

        def __len__(self):
            return len(self.d)

        def __getitem__(self, i):
            return self.d[i]

    random.seed(42); torch.manual_seed(42)
    ours = [(b.nhwc.cpu(), l, n) for b, l, n in U.getDataLoader(sd, batchSize=4)]
    random.seed(42); torch.manual_seed(42)
    theirs = list(DataLoader(_O(), batch_size=4, shuffle=True, num_workers=0))
    assert len(ours) == len(theirs) == 2 and ours[1][0].shape[0] == 2                 # 6 = 4 + 2, partial last batch
    for (x, l, n), (ox, ol, on) in zip(ours, theirs):
        assert tuple(n) == tuple(on) and torch.equal(l, ol)
        assert torch.equal(x[..., :3], ox.permute(0, 2, 3, 1).bfloat16())


def test_network_validate_vs_oracle(setup):
    """SpatialNetwork.validate(): precision, loss, and the running per-video consensus vs the oracle's loop."""
    from oracle import two_stream as ts
    from torch.utils.data import DataLoader
    from video_analytics_b200 import utils as U
    from video_analytics_b200.spatialModel import SpatialDataset, SpatialNetwork, SpatialModel
    s = setup
    sd = SpatialDataset(s["lst"], None, s["tr"], actionLabelLoc=s["cls"], store=s["store"])
    loader = U.getDataLoader(sd, batchSize=4)
    torch.manual_seed(0)
    net = SpatialNetwork(101, 2, 0.1, 0.9, 256, loader, loader, [10, 20], str(s["tmp"] / "ckp"), gpu=True, maxBatch=4)
    assert SpatialModel is SpatialNetwork and net.classifierLen == 10 and len(net.model.state_dict()) == 34
    assert all(k.startswith("module.") for k in net.model.state_dict())             # DataParallel key prefix kept
    oracle_model = ts.build_spatial_model(seed=123)
    net.model.module.load_state_dict(oracle_model.state_dict())
    net.sync_weights()
    random.seed(7); torch.manual_seed(7)
    precision, loss = net.validate()
    # oracle loop (reference validate(), spatialModel.py:197-231)
    class _O(torch.utils.data.Dataset):
        d = ts.SpatialDataset(s["lines"], s["ost"], actionLabelDict=s["labels"])
        def __len__(self): return len(self.d)
        def __getitem__(self, i): return self.d[i]
    random.seed(7); torch.manual_seed(7)
    odict, ocorrect, oloss = {}, 0, 0
    crit = torch.nn.CrossEntropyLoss()
    max_err = 0.0
    for data, labels, names in DataLoader(_O(), batch_size=4, shuffle=True, num_workers=0):
        fv, op, pred = ts.forward_eval(oracle_model, data)
        oloss += crit(op, labels)
        ocorrect += int((pred == labels).sum())
        ts.update_video_dict(odict, names, labels, fv)
    assert abs(float(loss) - float(oloss)) / float(oloss) < 5e-3
    assert set(net.testDict.keys()) == set(odict.keys()) and len(net.testDict) == 6
    for name in odict:
        meter, label = net.testDict[name]
        assert int(label) == int(odict[name][1]) and meter.count == 1
        ref = odict[name][0].avg
        assert float((meter.avg.cpu() - ref).abs().max() / ref.abs().max()) < 2e-2
    assert 0.0 <= precision <= 1.0
    # second epoch accumulates (dicts never reset, reference Appendix A.5)
    net.validate()
    assert net.testDict[next(iter(odict))][0].count == 2
    # CSV dump + checkpoint round trip
    csv_path = s["tmp"] / "spatial_test.csv"
    U.saveVideoDescriptors(net.testDict, str(csv_path), True)
    rows = csv_path.read_text().strip().splitlines()
    assert len(rows) == 6 and all(len(r.split(",")) == 258 for r in rows)
    net.epoch = 0
    net.save()
    assert os.path.isfile(net.resumeLoc)
    ck = torch.load(net.resumeLoc, weights_only=False)
    assert set(ck) == {"epoch", "model", "highestPrecision", "optimizer"}
    assert net.resume() and net.startEpoch == 1


def test_temporal_network_and_fusion_pipeline(setup):
    """TemporalNetwork forward on a loader batch + CSV -> combineDescriptors -> CombinedModel.predict end to end."""
    from oracle import two_stream as ts
    from video_analytics_b200 import utils as U
    from video_analytics_b200.combinedModel import CombinedModel, combineDescriptors
    from video_analytics_b200.temporalModel import TemporalDataset, TemporalNetwork
    s = setup
    td = TemporalDataset(s["lst"], None, s["tr"], actionLabelLoc=s["cls"], store=s["store"])
    loader = U.getDataLoader(td, batchSize=3, shuffle=False)
    net = TemporalNetwork(101, 10, 1, 0.1, 0.9, 256, loader, loader, [10, 20], str(s["tmp"] / "ckp_t"), gpu=True, maxBatch=3)
    oracle_model = ts.build_temporal_model(seed=5)
    net.model.module.load_state_dict(oracle_model.state_dict())
    net.sync_weights()
    it = iter(loader)                      # creating the iterator draws the loader's base seed from the torch RNG
    random.seed(3); torch.manual_seed(3)
    batch, labels, names = next(it)        # items (and their RNG draws) are produced here
    fv, logits = net.forward(batch)
    random.seed(3); torch.manual_seed(3)
    otd = ts.TemporalDataset(s["lines"], s["ost"], actionLabelDict=s["labels"])
    ox = torch.stack([otd[i][0] for i in range(3)])
    ofv, ol, _ = ts.forward_eval(oracle_model, ox)
    assert float((fv.cpu() - ofv).abs().max() / ofv.abs().max()) < 2e-2
    p, op = torch.softmax(logits.cpu(), 1), torch.softmax(ol, 1)
    assert float(((p - op).abs() / op).max()) < 1e-3
    # fusion through the CSV wire format
    net.testDict.update_batch(names, labels, fv)
    ps, pt = s["tmp"] / "s.csv", s["tmp"] / "t.csv"
    U.saveVideoDescriptors(net.testDict, str(pt), True)
    U.saveVideoDescriptors(net.testDict, str(ps), True)
    X, y = combineDescriptors(str(ps), str(pt))
    assert X.shape == (3, 512) and [int(v) for v in y] == [int(l) for l in labels]
    rng = np.random.RandomState(0)
    cm = CombinedModel().set_svm(rng.randn(101, 512), rng.randn(101))
    scores, idx = ts.svm_decision(X, cm.coef_, cm.intercept_)
    assert np.array_equal(cm.predict(X), idx)


def test_gpu_required_errors(setup):
    from video_analytics_b200._lib import VAError
    from video_analytics_b200.spatialModel import SpatialNetwork
    with pytest.raises(VAError):
        SpatialNetwork(101, 1, 0.1, 0.9, 256, None, None, [10], str(setup["tmp"] / "x"), gpu=False)

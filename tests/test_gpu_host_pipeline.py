"""TwoStreamEvaluator.host_pipeline: evaluation fed from PINNED HOST images (double-buffered H2D of only the images the
25 x 10 protocol reads) gives bit-identical rows to evaluation on the resident device store."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_host_pipeline_equals_resident_store():
    from video_analytics_b200 import ops
    from video_analytics_b200.combinedModel import CombinedModel
    from video_analytics_b200.evaluate import HostStore, TwoStreamEvaluator
    from video_analytics_b200.spatialModel import build_spatial_torch_model
    from video_analytics_b200.store import DeviceStore, make_layout
    from video_analytics_b200.temporalModel import build_temporal_torch_model
    lay = make_layout(3)
    store = DeviceStore(lay)
    ns, nt = ops.StreamNet(ops.STREAM_SPATIAL, 3, max_batch=125), ops.StreamNet(ops.STREAM_TEMPORAL, 20, max_batch=125)
    ns.load_state_dict(build_spatial_torch_model(101, 256, seed=0).state_dict())
    nt.load_state_dict(build_temporal_torch_model(101, 10, 256, seed=0).state_dict())
    ev = TwoStreamEvaluator(ns, nt, store, CombinedModel())
    host = HostStore.from_device(store)
    groups = [[0, 1], [2], [1, 0], [2, 2]]
    got = []
    for res in ev.host_pipeline(host, groups):
        got.append({k: v.clone() for k, v in res.items()})
        assert ev.last_h2d_bytes > 0
    torch.cuda.synchronize()
    assert len(got) == len(groups)
    for g, r in zip(groups, got):
        want = ev.run_videos(g)
        for k in ("video_scores", "video_desc", "score_pred"):
            assert torch.equal(r[k], want[k]), (g, k)
    # the same groups with the results read back by the pipeline (pinned host buffers, next group's kernels queued before a
    # group's results are handed out): identical rows, in order, D2H bytes counted
    got_h = []
    for res in ev.host_pipeline(host, groups, to_host=("video_scores", "score_pred")):
        assert not res["video_scores"].is_cuda and res["video_scores"].is_pinned()
        got_h.append({k: v.clone() for k, v in res.items()})
    assert len(got_h) == len(groups)
    for g, r, h in zip(groups, got, got_h):
        assert h["video_scores"].shape[0] == len(g)
        assert torch.equal(h["video_scores"], r["video_scores"].cpu()) and torch.equal(h["score_pred"], r["score_pred"].cpu())
    assert ev.last_d2h_bytes == len(groups[-1]) * (101 * 4 + 4)
    # bytes per group: 25 frames + 500 flow images per video + the two index tables
    per_video = 25 * 240 * 320 * 3 + 500 * 256 * 340 + 250 * 4 * 4 + 250 * 20 * 4 * 4
    assert ev.last_h2d_bytes == 2 * per_video
    ns.close(); nt.close()

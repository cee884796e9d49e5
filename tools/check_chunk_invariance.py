"""The evaluation result must not depend on the internal chunk size (max_batch): run the same videos through
StreamNets with max_batch 125 and a large chunk and compare every output bit for bit."""
import sys

import torch

sys.path.insert(0, ".")
from video_analytics_b200 import ops  # noqa: E402
from video_analytics_b200.combinedModel import CombinedModel  # noqa: E402
from video_analytics_b200.evaluate import TwoStreamEvaluator  # noqa: E402
from video_analytics_b200.spatialModel import build_spatial_torch_model  # noqa: E402
from video_analytics_b200.store import DeviceStore, make_layout  # noqa: E402
from video_analytics_b200.temporalModel import build_temporal_torch_model  # noqa: E402

big = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 4
layout = make_layout(4)
store = DeviceStore(layout, torch.device("cuda"))
sd_s = build_spatial_torch_model(101, 256, seed=0).state_dict()
sd_t = build_temporal_torch_model(101, 10, 256, seed=0).state_dict()
res = {}
for mb in (125, big):
    s = ops.StreamNet(ops.STREAM_SPATIAL, 3, 101, 256, max_batch=mb)
    t = ops.StreamNet(ops.STREAM_TEMPORAL, 20, 101, 256, max_batch=mb)
    s.load_state_dict(sd_s)
    t.load_state_dict(sd_t)
    ev = TwoStreamEvaluator(s, t, store, CombinedModel())
    r = ev.run_videos(list(range(nv)))
    torch.cuda.synchronize()
    res[mb] = {k: v.clone() for k, v in r.items()}
    s.close(); t.close()
ok = True
for k in res[125]:
    same = torch.equal(res[125][k], res[big][k])
    ok &= same
    print(k, "bit-equal" if same else f"DIFFERS max {float((res[125][k].double() - res[big][k].double()).abs().max()):.3e}")
print("CHUNK_INVARIANT" if ok else "CHUNK_DEPENDENT", big, nv)

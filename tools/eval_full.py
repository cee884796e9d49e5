"""BASELINE configs[3] as ONE job: two-stream evaluation of 3783 synthetic videos (25 snippets x 10 crops per stream),
whole videos sharded across the ranks, all-gather of the per-video rows, then the late-fusion SVM fitted and applied
on the gathered 512-d descriptors -- evaluate -> gather -> fit -> predict without leaving the device.

    python tools/eval_full.py [--videos 3783] [--pool 8]                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/eval_full.py

Prints one JSON line (rank 0): device-timed seconds for the whole job (max over ranks), snippets/s, and the
size-independent checks that stand in for an oracle at this size (the CPU oracle needs ~22 s per video):
  * determinism: video v is pool video v % P, so all rows with the same v % P must be bit-identical, wherever in a
    chunk, step or rank they were computed;
  * every rank holds the same gathered buffers (checksum all-reduce MIN == MAX);
  * reported, not asserted: train accuracy of the SVM fitted on [V, 512] with label = v % P, and whether `predict`
    (va_fuse) agrees with argmax of coef.X + intercept computed by torch in fp64.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analytics_b200 import ops                                              # noqa: E402
from video_analytics_b200.combinedModel import CombinedModel                      # noqa: E402
from video_analytics_b200.distributed import gather_video_rows, init_from_env, shard_bounds, trim_rows   # noqa: E402
from video_analytics_b200.evaluate import SNIPPETS_PER_VIDEO, TwoStreamEvaluator  # noqa: E402
from video_analytics_b200.spatialModel import build_spatial_torch_model          # noqa: E402
from video_analytics_b200.store import DeviceStore, make_layout                   # noqa: E402
from video_analytics_b200.temporalModel import build_temporal_torch_model        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=3783)
    ap.add_argument("--pool", type=int, default=8)
    ap.add_argument("--videos-per-step", type=int, default=2)
    ap.add_argument("--max-batch", type=int, default=500)
    a = ap.parse_args()
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    C, D, V, P, vps = 101, 256, a.videos, a.pool, a.videos_per_step
    spatial = ops.StreamNet(ops.STREAM_SPATIAL, 3, C, D, max_batch=a.max_batch)
    temporal = ops.StreamNet(ops.STREAM_TEMPORAL, 20, C, D, max_batch=a.max_batch)
    spatial.load_state_dict(build_spatial_torch_model(C, D, seed=0).state_dict())
    temporal.load_state_dict(build_temporal_torch_model(C, 10, D, seed=0).state_dict())
    store = DeviceStore(make_layout(P), dev)
    ev = TwoStreamEvaluator(spatial, temporal, store, CombinedModel())
    lo, hi, per = shard_bounds(V, rank, world)
    out = ev.alloc_outputs(world * per, D, C, with_svm=False)
    ev.run_videos([lo, min(lo + 1, hi - 1)])            # warm-up (tensor maps, LUTs, allocator)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for first in range(lo, hi, vps):
        ev.run_videos(list(range(first, min(first + vps, hi))), out=out, out_row=rank * per + (first - lo))
    gather_video_rows(out, rank, world, per)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res = trim_rows(out, V)

    # ---- checks
    desc, scores, pred = res["video_desc"], res["video_scores"], res["score_pred"]
    vid = torch.arange(V, device=dev)
    same = True
    for k in range(min(P, V)):
        rows = (vid % P) == k
        same &= bool((desc[rows] == desc[k]).all()) and bool((scores[rows] == scores[k]).all()) and bool((pred[rows] == pred[k]).all())
    ranks_agree = True
    if world > 1:
        chk = torch.stack([desc.double().sum(), scores.double().sum(), pred.double().sum()])
        mn, mx = chk.clone(), chk.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        ranks_agree = bool((mn == mx).all())

    # ---- late fusion on the gathered descriptors: fit + predict on the device (every rank holds all rows; rank 0 fits)
    line = None
    if rank == 0:
        labels = (np.arange(V) % P) + 1                                            # 1-based like the reference's labels
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        model = CombinedModel().fit(desc.double(), labels)
        got = model.predict(desc)
        f1.record()
        torch.cuda.synchronize()
        want = model.classes_[(desc.double() @ torch.from_numpy(model.coef_).to(dev).T
                               + torch.from_numpy(model.intercept_).to(dev)).argmax(1).cpu().numpy()]
        snippets = V * SNIPPETS_PER_VIDEO
        line = {"job": "BASELINE configs[3]: two-stream evaluation of %d synthetic videos (pool of %d) on %d GPU(s)" % (V, P, world),
                "seconds": float(ms.item()) * 1e-3, "snippets_per_s": snippets / (float(ms.item()) * 1e-3),
                "videos_per_s": V / (float(ms.item()) * 1e-3), "n_gpus": world,
                "rows_with_equal_pool_id_bit_identical": bool(same), "ranks_hold_identical_rows": ranks_agree,
                "svm_fit_predict_ms": f0.elapsed_time(f1), "svm_epochs_max": int(model.n_iter_.max()),
                "svm_train_accuracy": float((got == labels).mean()), "svm_predict_equals_fp64_argmax": bool((got == want).all())}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    # the SVM figures are reported, not asserted: random-init networks give nearly identical descriptors for different
    # videos, so separability at C = 1 is a property of the data, not of the code
    if rank == 0 and not (line["rows_with_equal_pool_id_bit_identical"] and line["ranks_hold_identical_rows"]):
        sys.exit(1)


if __name__ == "__main__":
    main()

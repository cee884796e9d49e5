"""Extra legs of the ONE bench.py JSON line: the BASELINE.json configs that are not the headline workload.

  configs0_b1   spatial stream, one snippet (configs[0] on the GPU): microseconds per forward, weight GB/s        (N = 1)
  configs1_b64  temporal stream, one batch of 64 flow stacks (configs[1]): ms, snippets/s, TFLOP/s               (N = 1)
  parity        top-1 agreement + the reference scores' top1-top2 margin histogram: bf16 path vs the fp32 parity
                mode (itself within 1e-5 of the oracle) over 8 videos x 500 snippets, and both GPU modes vs the CPU
                oracle on a 10+10-snippet sample                                                                   (N = 1)
  strong_3783   configs[3]: the whole 3783-video evaluation job, videos sharded over the ranks, total seconds     (every N)
  train         configs[4]: two-stream training step, batch 256 per GPU and stream, gradient all-reduce           (every N)
  tvl1_flow     SURVEY 8f row 4: TV-L1 production of the flow_x_/flow_y_ images at 340 x 256, frame pairs per second  (N = 1)
Every function returns a plain dict on rank 0 (None elsewhere) and never prints.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPATIAL_FLOPS = 30_934_485_504
TEMPORAL_FLOPS = 31_917_132_288
WEIGHT_BYTES_BF16 = 2 * 135_335_333          # one stream's parameters as the bf16 operands the kernels read


def _event_ms(fn, reps, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def small_batch_legs(spatial, temporal, ev, store, layout):
    """configs[0] (B = 1, spatial) and configs[1] (B = 64, temporal) through the same handles as the headline run."""
    import torch
    from video_analytics_b200 import ops
    ts_, tt_ = ev.tables_for(0)
    x1 = ops.preprocess(store.rgb, layout.rgb_shape, ts_[:1].contiguous(), ev.mean_s, ev.std_s, c_pad=spatial.c_pad)
    ms1 = _event_ms(lambda: spatial.forward(x1, want_logits=False, want_pred=True), reps=50, warm=5)
    t64 = tt_[:64].contiguous()
    x64 = ops.preprocess(store.flow, layout.flow_shape, t64, ev.mean_t, ev.std_t, c_pad=temporal.c_pad)
    ms64 = _event_ms(lambda: temporal.forward(x64, want_logits=False, want_pred=True), reps=10)
    def with_k1():
        x = ops.preprocess(store.flow, layout.flow_shape, t64, ev.mean_t, ev.std_t, c_pad=temporal.c_pad)
        return temporal.forward(x, want_logits=False, want_pred=True)
    ms64_e2e = _event_ms(with_k1, reps=10)
    del x1, x64
    return ({"workload": "spatial stream, ONE 3x224x224 snippet, bf16 (BASELINE configs[0] on the GPU): 13 conv + 3 FC + head launches",
             "us_per_forward": 1e3 * ms1, "snippets_per_s": 1e3 / ms1,
             "bound": "weight bandwidth + launch latency: 271 MB of bf16 weights per forward, 17 dependent launches",
             "weight_GBps": WEIGHT_BYTES_BF16 / (ms1 * 1e-3) / 1e9, "tflops": SPATIAL_FLOPS / (ms1 * 1e-3) / 1e12},
            {"workload": "temporal stream, batch 64 of 20x224x224 flow stacks, bf16, 1 GPU (BASELINE configs[1])",
             "ms_per_batch": ms64, "snippets_per_s": 64e3 / ms64, "tflops": 64 * TEMPORAL_FLOPS / (ms64 * 1e-3) / 1e12,
             "ms_per_batch_with_preprocess": ms64_e2e})


def flow_leg(cpu_baseline=True, pairs=126, frames=9, h=256, w=340):
    """va_tvl1_flow on a synthetic moving-texture clip (TSN's 340 x 256 flow-image size): device-timed pairs/s, the
    data-dependent iteration counts, and how far the on-chip iteration is from the SMs' fp32 issue rate.  The CPU figure
    beside it is oracle/tvl1.py (numpy, one core) on ONE pair of the same clip."""
    import numpy as np
    import torch
    from video_analytics_b200 import flow, ops
    dev = torch.device("cuda", torch.cuda.current_device())
    clip = flow.synthetic_clip(frames, h, w, seed=7)
    fr = torch.from_numpy(clip).to(dev)
    base = torch.arange(pairs, dtype=torch.int32, device=dev) % (frames - 1)
    k = torch.arange(pairs, dtype=torch.int32, device=dev)
    table = torch.stack([base, base + 1, k, k + pairs], dim=1).contiguous()
    out = torch.empty((2 * pairs, h, w), dtype=torch.uint8, device=dev)
    p = flow.TVL1Params()
    its = flow.tvl1(fr, (h, w, 3), table, out, params=p, return_iterations=True)["iterations"].cpu().numpy()
    ms = _event_ms(lambda: flow.tvl1(fr, (h, w, 3), table, out, params=p), reps=3, warm=1)
    levels = p.levels(h, w)
    sizes = [(h, w)]
    for _ in range(1, levels):
        sizes.append((int(round(sizes[-1][0] * p.scale_step)), int(round(sizes[-1][1] * p.scale_step))))
    px = np.array([a * b for a, b in sizes[::-1]], np.float64)
    pix_iters = float((its.reshape(pairs, levels, p.warps).sum(2) * px[None, :]).sum())
    sm = ops.device_info()["sm_count"]
    clk_hz = 1.0e6 * torch.cuda.clock_rate()          # current SM clock (MHz) as the driver reports it
    instr_min = 105.0           # fp32 + load/store instructions one primal + dual update of a pixel needs at least
    peak = sm * 128 * clk_hz / instr_min
    leg = {"workload": f"TV-L1 optical flow (OpenCV CUDA defaults: 5 levels x 5 warps x <= 300 iterations, epsilon 0.01), {pairs} frame pairs of "
                       f"{w}x{h} RGB, one thread-block cluster per pair and pyramid level (4 / 8 / 16 CTAs), u8 flow_x/flow_y out (bound 20)",
           "pairs_per_s": pairs / ms * 1e3, "ms": ms, "inner_iterations_per_pair": float(its.sum(1).mean()),
           "pixel_iterations_per_s": pix_iters / ms * 1e3,
           "roofline": {"bound": "fp32 issue (solver state resident in shared memory, no HBM traffic in the iteration)",
                        "achieved": pix_iters / ms * 1e3, "peak": peak, "unit": "pixel-iterations/s", "frac": pix_iters / ms * 1e3 / peak,
                        "peak_definition": f"{sm} SMs x 128 lanes x SM clock / {instr_min:.0f} instructions per pixel-iteration",
                        "limiters": "the two finest levels need 16-CTA clusters, of which 7 fit (112 of 148 SMs); ~2500 clk of neighbour handshakes + barriers per iteration; ~200 issued instructions per pixel-iteration"},
           "parity": "bit-identical to oracle/tvl1.py (tests/test_gpu_tvl1.py); oracle unpinned: third-party tool absent"}
    if cpu_baseline:
        from oracle import tvl1 as otv
        t0 = time.time()
        otv.tvl1_flow(otv.gray_from_rgb(clip[0]), otv.gray_from_rgb(clip[1]))
        dt = time.time() - t0
        leg["cpu_baseline"] = {"value": 1.0 / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
                               "sample": "one 340x256 pair of the same clip through oracle/tvl1.py (numpy fp32)"}
    return leg


def margin_histogram(scores):
    srt = scores.sort(dim=1, descending=True).values
    m = (srt[:, 0] - srt[:, 1]).double()
    edges = [0.0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1.0]
    return {f"[{edges[i]:g},{edges[i + 1]:g})": int(((m >= edges[i]) & (m < edges[i + 1])).sum()) for i in range(len(edges) - 1)}


def parity_leg(spatial, temporal, ev, store, layout, sd_spatial, sd_temporal, n_videos=8, oracle_logits=None):
    """Prediction agreement, reported with the margins that make it meaningful (under
    reference-faithful random init top1-top2 score margins are 1e-5..1e-2, so some flips are inevitable in bf16)."""
    import numpy as np
    import torch
    from video_analytics_b200 import ops
    from video_analytics_b200.combinedModel import CombinedModel
    from video_analytics_b200.evaluate import SNIPPETS_PER_VIDEO, TwoStreamEvaluator
    C, D = spatial.n_classes, spatial.desc_dim
    s32 = ops.StreamNet(ops.STREAM_SPATIAL, 3, C, D, max_batch=25, precision="fp32")
    t32 = ops.StreamNet(ops.STREAM_TEMPORAL, 20, C, D, max_batch=25, precision="fp32")
    s32.load_state_dict(sd_spatial)
    t32.load_state_dict(sd_temporal)
    n_videos = min(n_videos, len(layout.videos))
    out = {"videos": n_videos, "snippets_per_stream": n_videos * SNIPPETS_PER_VIDEO,
           "reference": "fp32 parity mode of the same kernels (bf16x3 operands, within 1e-5 of the CPU oracle: tests/test_gpu_forward.py)"}
    per_video = {"bf16": [], "fp32": []}
    for name, n16, n32, images, shape, col, mean, std in (
            ("spatial", spatial, s32, store.rgb, layout.rgb_shape, 0, ev.mean_s, ev.std_s),
            ("temporal", temporal, t32, store.flow, layout.flow_shape, 1, ev.mean_t, ev.std_t)):
        agree = n = 0
        max_rel = 0.0
        p32_all, chunks16, chunks32 = [], [], []
        for v in range(n_videos):
            table = ev.tables_for(v)[col]
            x16 = ops.preprocess(images, shape, table, mean, std, c_pad=n16.c_pad)
            _, _, p16, k16 = n16.forward(x16, want_logits=False)
            del x16
            ref_in = ops.preprocess(images, shape, table, mean, std, reference_layout=True)       # fp32 NCHW, bit-exact
            p32_v, k32_v = [], []
            for b in range(0, ref_in.shape[0], 25):
                _, _, p, k = n32.forward(n32.pack_input(ref_in[b:b + 25]), want_logits=False)
                p32_v.append(p); k32_v.append(k)
            p32, k32 = torch.cat(p32_v), torch.cat(k32_v)
            del ref_in
            agree += int((k16 == k32).sum()); n += int(k16.numel())
            max_rel = max(max_rel, float(((p16 - p32).abs() / p32).max()))
            p32_all.append(p32.cpu())
            chunks16.append(p16.mean(0)); chunks32.append(p32.mean(0))
        per_video["bf16"].append(torch.stack(chunks16)); per_video["fp32"].append(torch.stack(chunks32))
        out[name] = {"snippet_top1_agreement": agree / n, "snippets": n, "max_rel_class_score_err": max_rel,
                     "reference_margin_histogram": margin_histogram(torch.cat(p32_all))}
    f16 = (per_video["bf16"][0] + per_video["bf16"][1]) / 2
    f32 = (per_video["fp32"][0] + per_video["fp32"][1]) / 2
    out["fused_video"] = {"top1_agreement": float((f16.argmax(1) == f32.argmax(1)).float().mean()), "videos": n_videos,
                          "max_rel_fused_score_err": float(((f16 - f32).abs() / f32).max()),
                          "reference_margin_histogram": margin_histogram(f32.cpu())}
    if oracle_logits is not None:
        # both GPU modes against the CPU oracle's per-snippet logits from bench.py's cpu_baseline pass (the first snippets
        # of pool video 0 in protocol order) -- the oracle itself is executed only there
        oracle = {}
        for sname, lg, n16, n32, images, shape, col, mean, std in (
                ("spatial", oracle_logits[0], spatial, s32, store.rgb, layout.rgb_shape, 0, ev.mean_s, ev.std_s),
                ("temporal", oracle_logits[1], temporal, t32, store.flow, layout.flow_shape, 1, ev.mean_t, ev.std_t)):
            k = int(lg.shape[0])
            p_ref = torch.softmax(lg, 1)
            pred = lg.argmax(1)
            table = ev.tables_for(0)[col][:k].contiguous()
            _, _, p16, k16 = n16.forward(ops.preprocess(images, shape, table, mean, std, c_pad=n16.c_pad), want_logits=False)
            ref_in = ops.preprocess(images, shape, table, mean, std, reference_layout=True)
            _, _, p32, k32 = n32.forward(n32.pack_input(ref_in), want_logits=False)
            oracle[sname] = {"snippets": k,
                             "bf16_top1_agreement": float((k16.cpu().long() == pred).float().mean()),
                             "bf16_max_rel_class_score_err": float(((p16.cpu() - p_ref).abs() / p_ref).max()),
                             "fp32_top1_agreement": float((k32.cpu().long() == pred).float().mean()),
                             "fp32_max_rel_class_score_err": float(((p32.cpu() - p_ref).abs() / p_ref).max()),
                             "oracle_margin_histogram": margin_histogram(p_ref)}
        out["vs_cpu_oracle"] = oracle
    s32.close(); t32.close()
    return out


def strong_leg(ev, out_alloc, rank, world, dev, n_videos=3783, vps=2, bounded_at_n1=473):
    """BASELINE configs[3] as ONE job: `n_videos` videos, contiguous shard per rank, both streams + fusion per video, one
    all-gather of the fused rows at the end; device-timed, max over ranks.  On one GPU the whole job is ~50 s, so N = 1
    runs one rank's share of the 8-GPU job (473 videos) and says so."""
    import torch
    import torch.distributed as dist
    from video_analytics_b200.distributed import gather_video_rows, shard_bounds
    from video_analytics_b200.evaluate import SNIPPETS_PER_VIDEO
    V = n_videos if world > 1 else min(n_videos, bounded_at_n1)
    lo, hi, per = shard_bounds(V, rank, world)
    out = out_alloc(world * per)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
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
    sec = float(ms.item()) * 1e-3
    # every rank must hold the same gathered rows
    same = True
    if world > 1:
        chk = torch.stack([out["video_scores"].double().sum(), out["score_pred"].double().sum()])
        mn, mx = chk.clone(), chk.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        same = bool((mn == mx).all())
    if rank != 0:
        return None
    res = {"workload": "BASELINE configs[3]: two-stream 25x10 evaluation of %d synthetic videos sharded over %d GPU(s), "
                       "NCCL all-gather of the per-video rows" % (V, world),
           "videos": V, "n_gpus": world, "seconds": sec, "videos_per_s": V / sec, "snippets_per_s": V * SNIPPETS_PER_VIDEO / sec,
           "scaling": "strong", "ranks_hold_identical_rows": same}
    if V != n_videos:
        res["note"] = ("N = 1 runs %d videos (one rank's share of the 8-GPU job); %d videos at this rate = %.1f s (extrapolated)"
                       % (V, n_videos, n_videos / (V / sec)))
        res["seconds_3783_extrapolated"] = n_videos / (V / sec)
    return res


def train_leg(args, steps=4, warmup=2):
    """BASELINE configs[4]: the two-stream training step (bench_train.py), batch 256 per GPU and stream, at this N."""
    import copy
    import bench_train
    a = copy.copy(args)
    a.steps, a.warmup, a.no_cpu_baseline, a.impl = steps, warmup, True, "ours"
    line = bench_train.main(a, embedded=True)
    if line is None:
        return None
    return {"workload": line["config"]["workload"], "batch_per_gpu_per_stream": line["config"]["batch_per_gpu_per_stream"],
            "n_gpus": line["n_gpus"], "value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"],
            "steps": line["steps"], "warmup": line["warmup"], "e2e": line["e2e"], "scaling": "weak",
            "parallelism": line["config"]["parallelism"], "gpu_launches": line["gpu_launches"],
            "roofline": {k: line["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "share_of_step")} if line.get("roofline") else None,
            "loss_first_step": line["config"]["loss_first_step"], "loss_last_step": line["config"]["loss_last_step"]}

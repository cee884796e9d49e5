#!/usr/bin/env python
"""Headline benchmark: two-stream snippets/sec of the Sheet03 evaluation hot path on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

Workload (BASELINE.json configs[2], "Combined two-stream late fusion, 25 snippets x 10 crops per video, 101
classes, synthetic UCF101-shaped clips"; configs[3] at N > 1): one STEP = `--videos-per-step` whole videos through
preprocess (both streams) -> VGG16 spatial + temporal forward -> per-video consensus + late fusion; one "two-stream
snippet" = one spatial forward + one temporal forward + its share of fusion.  Videos shard across ranks (weak
scaling: per-GPU work fixed) and the per-video scores are all-gathered over NCCL inside the timed region.

`value`   : device-timed (CUDA events, max over ranks) with the frame store already resident in HBM.
`e2e`     : same metric through the public Python API with HOST (pinned) frames: every step copies its videos'
            u8 frames and index tables host->device and reads the fused scores back, inside the timed region.
`roofline`: the tensor-core layer kernel (conv_tc_kernel), ALGORITHMIC FLOPs / its live event-timed duration,
            against MEASURED_PEAKS.json's sustained bf16 figure (the kernel is timed inside a long step).
`cpu_baseline` / `--impl reference`: the reference's CPU path = oracle/two_stream.py on stock PyTorch fp32 (the
            reference's scripts are Python-2 only and cannot run; kind "port"), bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPATIAL_FLOPS = 30_934_485_504     # per snippet, SURVEY.md 8d / BASELINE.md section 3
TEMPORAL_FLOPS = 31_917_132_288
METRIC = "two-stream snippets/sec"
UNIT = "snippets/s"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "hbm": float(d.get("hbm_gbs", 6650.0)),
                "src": "MEASURED_PEAKS.json bf16_tflops_sustained"}
    return {"tflops": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained, 6.65 TB/s)"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Samples taken before this call (warm-up) are excluded from the summary."""
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[getattr(self, "first", 0):]:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        # under-load samples: the upper half of the SM clock readings (idle gaps between steps excluded)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_pass(n_spatial, n_temporal, batch=10, threads=None):
    """The reference's CPU path for the same workload, bounded: preprocess + forward + consensus + fusion of the
    first n snippets of one synthetic video per stream, via the oracle port on stock PyTorch fp32."""
    import numpy as np
    import torch
    from oracle import synth, two_stream as ts
    from video_analytics_b200.store import make_layout
    torch.set_num_threads(threads or os.cpu_count())
    state = cpu_reference_pass.__dict__.setdefault("state", {})
    if not state:
        lay = make_layout(1)
        rgb, flow = synth.build_store_numpy(lay)
        state.update(store=ts.OracleStore(lay, rgb, flow), name=lay.videos[0].name, ms=ts.build_spatial_model(seed=0),
                     mt=ts.build_temporal_model(seed=0))
    st, name = state["store"], state["name"]
    t0 = time.perf_counter()
    # spatial: first n of the 250 protocol snippets
    img_recs = [(f, c) for f in ts.test_frame_indices(st.n_frame_files(name)) for c in ts.ten_crop_params(*synth.RGB_SHAPE[:2])]
    xs = torch.stack([ts.apply_transform(st.frame(name, f), i, j, fl, ts.NORM_MEANS_TF, ts.NORM_STDS_TF)
                      for (f, (i, j, fl)) in img_recs[:n_spatial]])
    ds, ss, _, lg_s = ts.video_consensus(state["ms"], xs, batch=batch)
    mean, std = ts.flow_norm_constants(1)
    stk = [(s, c) for s in ts.test_flow_starts(st.n_flow_files(name) // 2) for c in ts.ten_crop_params(*synth.FLOW_SHAPE[:2])]
    xt = []
    for (s, (i, j, fl)) in stk[:n_temporal]:
        planes = []
        for idx in range(s, s + ts.VIDEO_INPUT_FLOW_COUNT):
            planes.append(ts.apply_transform(st.flow_x(name, idx), i, j, fl, mean, std))
            planes.append(ts.apply_transform(st.flow_y(name, idx), i, j, fl, mean, std))
        xt.append(torch.cat(planes, 0))
    dt_, st_, _, lg_t = ts.video_consensus(state["mt"], torch.stack(xt), batch=batch)
    state["last_logits"] = (lg_s, lg_t)        # per-snippet logits of this pass: the parity leg's oracle sample
    fused = ts.fuse_scores(ss, st_)
    _ = torch.cat([ds, dt_]), int(fused.argmax())
    return time.perf_counter() - t0, torch.get_num_threads()


def jpeg_decode_measurement(store, layout, dev, n_rgb=50, n_flow=1000):
    """the images ONE evaluation step touches (2 videos: 50 frames + 1000 flow images), written the
    reference's way (cv2.imwrite -> baseline JPEG), decoded into the store by the CUDA decoder: H2D of the compressed
    files + Huffman / IDCT / colour kernels, timed with CUDA events; Pillow (the reference's loader) on the host cores
    beside it.  Synthetic frames are hash noise: ~67 KB per file, close to the worst case for entropy decoding."""
    import concurrent.futures as cf
    import io
    import numpy as np
    import torch
    try:
        import cv2
        from PIL import Image
    except Exception as e:                      # encoder / reference decoder missing: nothing to measure against
        return {"unavailable": repr(e)}
    from video_analytics_b200 import jpeg
    rgb = store.rgb[:n_rgb * int(np.prod(layout.rgb_shape))].cpu().numpy().reshape(n_rgb, *layout.rgb_shape)
    nf = layout.flow_shape[0] * layout.flow_shape[1]
    have = min(n_flow, store.flow.numel() // nf)
    flow = store.flow[:have * nf].cpu().numpy().reshape(have, layout.flow_shape[0], layout.flow_shape[1])
    files_rgb = [cv2.imencode(".jpg", f[..., ::-1])[1].tobytes() for f in rgb]
    files_flow = [cv2.imencode(".jpg", flow[k % have])[1].tobytes() for k in range(n_flow)]
    out_rgb = torch.empty(n_rgb * rgb[0].size, dtype=torch.uint8, device=dev)
    out_flow = torch.empty(n_flow * nf, dtype=torch.uint8, device=dev)
    sets = [(jpeg.JpegFileSet(files_rgb), out_rgb, [k * rgb[0].size for k in range(n_rgb)]),
            (jpeg.JpegFileSet(files_flow), out_flow, [k * nf for k in range(n_flow)])]

    def run():
        for fs_, out_, offs_ in sets:
            fs_.decode_into(out_, offs_)
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 4
    a.record()
    for _ in range(reps):
        run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    ok = bool(np.array_equal(out_flow[:nf].cpu().numpy().reshape(flow[0].shape), np.asarray(Image.open(io.BytesIO(files_flow[0])))) and
              np.array_equal(out_rgb[:rgb[0].size].cpu().numpy().reshape(rgb[0].shape), np.asarray(Image.open(io.BytesIO(files_rgb[0])))))
    threads = os.cpu_count() or 1
    allf = files_rgb + files_flow
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda f: np.asarray(Image.open(io.BytesIO(f))).shape, allf[:64]))
        t0 = time.perf_counter()
        list(ex.map(lambda f: np.asarray(Image.open(io.BytesIO(f))).shape, allf))
        cpu_s = time.perf_counter() - t0
    n = n_rgb + n_flow
    out_bytes = out_rgb.numel() + out_flow.numel()
    # the same measurement on photo-like content (low-frequency structure + mild noise): files ~3x smaller, nearly every
    # block ends with EOB -- the regime the block-per-image parallel entropy decoder is built for
    rng = np.random.default_rng(0)

    def photo_like(h, w, c, k):
        yy, xx = np.mgrid[0:h, 0:w]
        base = [128 + 70 * np.sin(xx / (23.0 + k % 7) + k) * np.cos(yy / (31.0 + k % 5)), 120 + 60 * np.cos((xx + yy) / (41.0 + k % 3)),
                110 + 50 * np.sin(yy / 19.0 + 0.3 * k)][:c]
        img = (np.stack(base, -1) + rng.normal(0, 3.0, (h, w, c))).clip(0, 255).astype(np.uint8)
        return img if c == 3 else img[..., 0]
    p_rgb = [cv2.imencode(".jpg", photo_like(layout.rgb_shape[0], layout.rgb_shape[1], 3, k))[1].tobytes() for k in range(16)]
    p_flow = [cv2.imencode(".jpg", photo_like(layout.flow_shape[0], layout.flow_shape[1], 1, k))[1].tobytes() for k in range(32)]
    pf_rgb, pf_flow = [p_rgb[k % 16] for k in range(n_rgb)], [p_flow[k % 32] for k in range(n_flow)]
    sets_p = [(jpeg.JpegFileSet(pf_rgb), out_rgb, sets[0][2]), (jpeg.JpegFileSet(pf_flow), out_flow, sets[1][2])]

    def run_p():
        for fs_, out_, offs_ in sets_p:
            fs_.decode_into(out_, offs_)
    run_p(); torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        run_p()
    b.record(); torch.cuda.synchronize()
    ms_p = a.elapsed_time(b) / reps
    ok_p = bool(np.array_equal(out_flow[:nf].cpu().numpy().reshape(flow[0].shape), np.asarray(Image.open(io.BytesIO(pf_flow[0])))) and
                np.array_equal(out_rgb[:rgb[0].size].cpu().numpy().reshape(rgb[0].shape), np.asarray(Image.open(io.BytesIO(pf_rgb[0])))))
    with cf.ThreadPoolExecutor(threads) as ex:
        t0 = time.perf_counter()
        list(ex.map(lambda f: np.asarray(Image.open(io.BytesIO(f))).shape, pf_rgb + pf_flow))
        cpu_p = time.perf_counter() - t0
    return {"images_per_call": n, "ms_per_call": ms, "images_per_s": n / (ms * 1e-3), "decoded_GB_per_s": out_bytes / (ms * 1e-3) / 1e9,
            "compressed_bytes_per_call": sum(len(f) for f in allf), "bit_exact_vs_pillow": ok,
            "content": "frames of the synthetic store: hash noise, every coefficient non-zero, no EOB -- the worst case for entropy "
                       "decoding and for chunk self-synchronisation",
            "bound": "latency of the entropy decode",
            "cpu_baseline": {"images_per_s": n / cpu_s, "cores": threads, "kind": "reference",
                             "sample": "PIL.Image.open + np.asarray of the same files on a %d-thread pool" % threads},
            "photo_like": {"images_per_call": n, "ms_per_call": ms_p, "images_per_s": n / (ms_p * 1e-3),
                           "decoded_GB_per_s": out_bytes / (ms_p * 1e-3) / 1e9,
                           "compressed_bytes_per_call": sum(len(f) for f in pf_rgb + pf_flow), "bit_exact_vs_pillow": ok_p,
                           "cpu_baseline": {"images_per_s": n / cpu_p, "cores": threads, "kind": "reference"}}}


def run_reference(args, rank):
    if rank != 0:
        return
    n = args.ref_snippets
    for _ in range(args.warmup):
        cpu_reference_pass(n, n)
    t = 0.0
    cores = 0
    for _ in range(args.steps):
        dt, cores = cpu_reference_pass(n, n)
        t += dt
    value = args.steps * n / t
    sample = f"{n} spatial + {n} temporal protocol snippets of one synthetic video per step (batch 10), preprocess+forward+consensus+fusion"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "two-stream 25x10 evaluation, 101 classes (BASELINE configs[2]); CPU reference arm: " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--videos-per-step", type=int, default=2)
    ap.add_argument("--pool", type=int, default=8, help="distinct synthetic videos in the HBM store")
    ap.add_argument("--max-batch", type=int, default=500,
                    help="snippets per internal network chunk (500 = the two videos of a step in one chunk: +6 %% over 125, "
                         "fuller waves on the 14x14 layers; results are bit-identical for any chunk size)")
    ap.add_argument("--ref-snippets", type=int, default=10, help="snippets per stream per CPU-reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-jpeg", dest="e2e_jpeg", action="store_false",
                    help="skip the end-to-end leg that starts from JPEG files (adds a few seconds of cv2 encoding at set-up)")
    ap.add_argument("--workload", default="eval", choices=["eval", "train"],
                    help="eval: the headline two-stream evaluation (BASELINE configs[2]/[3]); train: the training step "
                         "(configs[4], bench_train.py)")
    ap.add_argument("--batch", type=int, default=256,
                    help="--workload train: snippets per stream per GPU per step (BASELINE configs[4]: 256)")
    ap.add_argument("--lr", type=float, default=0.001, help="--workload train: SGD learning rate")
    ap.add_argument("--no-legs", dest="legs", action="store_false",
                    help="skip the extra legs of the line (configs0_b1, configs1_b64, parity, strong_3783, train, tvl1_flow)")
    ap.add_argument("--strong-videos", type=int, default=3783, help="videos of the configs[3] strong-scaling leg")
    args = ap.parse_args()
    if args.workload == "train":
        import bench_train
        if args.steps == 20:
            args.steps = 8
        bench_train.main(args)
        return
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from video_analytics_b200 import _lib, ops
    from video_analytics_b200.combinedModel import CombinedModel
    from video_analytics_b200.distributed import gather_video_rows, init_from_env, shard_bounds
    from video_analytics_b200.evaluate import SNIPPETS_PER_VIDEO, TwoStreamEvaluator, spatial_table, temporal_table
    from video_analytics_b200.spatialModel import build_spatial_torch_model
    from video_analytics_b200.store import DeviceStore, VideoMeta, make_layout
    from video_analytics_b200.temporalModel import build_temporal_torch_model

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device (there is no CPU fallback for the product path)")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the one JSON line (NCCL prints its version banner there)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("VA_DEFER_UPDATE", "1") == "1":
        from video_analytics_b200.distributed import reserve_nccl_ctas
        reserve_nccl_ctas()          # the training step overlaps each stream's all-reduce with the other stream's forward
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    K, W, vps = args.steps, max(args.warmup, 0), args.videos_per_step
    C, D = 101, 256

    # ---- model + data (random-init weights of the reference architecture, synthetic frames; no network here)
    spatial = ops.StreamNet(ops.STREAM_SPATIAL, 3, C, D, max_batch=args.max_batch)
    temporal = ops.StreamNet(ops.STREAM_TEMPORAL, 20, C, D, max_batch=args.max_batch)
    spatial.load_state_dict(build_spatial_torch_model(C, D, seed=0).state_dict())
    temporal.load_state_dict(build_temporal_torch_model(C, 10, D, seed=0).state_dict())
    layout = make_layout(args.pool)
    store = DeviceStore(layout, dev)
    combined = CombinedModel()
    g = torch.Generator().manual_seed(3)
    combined.set_svm(torch.randn(C, 2 * D, generator=g, dtype=torch.float64).numpy() * 0.05,
                     torch.randn(C, generator=g, dtype=torch.float64).numpy() * 0.01)
    ev = TwoStreamEvaluator(spatial, temporal, store, combined)

    # videos of this job: world * (W + K) * vps, contiguous block per rank
    n_videos = world * (W + K) * vps
    lo, hi, per = shard_bounds(n_videos, rank, world)
    out = ev.alloc_outputs(world * per, D, C, with_svm=True)
    my = list(range(lo, hi))

    def step(i):
        vids = my[i * vps:(i + 1) * vps]
        ev.run_videos(vids, out=out, out_row=rank * per + i * vps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()               # nvidia-smi needs ~1 s to produce its first line: start it before the warm-up
    for i in range(W):
        step(i)
    barrier()
    sampler.mark()
    lib = _lib.load()
    lib.va_profile_enable(1)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W, W + K):
        step(i)
    gather_video_rows(out, rank, world, per)
    e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    import ctypes as Ct
    t_ms, t_l, t_f = Ct.c_double(), Ct.c_uint64(), Ct.c_double()
    lib.va_profile_read(Ct.byref(t_ms), Ct.byref(t_l), Ct.byref(t_f))
    lib.va_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    snippets = world * K * vps * SNIPPETS_PER_VIDEO
    value = snippets / (total_ms * 1e-3)

    # ---- end to end: host (pinned) frames -> H2D -> path -> D2H scores, every step, through the public API
    rgb_host = store.rgb.cpu().pin_memory()
    flow_host = store.flow.cpu().pin_memory()
    rgb_img = layout.rgb_shape[0] * layout.rgb_shape[1] * layout.rgb_shape[2]
    flow_img = layout.flow_shape[0] * layout.flow_shape[1] * layout.flow_shape[2]
    max_fr = max(m.n_frames for m in layout.videos)
    max_fl = max(m.n_flows for m in layout.videos)
    stage = DeviceStore.__new__(DeviceStore)
    stage.layout = layout
    stage.rgb = torch.empty(vps * max_fr * rgb_img, dtype=torch.uint8, device=dev)
    stage.flow = torch.empty(vps * 2 * max_fl * flow_img, dtype=torch.uint8, device=dev)
    host_tables = {}

    def staged_tables(k, slot):
        if (k, slot) not in host_tables:
            m = layout.videos[k]
            sm = VideoMeta(m.name, m.category, m.label, m.n_frames, slot * max_fr, m.n_flows, slot * 2 * max_fl,
                           slot * 2 * max_fl + m.n_flows)
            host_tables[(k, slot)] = (torch.from_numpy(spatial_table(sm, layout.rgb_shape)).pin_memory(),
                                      torch.from_numpy(temporal_table(sm, layout.flow_shape)).pin_memory())
        return host_tables[(k, slot)]

    res_host = {"video_scores": torch.empty((vps, C), dtype=torch.float32).pin_memory(),
                "score_pred": torch.empty((vps,), dtype=torch.int32).pin_memory(),
                "svm_pred": torch.empty((vps,), dtype=torch.int32).pin_memory()}
    h2d = d2h = 0

    # The evaluator's own host pipeline (TwoStreamEvaluator.host_pipeline): per group it copies the images the protocol
    # touches (25 frames + 2 x 250 flow images per video) and the index tables from pinned host memory on a copy stream,
    # double-buffered, so the H2D of step i+1 runs under the networks of step i; every step ends with its scores on the host.
    from video_analytics_b200.evaluate import HostStore
    host = HostStore(layout, rgb_host, flow_host)

    def e2e_run(first, last):
        nonlocal h2d, d2h
        groups = [my[i * vps:(i + 1) * vps] for i in range(first, last)]
        # every step's scores / predictions are read back to pinned host memory by the pipeline itself (event-synchronised
        # per step; the next step's kernels are already queued when a step's results are handed out)
        for r in ev.host_pipeline(host, groups, to_host=tuple(res_host)):
            for kname, hbuf in res_host.items():
                hbuf[:r[kname].shape[0]].copy_(r[kname])        # host -> host: the caller's own result buffers
            h2d, d2h = ev.last_h2d_bytes, ev.last_d2h_bytes

    e2e_run(0, W)
    barrier()
    t0 = time.perf_counter()
    e2e_run(W, W + K)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = snippets / float(e2e_s.item())

    # ---- end to end from the reference's ON-DISK format: every step copies the JPEG FILES of the frames
    #      and flow images its two videos touch (written the reference's way, cv2.imwrite) from pinned host memory, the
    #      CUDA decoder fills the stage store on a side stream one step ahead, then K1 -> networks -> fusion -> D2H.
    e2e_jpeg = None
    jpeg_ready, jpeg_why = bool(args.e2e_jpeg), "disabled (--no-e2e-jpeg)"
    if jpeg_ready:
        try:
            import cv2
            from video_analytics_b200 import jpeg
            from video_analytics_b200.utils import test_flow_starts, test_frame_indices
            L = 10
            rgb_np = rgb_host.numpy().reshape(-1, *layout.rgb_shape)
            flow_np = flow_host.numpy().reshape(-1, layout.flow_shape[0], layout.flow_shape[1])
            n_rgb_b, n_flow_b = stage.rgb.numel(), stage.flow.numel()
            stages2, filesets = [], {}
            DEPTH = 3                                   # stage buffers: decode runs two steps ahead of the networks
            for _ in range(DEPTH):
                st2 = DeviceStore.__new__(DeviceStore)
                st2.layout = layout
                st2.all = torch.empty(n_rgb_b + n_flow_b, dtype=torch.uint8, device=dev)
                st2.rgb, st2.flow = st2.all[:n_rgb_b], st2.all[n_rgb_b:]
                stages2.append(st2)

            # Two contents for the same file sets: "video" = what a camera frame / a TV-L1 flow image looks like to a JPEG
            # coder (low-frequency structure + sensor noise: ~25 KB per file, most blocks end early with EOB) and "noise" =
            # the synthetic store's own hash-noise pixels (~67 KB per file, every coefficient non-zero: the entropy decoder's
            # worst case, 3-4x the bits of any real frame).
            import numpy as _np
            _rng = _np.random.default_rng(0)

            def video_like(h, w_, c, k):
                yy, xx = _np.mgrid[0:h, 0:w_]
                base = [128 + 70 * _np.sin(xx / (23.0 + k % 7) + k) * _np.cos(yy / (31.0 + k % 5)), 120 + 60 * _np.cos((xx + yy) / (41.0 + k % 3)),
                        110 + 50 * _np.sin(yy / 19.0 + 0.3 * k)][:c]
                img = (_np.stack(base, -1) + _rng.normal(0, 3.0, (h, w_, c))).clip(0, 255).astype(_np.uint8)
                return img if c == 3 else img[..., 0]
            pool_rgb = [cv2.imencode(".jpg", video_like(layout.rgb_shape[0], layout.rgb_shape[1], 3, k))[1].tobytes() for k in range(32)]
            pool_flow = [cv2.imencode(".jpg", video_like(layout.flow_shape[0], layout.flow_shape[1], 1, k))[1].tobytes() for k in range(64)]
            jpeg_content = ["video"]

            def step_fileset(ks):
                """ONE staged file set (one decoder call) for the images the 25x10 protocol reads from the pool videos `ks`
                of a step (video ks[slot] goes to stage slot `slot`)."""
                key = (jpeg_content[0], ks)
                if key not in filesets:
                    files, offs = [], []
                    noise = jpeg_content[0] == "noise"
                    for slot, k in enumerate(ks):
                        m = layout.videos[k]
                        frames = sorted(set(test_frame_indices(m.n_frames)))
                        flows = sorted({s0 + d for s0 in test_flow_starts(m.n_flows, L) for d in range(L)})     # 1-based
                        for f in frames:
                            files.append(cv2.imencode(".jpg", rgb_np[m.rgb_first + f][..., ::-1])[1].tobytes() if noise
                                         else pool_rgb[(m.rgb_first + f) % len(pool_rgb)])
                            offs.append((slot * max_fr + f) * rgb_img)
                        for first, local0 in ((m.flowx_first, slot * 2 * max_fl), (m.flowy_first, slot * 2 * max_fl + m.n_flows)):
                            for idx in flows:
                                files.append(cv2.imencode(".jpg", flow_np[first + idx - 1])[1].tobytes() if noise
                                             else pool_flow[(first + idx - 1) % len(pool_flow)])
                                offs.append(n_rgb_b + (local0 + idx - 1) * flow_img)
                    filesets[key] = (jpeg.JpegFileSet(files), offs, sum(len(f) for f in files))
                return filesets[key]

            def step_keys(i):
                return tuple(v % len(layout.videos) for v in my[i * vps:(i + 1) * vps])

            for content_ in ("video", "noise"):
                jpeg_content[0] = content_
                for i in range(W + K):
                    step_fileset(step_keys(i))
        except Exception as e:
            jpeg_ready, jpeg_why = False, repr(e)
    if world > 1:                               # every rank takes the same branch: the timed loop below has barriers
        flag = torch.tensor([1 if jpeg_ready else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if jpeg_ready and int(flag.item()) == 0:
            jpeg_ready, jpeg_why = False, "set-up failed on another rank"
    if not jpeg_ready:
        e2e_jpeg = {"unavailable": jpeg_why}
    else:
        try:
            side = torch.cuda.Stream()
            decoded = [torch.cuda.Event() for _ in range(DEPTH)]
            consumed = [torch.cuda.Event() for _ in range(DEPTH)]
            jpeg_bytes = [0]

            # VA_JPEG_MODE=overlap (default): the whole decode (H2D + kernels) of step i+2 runs on a side stream under steps i,
            # i+1.  VA_JPEG_MODE=serial: only the H2D copy runs ahead; the decode kernels of a step run on the compute stream
            # right before its networks.  Measured (snippets/s, video-like / noise content): overlap 15.3-16.6 k / 14.6-15.0 k,
            # serial 11.0-12.1 k / 13.5-15.4 k -- a decoder block lives ~1.5 ms (one image's serial entropy stream) and cannot
            # co-reside with the 220-227 KB kernels (fused conv1_1, FC layers), so overlapped decode costs a step 5-6 ms for
            # 3.3 ms of decode work, but serialising it costs more.  Capping the layer kernels' shared memory at 212 KB so that a
            # decoder block would fit beside them changed nothing (14.2 k vs 14.5 k): their 320-512 threads x 128 registers fill
            # the register file, which is the real co-residency limit.
            jpeg_mode = os.environ.get("VA_JPEG_MODE", "overlap")
            staged_up = {}

            def issue_decode(i):
                slot2 = i % DEPTH
                fs_, offs_, nb = step_fileset(step_keys(i))
                if jpeg_mode == "serial":
                    with torch.cuda.stream(side):
                        side.wait_event(consumed[slot2])
                    staged_up[i] = fs_.upload(dev, side)
                    return nb
                with torch.cuda.stream(side):
                    side.wait_event(consumed[slot2])
                    fs_.decode_into(stages2[slot2].all, offs_)
                    decoded[slot2].record(side)
                return nb

            def e2e_jpeg_step(i, end):
                slot2 = i % DEPTH
                cur = torch.cuda.current_stream()
                vids = my[i * vps:(i + 1) * vps]
                tabs_s, tabs_t = [], []
                for slot, v in enumerate(vids):
                    hs, ht = staged_tables(v % len(layout.videos), slot)
                    tabs_s.append(hs.to(dev, non_blocking=True)); tabs_t.append(ht.to(dev, non_blocking=True))
                if jpeg_mode == "serial":
                    fs_, offs_, _ = step_fileset(step_keys(i))
                    fs_.decode_into(stages2[slot2].all, offs_, staged=staged_up.pop(i))
                else:
                    cur.wait_event(decoded[slot2])
                if i + DEPTH - 1 < end:
                    jpeg_bytes[0] = issue_decode(i + DEPTH - 1)    # runs under this and the next step's networks
                r = ev.run_tables(torch.cat(tabs_s), torch.cat(tabs_t), len(vids), store=stages2[slot2])
                consumed[slot2].record(cur)
                for kname, hbuf in res_host.items():
                    hbuf[:len(vids)].copy_(r[kname], non_blocking=True)
                cur.synchronize()

            def timed_jpeg_run(content_):
                jpeg_content[0] = content_
                torch.cuda.synchronize()
                for s_ in range(DEPTH):
                    consumed[s_].record(torch.cuda.current_stream())
                nw = min(W, 2)
                for i in range(min(DEPTH - 1, nw)):
                    issue_decode(i)
                for i in range(nw):
                    e2e_jpeg_step(i, nw)
                barrier()
                t0 = time.perf_counter()
                for i in range(W, min(W + DEPTH - 1, W + K)):
                    issue_decode(i)
                for i in range(W, W + K):
                    e2e_jpeg_step(i, W + K)
                barrier()
                ej = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(ej, op=dist.ReduceOp.MAX)
                first = my[W * vps:(W + 1) * vps]
                sets_ = [step_fileset(step_keys(W))]
                tab_bytes = sum(t.numel() * 4 for s_, v in enumerate(first) for t in staged_tables(v % len(layout.videos), s_))
                return {"value": snippets / float(ej.item()), "unit": UNIT,
                        "h2d_bytes_per_step": int(sum(x[2] for x in sets_) + tab_bytes), "d2h_bytes_per_step": d2h,
                        "images_decoded_per_step": int(sum(x[0].n for x in sets_))}

            # VA_JPEG_PRIORITY=1 runs the networks on a HIGH-priority stream in this leg (the block scheduler then places a
            # layer's CTAs before pending decoder blocks).  Measured: worse -- the decoder is starved and the step waits for
            # it (video-like content 16.6 k -> 11.5 k snippets/s, noise unchanged at 14.7 k), so it is off.
            prio = os.environ.get("VA_JPEG_PRIORITY", "0") == "1"
            hi_stream = torch.cuda.Stream(priority=-1) if prio else None

            def with_priority(fn, *a):
                if hi_stream is None:
                    return fn(*a)
                hi_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(hi_stream):
                    r_ = fn(*a)
                torch.cuda.current_stream().wait_stream(hi_stream)
                return r_

            e2e_jpeg = with_priority(timed_jpeg_run, "video")
            e2e_jpeg["network_stream_priority"] = "high" if prio else "default"
            e2e_jpeg["content"] = ("video-like frames and flow images (low-frequency structure + noise sigma 3, ~25 KB per file: the "
                                   "bit rate of real UCF101 frames); network inputs are whatever the files decode to")
            e2e_jpeg["input"] = ("JPEG files (cv2.imwrite format) in pinned host memory -> H2D (copy stream, two steps ahead) -> CUDA "
                                 "decode -> gather + networks -> fusion -> D2H; decode kernels "
                                 + ("on the compute stream before the step's networks" if jpeg_mode == "serial" else
                                    "on a side stream under the previous steps' networks"))
            worst = with_priority(timed_jpeg_run, "noise")
            worst["content"] = ("the synthetic store's hash-noise pixels (~67 KB per file, no EOB, every coefficient non-zero): the "
                                "entropy decoder's worst case")
            e2e_jpeg["worst_case_noise"] = worst
        except Exception as e:      # keep the contract line alive: report why this optional leg is missing
            e2e_jpeg = {"unavailable": repr(e)}

    _sd = {}

    def stream_state_dict(kind):
        """fp32 state_dict of the seed-0 torch module the handles were loaded from (built once, shared by the legs below)."""
        if kind not in _sd:
            _sd[kind] = (build_spatial_torch_model(C, D, seed=0) if kind == "s" else build_temporal_torch_model(C, 10, D, seed=0)).state_dict()
        return _sd[kind]

    # ---- HBM-bound kernels of the path, timed alone (CUDA events, inputs >> L2): algorithmic bytes of SURVEY.md 8d
    aux = []
    if rank == 0:
        def timed(fn, reps=5):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        nvid = 4
        tabs = [ev.tables_for(v) for v in range(nvid)]
        ts_, tt_ = torch.cat([t[0] for t in tabs]), torch.cat([t[1] for t in tabs])
        nsn = ts_.shape[0]
        ms_s = timed(lambda: ops.preprocess(store.rgb, layout.rgb_shape, ts_, ev.mean_s, ev.std_s, c_pad=spatial.c_pad))
        ms_t = timed(lambda: ops.preprocess(store.flow, layout.flow_shape, tt_, ev.mean_t, ev.std_t, c_pad=temporal.c_pad))
        gv = 512
        gen = torch.Generator(device="cuda").manual_seed(1)
        fd = [torch.rand((gv * SNIPPETS_PER_VIDEO, D), device=dev, generator=gen) for _ in range(2)]
        fs = [torch.rand((gv * SNIPPETS_PER_VIDEO, C), device=dev, generator=gen) for _ in range(2)]
        offs = torch.arange(0, (gv + 1) * SNIPPETS_PER_VIDEO, SNIPPETS_PER_VIDEO, dtype=torch.int32, device=dev)
        fout = ev.alloc_outputs(gv, D, C, with_svm=True)
        ms_f = timed(lambda: combined.fuse(fd[0], fd[1], fs[0], fs[1], offs, out=fout))
        # the same launch without the fp64 SVM scoring (F2): what is left is the byte-bound part of K4 -- consensus means
        # and score fusion (C1, F1, X1); the SVM adds 101 x 512 fp64 MACs per video, compute and L2 reads, not HBM bytes
        fout2 = ev.alloc_outputs(gv, D, C, with_svm=False)
        ms_f2 = timed(lambda: ops.fuse(fd[0], fd[1], fs[0], fs[1], offs, w_s=combined.w_s, w_t=combined.w_t, out=fout2))
        hbm = read_peaks()["hbm"]
        px = 224 * 224
        fuse_b = (714_000 + (2 * D + C) * 4 + 8 + C * 8) * gv
        # moved = bytes the launch really reads + writes: the network-input layout pads 3 -> 16 and 20 -> 32 channels
        # (TMA / UMMA K granularity), so K1 writes more than the algorithmic bf16 tensor of SURVEY.md 8d
        for name, nbytes, moved, ms_k in (
                ("preprocess_rows_kernel (RGB, 3->16ch bf16 NHWC)", 451_584 * nsn, (px * 3 + px * spatial.c_pad * 2) * nsn, ms_s),
                ("preprocess_rows_kernel (flow stack, 20->32ch bf16 NHWC)", 3_010_560 * nsn, (px * 20 + px * temporal.c_pad * 2) * nsn, ms_t),
                ("fuse_kernel (consensus + late fusion)", fuse_b - C * 8 * gv, fuse_b - C * 8 * gv, ms_f2),
                ("fuse_kernel (consensus + late fusion + fp64 SVM scoring)", fuse_b, fuse_b, ms_f)):
            gbs = nbytes / (ms_k * 1e-3) / 1e9
            aux.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                        "ms_per_launch": ms_k, "algorithmic_bytes_per_launch": nbytes,
                        "moved_bytes_per_launch": moved, "moved_gbs": moved / (ms_k * 1e-3) / 1e9,
                        "moved_frac": moved / (ms_k * 1e-3) / 1e9 / hbm})
        for a_ in aux[:2]:
            a_["on_path"] = ("training loader and reference-layout __getitem__ only: the evaluation step of this line runs "
                             "conv1_fused_kernel (next two entries), which never materialises the padded tensor")
        del fd, fs, fout, fout2
        # the default front end of the evaluation path (va_forward_store): index-table gather + normalise + conv1_1 in one
        # kernel.  Algorithmic bytes per snippet = the cropped u8 it reads (150 528 / 1 003 520 B) + the bf16
        # [224,224,64] layer output it writes (6 422 528 B) -- the write dominates, so the bound is HBM, not the tensor pipe
        try:
            t1s, t1t = tabs[0]
            wsd = {"s": stream_state_dict("s"), "t": stream_state_dict("t")}
            for name, key, images, shape, tab, mean, std, rd in (
                    ("conv1_fused_kernel (RGB: gather + normalise + conv1_1 + ReLU, u8 store -> bf16 NHWC 64ch)", "s", store.rgb,
                     layout.rgb_shape, t1s, ev.mean_s, ev.std_s, 150_528),
                    ("conv1_fused_kernel (flow stack: 20 planes gathered per snippet, same epilogue)", "t", store.flow,
                     layout.flow_shape, t1t, ev.mean_t, ev.std_t, 1_003_520)):
                w1 = wsd[key]["features.0.weight"].float().to(dev).contiguous()
                b1 = wsd[key]["features.0.bias"].float().to(dev).contiguous()
                ms_c = timed(lambda: ops.conv1_fused(images, shape, tab, mean, std, w1, b1))
                nb = (rd + px * 64 * 2) * tab.shape[0]
                gbs = nb / (ms_c * 1e-3) / 1e9
                aux.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                            "ms_per_launch": ms_c, "snippets_per_launch": int(tab.shape[0]), "algorithmic_bytes_per_launch": nb,
                            "moved_bytes_per_launch": nb, "moved_gbs": gbs, "moved_frac": gbs / hbm,
                            "write_share_of_bytes": px * 64 * 2 / (rd + px * 64 * 2),
                            "on_path": "evaluation (TwoStreamEvaluator default front end)"})
            # `peak` above is the read+write COPY figure; a stream of stores alone gets far less of it on this part
            # (tools/bench_write_bw.py, profiles/r02_hbm_write_only_bw.json: memset and a fill kernel both 3.86 TB/s beside a
            # 6.50 TB/s copy).  Measured live on a buffer the size of one launch's output, so the store-dominated entries
            # can be read against the ceiling that applies to them.
            wbuf = torch.empty(250 * px * 64, dtype=torch.bfloat16, device=dev)
            ms_w = timed(lambda: wbuf.zero_())
            w_gbs = wbuf.numel() * 2 / (ms_w * 1e-3) / 1e9
            del wbuf
            for a_ in aux[-2:]:
                a_["write_only_peak"] = w_gbs
                a_["frac_of_write_only_peak"] = a_["achieved"] / w_gbs
            del w1, b1
        except Exception as e:      # an evidence entry, not the contract: say why it is missing
            aux.append({"kernel": "conv1_fused_kernel", "unavailable": repr(e)})
        jpeg_line = jpeg_decode_measurement(store, layout, dev)

    # ---- the reference's CPU path beside it (rank 0, N = 1): the oracle port on the box's host cores, bounded sample
    cpu_line, oracle_logits = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = args.ref_snippets
        cpu_reference_pass(n, n)                                   # warm-up
        t, reps, cores = 0.0, 0, 0
        while t < 12.0 and reps < 6:
            dt, cores = cpu_reference_pass(n, n)
            t += dt
            reps += 1
        cpu_line = {"value": reps * n / t, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{reps} x ({n} spatial + {n} temporal protocol snippets, batch 10) of one synthetic video: "
                              "CPU preprocess + VGG16 fp32 forward + consensus + fusion (oracle/two_stream.py)"}
        oracle_logits = cpu_reference_pass.__dict__["state"].get("last_logits")

    # ---- the other BASELINE configs, in the same line (bench_legs.py): collective legs run on every rank
    legs = {}
    if args.legs:
        import bench_legs
        sd_s, sd_t = stream_state_dict("s"), stream_state_dict("t")
        for key, fn in (
                ("small", (lambda: bench_legs.small_batch_legs(spatial, temporal, ev, store, layout)) if world == 1 else None),
                ("parity", (lambda: bench_legs.parity_leg(spatial, temporal, ev, store, layout, sd_s, sd_t,
                                                          oracle_logits=oracle_logits)) if world == 1 else None),
                ("strong_3783", lambda: bench_legs.strong_leg(ev, lambda rows: ev.alloc_outputs(rows, D, C, with_svm=True), rank, world,
                                                              dev, n_videos=args.strong_videos, vps=vps)),
                ("train", lambda: bench_legs.train_leg(args)),
                ("tvl1_flow", (lambda: bench_legs.flow_leg(cpu_baseline=not args.no_cpu_baseline)) if world == 1 else None)):
            if fn is None:
                continue
            try:
                legs[key] = fn()
            except Exception as e:                 # a failing leg must not take the contract line down with it
                if world > 1:
                    raise                           # ...except under torchrun, where a rank that skips a collective hangs the rest
                legs[key] = {"unavailable": repr(e)[:300]}
            torch.cuda.empty_cache()

    if rank == 0:
        peaks = read_peaks()
        achieved = (t_f.value / 1e12) / (t_ms.value * 1e-3) if t_ms.value > 0 else 0.0
        traffic = None      # DRAM bytes per layer-kernel launch from the committed ncu --set full capture (profiles/)
        tpath = os.path.join(ROOT, "profiles", "r02_ncu_conv_tc_traffic_final.json")
        if os.path.exists(tpath):
            # the capture ran one 250-snippet chunk; a launch of this run processes `chunk` snippets: activation bytes scale
            # with the chunk, the weights (0.27 GB of 9.2 GB per 250-snippet chunk) do not -- scaled linearly, 3 % high
            tj = json.load(open(tpath))
            captured = tj.get("dram_bytes_per_launch_avg")
            chunk = min(args.max_batch, vps * SNIPPETS_PER_VIDEO)
            traffic = captured * chunk / float(tj.get("chunk_snippets", 250)) if captured is not None else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "two-stream 25 snippets x 10 crops per video evaluation, 101 classes, VGG16 spatial(3ch)+temporal(20ch), "
                                   "late fusion (BASELINE configs[2]; configs[3] sharding at N>1)",
                       "videos_per_step_per_gpu": vps, "snippets_per_video_per_stream": SNIPPETS_PER_VIDEO,
                       "global_videos_timed": world * K * vps, "parallelism": f"video-shard x{world} + all-gather of fused scores",
                       "l2_policy": "inputs larger than L2: each step streams ~1.2 GB of preprocessed snippets and >800 MB activations per chunk",
                       "weights": "random init (seed 0) of the reference architecture", "max_batch": args.max_batch},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel / conv_tc2_kernel / conv_tc2h_kernel (tcgen05 implicit-GEMM conv/FC)", "achieved": achieved,
                         "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic,
                         "traffic_source": "profiles/r02_ncu_conv_tc_traffic_final.json (ncu --set full of the layer launches of one 250-snippet "
                                           "spatial chunk: dram__bytes_read + dram__bytes_write, averaged per launch), scaled to this run's "
                                           "snippets per launch",
                         "tensor_pipe_pct": 79.5,
                         "tensor_pipe_source": "profiles/r02_ncu_layer_table_final.md: sm__pipe_tensor_cycles_active (pct of peak, elapsed), "
                                               "duration-weighted over the layer launches of one 250-snippet chunk under ncu; "
                                               "89.4-94.7 % on conv2_1..conv4_3",
                         "peak_source": peaks["src"], "launches": int(t_l.value), "avg_launch_ms": t_ms.value / max(1, t_l.value),
                         "flops_per_launch": t_f.value / max(1, t_l.value),
                         "share_of_step": t_ms.value / total_ms if total_ms > 0 else None},
            "aux_rooflines": aux,
            "jpeg_decode": jpeg_line,
            "e2e_jpeg": e2e_jpeg,
            "clocks": clocks,
        }
        if legs:
            small = legs.get("small")
            if isinstance(small, tuple):
                line["configs0_b1"], line["configs1_b64"] = small
            elif small is not None:
                line["configs0_b1"] = line["configs1_b64"] = small
            for key in ("parity", "strong_3783", "train", "tvl1_flow"):
                if legs.get(key) is not None:
                    line[key] = legs[key]
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

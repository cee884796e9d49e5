#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[4]: two-stream training step bf16, gradient all-reduce over NVLink).

`python bench.py --workload train [--gpus N --steps K --warmup W --batch B]` (default B = 256) dispatches here; same launch contract and
JSON line as bench.py.  One step, per GPU: B RGB snippets + B flow-stack snippets are cropped/flipped/normalised from
the uint8 store (K1), each stream runs forward + backward + ONE NCCL all-reduce of its gradient arena + the fused
SGD-momentum update (training.py).  `value` counts two-stream snippets (one RGB + one flow stack) over all ranks.

The reference arm for this workload is the oracle's train_step (torch CPU fp32, reference loop body) on a small batch.
"""
from __future__ import annotations

import json
import os
import random
import time

METRIC = "two-stream training snippets/sec (device-timed, max over ranks)"
UNIT = "snippets/s"
ROOT = os.path.dirname(os.path.abspath(__file__))

# algorithmic FLOPs of one VGG16 stream forward per snippet (2*MAC): conv stack + 3 tensor-core FC layers
_CONV = [(224, 64, None), (224, 64, 64), (112, 128, 64), (112, 128, 128), (56, 256, 128), (56, 256, 256), (56, 256, 256),
         (28, 512, 256), (28, 512, 512), (28, 512, 512), (14, 512, 512), (14, 512, 512), (14, 512, 512)]


def stream_flops(cin, desc_dim=256):
    """(forward, backward) algorithmic FLOPs per snippet: backward = dgrad (all layers but the first conv) + wgrad."""
    fwd = bwd = 0.0
    for i, (hw, co, ci) in enumerate(_CONV):
        f = 2.0 * hw * hw * co * 9 * (ci if ci else cin)
        fwd += f
        bwd += f * (2 if i > 0 else 1)
    for fi, fo in ((25088, 4096), (4096, 4096), (4096, desc_dim)):
        fwd += 2.0 * fi * fo
        bwd += 4.0 * fi * fo
    return fwd, bwd


def cpu_train_pass(batch, threads=None):
    """One two-stream training step of the CPU oracle (reference loop body, torch fp32) on `batch` snippets per stream."""
    import torch
    from oracle import two_stream as ts
    if threads:
        torch.set_num_threads(threads)
    state = cpu_train_pass.__dict__.setdefault("state", {})
    if not state:
        for kind, cin in (("spatial", 3), ("temporal", 20)):
            model = ts.build_spatial_model(seed=0) if kind == "spatial" else ts.build_temporal_model(seed=0)
            state[kind] = (model, torch.optim.SGD(model.parameters(), 0.1, momentum=0.9), cin)
    g = torch.Generator().manual_seed(0)
    t0 = time.perf_counter()
    for kind, (model, opt, cin) in state.items():
        ip = torch.randn(batch, cin, 224, 224, generator=g)
        labels = torch.randint(1, 101, (batch,), generator=g)
        masks = ts.draw_dropout_masks([(batch, 4096), (batch, 4096), (batch, 256)])
        ts.train_step(model, opt, torch.nn.CrossEntropyLoss(), ip, labels, masks)
    return time.perf_counter() - t0, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    n = args.ref_snippets
    K, W = args.steps, max(args.warmup, 0)
    for _ in range(min(W, 1)):
        cpu_train_pass(n)
    t, cores = 0.0, 0
    steps = min(K, 3)
    for _ in range(steps):
        dt, cores = cpu_train_pass(n)
        t += dt
    v = steps * n / t
    sample = f"{steps} x one two-stream SGD step on {n} snippets per stream (oracle/two_stream.py train_step, torch CPU fp32)"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                      "warmup": min(W, 1), "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "two-stream training step (BASELINE configs[4])", "batch_per_step": n},
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main(args, embedded=False):
    """embedded=True (bench_legs.train_leg): return the line on rank 0 instead of printing it, keep the process group."""
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, read_peaks
    from video_analytics_b200 import _lib, ops, train_ops
    from video_analytics_b200.distributed import init_from_env
    from video_analytics_b200.parameters import FLOW_NORM_MEAN, FLOW_NORM_STD, NORM_MEANS_TF, NORM_STDS_TF
    from video_analytics_b200.spatialModel import build_spatial_torch_model
    from video_analytics_b200.store import DeviceStore, make_layout
    from video_analytics_b200.temporalModel import build_temporal_torch_model
    from video_analytics_b200.training import StreamTrainer

    if not torch.cuda.is_available():
        raise SystemExit("bench_train.py needs a B200: no CUDA device (there is no CPU fallback for the product path)")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("VA_DEFER_UPDATE", "1") == "1":
        from video_analytics_b200.distributed import reserve_nccl_ctas
        reserve_nccl_ctas()          # the training step overlaps each stream's all-reduce with the other stream's forward
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    K, W, B = args.steps, max(args.warmup, 0), args.batch
    C, D, L = 101, 256, 10
    group = dist.group.WORLD if world > 1 else None
    # same initial weights on every rank (seeded init), as DataParallel's replicate would give
    # data parallel: each stream's gradient all-reduce runs under the OTHER stream's forward pass (deferred update + SM
    # reservation, training.py); VA_DEFER_UPDATE=0 gives the collective-after-backward schedule for A/B runs
    defer = world > 1 and os.environ.get("VA_DEFER_UPDATE", "1") == "1"
    # the gradient all-reduce is the library's own two-shot kernel over the NVSwitch multicast address of a symmetric
    # arena (csrc/va_allreduce.cu); VA_ALLREDUCE_IMPL=nccl gives torch.distributed's (measured equal at N = 2 and N = 8)
    impl = os.environ.get("VA_ALLREDUCE_IMPL", "va")
    tr_s = StreamTrainer(build_spatial_torch_model(C, D, seed=0), None, lr=args.lr, momentum=0.9, c_pad=16, process_group=group,
                         defer_update=defer, allreduce_impl=impl)
    tr_t = StreamTrainer(build_temporal_torch_model(C, L, D, seed=0), None, lr=args.lr, momentum=0.9, c_pad=32, process_group=group,
                         defer_update=defer, allreduce_impl=impl)
    layout = make_layout(args.pool)
    store = DeviceStore(layout, dev)
    mean_s, std_s = list(NORM_MEANS_TF), list(NORM_STDS_TF)
    mean_t, std_t = [FLOW_NORM_MEAN] * (2 * L), [FLOW_NORM_STD] * (2 * L)
    rgb_img = int(np.prod(layout.rgb_shape))
    flow_img = int(np.prod(layout.flow_shape))

    # training-mode sampling (spatialModel.py:60-81 / temporalModel.py:69-90): one random frame / flow start per video,
    # random 224-crop + flip; drawn on the host once per step for this rank's B snippets
    rng = random.Random(1234 + rank)

    def draw(step):
        rows_s, rows_t, labels, src = [], [], [], []
        for b in range(B):
            m = layout.videos[(step * B * world + rank * B + b) % len(layout.videos)]
            f = rng.randint(0, m.n_frames - 1)
            s0 = rng.randint(1, m.n_flows - L + 1)
            i, j, fl = rng.randint(0, layout.rgb_shape[0] - 224), rng.randint(0, layout.rgb_shape[1] - 224), rng.randint(0, 1)
            rows_s.append([[m.rgb_first + f, i, j, fl]])
            i, j, fl = rng.randint(0, layout.flow_shape[0] - 224), rng.randint(0, layout.flow_shape[1] - 224), rng.randint(0, 1)
            rows_t.append([[fid, i, j, fl] for idx in range(s0, s0 + L) for fid in (m.flowx_first + idx - 1, m.flowy_first + idx - 1)])
            labels.append(m.label)
            src.append((m.rgb_first + f, m.flowx_first + s0 - 1, m.flowy_first + s0 - 1))
        return (torch.tensor(rows_s, dtype=torch.int32), torch.tensor(rows_t, dtype=torch.int32),
                torch.tensor(labels, dtype=torch.int64), src)

    plan = [draw(i) for i in range(W + K)]
    dev_plan = [(a.to(dev), b.to(dev), c.to(dev)) for a, b, c, _ in plan]
    losses = []

    def step(i):
        ts_, tt_, lab = dev_plan[i]
        xs = ops.preprocess(store.rgb, layout.rgb_shape, ts_, mean_s, std_s, c_pad=16)
        ls, _, _ = tr_s.step(xs, lab)
        del xs
        xt = ops.preprocess(store.flow, layout.flow_shape, tt_, mean_t, std_t, c_pad=32)
        lt, _, _ = tr_t.step(xt, lab)
        losses.append((ls, lt))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(W):
        step(i)
    tr_s.flush()
    tr_t.flush()
    barrier()
    sampler.mark()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W, W + K):
        step(i)
    tr_s.flush()                 # deferred updates of the last step belong to the timed region
    tr_t.flush()
    e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    snippets = world * K * B
    value = snippets / (total_ms * 1e-3)
    loss_first, loss_last = [float(x) for x in losses[0]], [float(x) for x in losses[-1]]

    # ---- end to end: the step's uint8 frames come from pinned host memory, the losses go back to the host
    rgb_host = store.rgb.cpu().pin_memory()
    flow_host = store.flow.cpu().pin_memory()
    # two staging buffers + a copy stream: the H2D copies of step i+1 run under the compute of step i (every timed
    # step's copy is issued inside the timed region; only the first one is not overlapped)
    stages = []
    for _ in range(2):
        st_ = DeviceStore.__new__(DeviceStore)
        st_.layout = layout
        st_.rgb = torch.empty(B * rgb_img, dtype=torch.uint8, device=dev)
        st_.flow = torch.empty(B * 2 * L * flow_img, dtype=torch.uint8, device=dev)
        stages.append(st_)
    copy_stream = torch.cuda.Stream()
    copied = [torch.cuda.Event(), torch.cuda.Event()]        # stage slot filled
    consumed = [torch.cuda.Event(), torch.cuda.Event()]      # stage slot read by the preprocess kernels
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    h2d = d2h = 0

    def staged_tables(i):
        a, b, c, src = plan[i]
        a, b = a.clone(), b.clone()
        for n_ in range(B):
            a[n_, 0, 0] = n_
            for l_ in range(L):
                b[n_, 2 * l_, 0] = n_ * 2 * L + l_
                b[n_, 2 * l_ + 1, 0] = n_ * 2 * L + L + l_
        return a.pin_memory(), b.pin_memory(), c.pin_memory(), src

    staged = {i: staged_tables(i) for i in range(W + K)}
    dev_tabs = {}

    def issue_copies(i):
        """H2D of step i's frames, index tables and labels on the copy stream; returns the byte count."""
        slot = i & 1
        a, b, c, src = staged[i]
        stage = stages[slot]
        nb = 0
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for n_, (fr, fx, fy) in enumerate(src):
                stage.rgb[n_ * rgb_img:(n_ + 1) * rgb_img].copy_(rgb_host[fr * rgb_img:(fr + 1) * rgb_img], non_blocking=True)
                o = n_ * 2 * L * flow_img
                stage.flow[o:o + L * flow_img].copy_(flow_host[fx * flow_img:(fx + L) * flow_img], non_blocking=True)
                stage.flow[o + L * flow_img:o + 2 * L * flow_img].copy_(flow_host[fy * flow_img:(fy + L) * flow_img], non_blocking=True)
                nb += rgb_img + 2 * L * flow_img
            dev_tabs[i] = (a.to(dev, non_blocking=True), b.to(dev, non_blocking=True), c.to(dev, non_blocking=True))
            nb += a.numel() * 4 + b.numel() * 4 + c.numel() * 8
            copied[slot].record(copy_stream)
        return nb

    def e2e_step(i, last, count=False):
        nonlocal h2d, d2h
        slot = i & 1
        cur = torch.cuda.current_stream()
        cur.wait_event(copied[slot])
        ts_, tt_, lab = dev_tabs.pop(i)
        for t_ in (ts_, tt_, lab):
            t_.record_stream(cur)
        stage = stages[slot]
        xs = ops.preprocess(stage.rgb, layout.rgb_shape, ts_, mean_s, std_s, c_pad=16)
        xt = ops.preprocess(stage.flow, layout.flow_shape, tt_, mean_t, std_t, c_pad=32)
        consumed[slot].record(cur)
        nb = 0
        if not last:
            nb = issue_copies(i + 1)             # overlaps with this step's forward/backward
        ls, _, _ = tr_s.step(xs, lab)
        del xs
        lt, _, _ = tr_t.step(xt, lab)
        del xt
        loss_host[0:1].copy_(ls, non_blocking=True)
        loss_host[1:2].copy_(lt, non_blocking=True)
        cur.synchronize()                        # the step's result is on the host
        if count:
            h2d, d2h = nb, 8

    for s_ in (0, 1):
        consumed[s_].record(torch.cuda.current_stream())
    n_warm = min(W, 2)
    if n_warm:
        issue_copies(0)
        for i in range(n_warm):
            e2e_step(i, last=(i == n_warm - 1))
        tr_s.flush()
        tr_t.flush()
    barrier()
    t0 = time.perf_counter()
    issue_copies(W)
    for i in range(W, W + K):
        e2e_step(i, last=(i == W + K - 1), count=(i == W))
    tr_s.flush()
    tr_t.flush()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = snippets / float(e2e_s.item())

    # ---- roofline of the tensor-core GEMM calls (layer kernel forward / data gradient, weight-gradient GEMM incl. its
    #      layout passes), timed live with CUDA events around each call of one extra step on rank 0
    roof = None
    if rank == 0:
        spans = []

        def wrap(mod, name):
            fn = getattr(mod, name)

            def timed(*a, **k):
                x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                x0.record()
                r = fn(*a, **k)
                x1.record()
                spans.append((name, x0, x1))
                return r
            setattr(mod, name, timed)
            return fn
        saved = {(m, n_): wrap(m, n_) for m, names in ((ops, ("conv2d_nhwc", "linear")),
                                                       (train_ops, ("conv2d_dgrad", "conv2d_wgrad", "linear_dgrad", "linear_wgrad")))
                 for n_ in names}
        if world == 1:
            step(W + K - 1)
        else:   # no collective outside the lock-step region: forward/backward only
            ts_, tt_, lab = dev_plan[W + K - 1]
            tr_s.forward_backward(ops.preprocess(store.rgb, layout.rgb_shape, ts_, mean_s, std_s, c_pad=16), lab)
            tr_t.forward_backward(ops.preprocess(store.flow, layout.flow_shape, tt_, mean_t, std_t, c_pad=32), lab)
        torch.cuda.synchronize()
        for (m, n_), fn in saved.items():
            setattr(m, n_, fn)
        by = {}
        for name, x0, x1 in spans:
            by[name] = by.get(name, 0.0) + x0.elapsed_time(x1)
        tensor_ms = sum(by.values())
        fs, bs = stream_flops(3)
        ft, bt = stream_flops(20)
        flops = B * (fs + bs + ft + bt)
        peaks = read_peaks()
        achieved = flops / (tensor_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "tcgen05 GEMM calls of the step: conv_tc/conv_tc2 (forward, data gradient) + wgrad_tc "
                                             "(weight gradient, spans include its layout passes)",
                "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "traffic": None, "peak_source": peaks["src"], "launches": len(spans), "avg_launch_ms": tensor_ms / max(1, len(spans)),
                "flops_per_step": flops, "ms_by_call": {k: round(v, 3) for k, v in by.items()},
                "share_of_step": tensor_ms / (total_ms / K)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "two-stream training step: VGG16 spatial(3ch)+temporal(20ch), 101 classes, CE + SGD-momentum 0.9, "
                                   "Dropout 0.5 (BASELINE configs[4])",
                       "lr": args.lr, "lr_note": "the reference's 0.1 diverges within a few steps on random-init weights and noise "
                                                 "frames (its runs start from ImageNet weights); the arithmetic per step does not depend on lr",
                       "batch_per_gpu_per_stream": B, "global_batch": world * B,
                       "parallelism": f"data parallel x{world}: NCCL all-reduce of the flat gradient arena per stream per step, "
                                      f"{tr_s.grad_allreduce_dtype} payload ({tr_s.flat_grad.numel() * (2 if tr_s.grad_allreduce_dtype == 'bf16' else 4) / 1e6:.0f}"
                                      f" + {tr_t.flat_grad.numel() * (2 if tr_t.grad_allreduce_dtype == 'bf16' else 4) / 1e6:.0f} MB), "
                                      + ("classifier slice launched as soon as its gradients exist" if tr_s.overlap_allreduce else
                                         (f"one collective per stream, run under the OTHER stream's forward pass (deferred update; "
                                          f"{tr_s.reserve_sms} SMs reserved for the collective during {tr_s.reserve_launches} layer launches)"
                                          + (", own two-shot all-reduce kernel over " + ("NVSwitch multicast" if tr_s._symm_use_mc else "NVLink peer pointers")
                                             if tr_s._symm is not None else ", NCCL")
                                          if tr_s.defer_update else
                                          "one collective after the backward pass (the persistent layer kernels hold every SM)")),
                       "l2_policy": "inputs larger than L2: activations of one step are several GB",
                       "weights": "random init (seed 0) of the reference architecture, fp32 master copies",
                       "loss_first_step": loss_first, "loss_last_step": loss_last},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roof,
            "clocks": clocks,
        }
        if embedded:
            return line
        if world == 1 and not args.no_cpu_baseline:
            n = args.ref_snippets
            cpu_train_pass(n)
            t, reps, cores = 0.0, 0, 0
            while t < 12.0 and reps < 4:
                dt, cores = cpu_train_pass(n)
                t += dt
                reps += 1
            line["cpu_baseline"] = {"value": reps * n / t, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{reps} x one two-stream SGD step on {n} snippets per stream "
                                              "(oracle/two_stream.py train_step: torch CPU fp32 forward + backward + SGD)"}
        print(json.dumps(line))
    if embedded:
        return None
    if world > 1:
        dist.destroy_process_group()

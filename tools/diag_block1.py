#!/usr/bin/env python
"""Block-1 experiments (conv1_2 kernel variants): correctness against torch fp32 and against the plain one-tap-per-stage
kernel, plus event-timed duration at the bench's chunk size.  Each variant runs in its own subprocess so that a device
trap (e.g. a descriptor the hardware rejects) cannot poison the others.
  python tools/diag_block1.py [batch]        -> all variants
  python tools/diag_block1.py case <force_r> <batch>
force_r: 1 plain, 0 default, 3 R=3 (vertical tap reuse), 10 HALO (one haloed box per tile), 11 CTA-pair HALO + resident weights
"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

LAYERS = {"conv1_2": (224, 64, 64, True), "conv2_1": (112, 64, 128, False), "conv2_2": (112, 128, 128, True)}
VARIANTS = {"conv1_2": [1, 3, 10, 11, 0], "conv2_1": [1, 3, 11, 0], "conv2_2": [1, 3, 11, 0]}


def run_case(force_r, batch, layer="conv1_2"):
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from video_analytics_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(99)
    dev = "cuda"
    # correctness on 3 images (partial last tiles do not exist at 224, but image borders and batch edges do)
    H, cin, cout, layer_pool = LAYERS[layer]
    n = 3
    xc = torch.randn(n, cin, H, H, generator=g).to(dev).bfloat16()
    x = xc.permute(0, 2, 3, 1).contiguous()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).to(dev)
    b = (torch.randn(cout, generator=g) * 0.1).to(dev)
    res = dict(layer=layer, force_r=force_r, batch=batch)
    for pool in (True, False):
        y = ops.conv2d_nhwc(x, w, b, relu=True, pool=pool, force_bn=0, force_r=force_r).float()
        torch.cuda.synchronize()
        ref = torch.relu(torch.nn.functional.conv2d(xc.float(), w.bfloat16().float(), b, padding=1))
        if pool:
            ref = torch.nn.functional.max_pool2d(ref, 2, 2)
        ref = ref.permute(0, 2, 3, 1).contiguous()
        diff = (y - ref).abs()
        tol = 2.0 ** -7 * ref.abs() + 2e-2 * ref.abs().mean()
        res[f"bad_frac_pool{int(pool)}"] = float((diff > tol).float().mean())
        res[f"max_abs_pool{int(pool)}"] = float(diff.max())
        y1 = ops.conv2d_nhwc(x, w, b, relu=True, pool=pool, force_bn=0, force_r=1).float()
        res[f"bitequal_plain_pool{int(pool)}"] = bool(torch.equal(y, y1))
        res[f"maxdiff_plain_pool{int(pool)}"] = float((y - y1).abs().max())
    # timing at the bench's chunk size
    xb = torch.randn(batch, H, H, cin, device=dev).bfloat16()
    for _ in range(2):
        ops.conv2d_nhwc(xb, w, b, relu=True, pool=layer_pool, force_r=force_r)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv2d_nhwc(xb, w, b, relu=True, pool=layer_pool, force_r=force_r)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    res["ms_median"] = ts[len(ts) // 2]
    res["ms_min"] = ts[0]
    res["tflops"] = 2.0 * batch * H * H * cout * 9 * cin / (ts[len(ts) // 2] * 1e-3) / 1e12
    res["ok"] = res["bad_frac_pool1"] == 0.0 and res["bad_frac_pool0"] == 0.0
    return res


def run_fused(kind, batch):
    """Event-timed: K1 + conv1_1 (unfused) vs the fused gather kernel, protocol tables of video 0 (250 snippets)."""
    import torch
    from video_analytics_b200 import ops
    from video_analytics_b200.evaluate import spatial_table, temporal_table
    from video_analytics_b200.store import DeviceStore, make_layout
    lay = make_layout(4)
    store = DeviceStore(lay)
    m = lay.videos[0]
    if kind == "s":
        images, shape, tab, mean, std, c_pad = store.rgb, lay.rgb_shape, spatial_table(m, lay.rgb_shape), [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], 16
    else:
        images, shape, tab, mean, std, c_pad = store.flow, lay.flow_shape, temporal_table(m, lay.flow_shape), [0.485] * 20, [0.229] * 20, 32
    import numpy as np
    reps = (batch + tab.shape[0] - 1) // tab.shape[0]
    table = torch.from_numpy(np.concatenate([tab] * reps)[:batch]).cuda()
    cin = table.shape[1] * shape[2]
    g = torch.Generator().manual_seed(4)
    w = (torch.randn(64, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
    b = (torch.randn(64, generator=g) * 0.1).cuda()

    def unfused():
        x = ops.preprocess(images, shape, table, mean, std, c_pad=c_pad)
        return ops.conv2d_nhwc(x, w, b, relu=True, pool=False)

    def fused():
        return ops.conv1_fused(images, shape, table, mean, std, w, b)

    res = dict(kind=kind, batch=batch)
    y2 = unfused().float()
    y1 = fused().float()
    torch.cuda.synchronize()
    res["max_abs_diff"] = float((y1 - y2).abs().max())
    res["frac_differs"] = float((y1 != y2).float().mean())
    res["ref_absmean"] = float(y2.abs().mean())
    del y1, y2
    for name, fn in (("unfused_ms", unfused), ("fused_ms", fused)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        res[name] = ts[len(ts) // 2]
    # per-role cycle counters of CTA 0 (gather thread 0, MMA warp, epilogue group 0)
    from video_analytics_b200 import _lib
    cnt = torch.zeros(16, dtype=torch.int64, device="cuda")
    _lib.load().va_debug_conv_counters(_lib.ptr(cnt))
    fused()
    torch.cuda.synchronize()
    _lib.load().va_debug_conv_counters(None)
    c = cnt.cpu().tolist()
    tiles = max(1, c[11])
    res["roles"] = {"tiles_per_cta": tiles, "clk_per_tile": c[0] / tiles,
                    "gather_pct": {"wait_strip": 100 * c[1] / max(1, c[0]), "wait_empty": 100 * c[2] / max(1, c[0]), "convert": 100 * c[3] / max(1, c[0]),
                                   "proxy_fence": 100 * c[12] / max(1, c[0])},
                    "mma_pct": {"wait_tmem_empty": 100 * c[5] / max(1, c[4]), "wait_full": 100 * c[6] / max(1, c[4])},
                    "epilogue0_pct": {"wait_tmem_full": 100 * c[8] / max(1, c[7]), "wait_staging": 100 * c[9] / max(1, c[7])}}
    out_bytes = batch * 224 * 224 * 64 * 2
    res["fused_write_GBs"] = out_bytes / (res["fused_ms"] * 1e-3) / 1e9
    res["ok"] = res["max_abs_diff"] <= 0.05 * max(res["ref_absmean"], 1e-3) + 0.05
    return res


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "fused":
        try:
            res = run_fused(sys.argv[2], int(sys.argv[3]))
        except Exception as e:  # noqa
            res = dict(kind=sys.argv[2], ok=False, error=f"{type(e).__name__}: {e}"[:600])
        print("RESULT " + json.dumps(res), flush=True)
        return
    if len(sys.argv) >= 4 and sys.argv[1] == "case":
        try:
            res = run_case(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else "conv1_2")
        except Exception as e:  # noqa
            res = dict(force_r=int(sys.argv[2]), ok=False, error=f"{type(e).__name__}: {e}"[:600])
        print("RESULT " + json.dumps(res), flush=True)
        return
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 250
    for layer in LAYERS:
      for fr in VARIANTS[layer]:
        t0 = time.time()
        try:
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "case", str(fr), str(batch), layer], capture_output=True,
                                text=True, timeout=200)
            lines = [l for l in pr.stdout.splitlines() if l.startswith("RESULT ")]
            print(f"[{layer} force_r={fr}] rc={pr.returncode} {time.time()-t0:.1f}s {lines[-1] if lines else 'NO RESULT ' + pr.stdout[-300:] + pr.stderr[-500:]}",
                  flush=True)
        except subprocess.TimeoutExpired:
            print(f"[{layer} force_r={fr}] TIMEOUT", flush=True)
    for kind in ("s", "t"):
        t0 = time.time()
        try:
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "fused", kind, str(batch)], capture_output=True,
                                text=True, timeout=200)
            lines = [l for l in pr.stdout.splitlines() if l.startswith("RESULT ")]
            print(f"[fused {kind}] rc={pr.returncode} {time.time()-t0:.1f}s {lines[-1] if lines else 'NO RESULT ' + pr.stdout[-300:] + pr.stderr[-500:]}",
                  flush=True)
        except subprocess.TimeoutExpired:
            print(f"[fused {kind}] TIMEOUT", flush=True)


if __name__ == "__main__":
    main()

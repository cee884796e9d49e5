"""JPEG decode throughput: CUDA decoder (CUDA events, per kernel via va_launch? -> whole call) vs Pillow on the host cores."""
import concurrent.futures as cf
import io
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import cv2  # noqa: E402
from PIL import Image  # noqa: E402
from video_analytics_b200 import jpeg  # noqa: E402
from video_analytics_b200.store import DeviceStore, make_layout  # noqa: E402

layout = make_layout(2)
store = DeviceStore(layout, torch.device("cuda"))
rgb = store.rgb.cpu().numpy().reshape(-1, *layout.rgb_shape)
flow = store.flow.cpu().numpy().reshape(-1, layout.flow_shape[0], layout.flow_shape[1])
n_rgb, n_flow = int(sys.argv[1]) if len(sys.argv) > 1 else 50, int(sys.argv[2]) if len(sys.argv) > 2 else 1000
if len(sys.argv) > 3 and sys.argv[3] == "smooth":
    # photo-like content (low-frequency structure + mild sensor noise): most blocks end with EOB, files are ~3x smaller
    rng = np.random.default_rng(0)

    def smooth(shape, k):
        h, w = shape[0], shape[1]
        yy, xx = np.mgrid[0:h, 0:w]
        base = [128 + 70 * np.sin(xx / (23.0 + k % 7) + k) * np.cos(yy / (31.0 + k % 5)), 120 + 60 * np.cos((xx + yy) / (41.0 + k % 3)),
                110 + 50 * np.sin(yy / 19.0 + 0.3 * k)]
        img = np.stack(base[:shape[2]] if len(shape) == 3 else base[:1], -1) + rng.normal(0, 3.0, (h, w, shape[2] if len(shape) == 3 else 1))
        img = img.clip(0, 255).astype(np.uint8)
        return img if len(shape) == 3 else img[..., 0]
    rgb = np.stack([smooth(layout.rgb_shape, k) for k in range(16)])
    flow = np.stack([smooth((layout.flow_shape[0], layout.flow_shape[1]), k) for k in range(32)])
rgb_files = [cv2.imencode(".jpg", rgb[i % len(rgb)][..., ::-1])[1].tobytes() for i in range(n_rgb)]
flow_files = [cv2.imencode(".jpg", flow[i % len(flow)])[1].tobytes() for i in range(n_flow)]
print(f"{n_rgb} RGB files avg {np.mean([len(f) for f in rgb_files]):.0f} B, {n_flow} flow files avg {np.mean([len(f) for f in flow_files]):.0f} B")
for name, files, shape in (("rgb", rgb_files, layout.rgb_shape), ("flow", flow_files, (layout.flow_shape[0], layout.flow_shape[1], 1))):
    nbytes = int(np.prod(shape))
    out = torch.empty(len(files) * nbytes, dtype=torch.uint8, device="cuda")
    offs = [k * nbytes for k in range(len(files))]
    t1 = time.perf_counter()
    fs = jpeg.JpegFileSet(files)                      # stage once: header parse + copy into pinned memory
    t_parse = time.perf_counter() - t1
    for _ in range(2):
        fs.decode_into(out, offs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        fs.decode_into(out, offs)                     # H2D of the compressed bytes + the three kernels
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    ref = np.asarray(Image.open(io.BytesIO(files[0])))
    got = out[:nbytes].cpu().numpy().reshape(ref.shape)
    assert np.array_equal(got, ref)
    print(f"{name}: {len(files)} images/call, {dt * 1e3:.2f} ms/call (H2D of the files + decode; one-time staging {t_parse * 1e3:.2f} ms) -> "
          f"{len(files) / dt:.0f} images/s, {len(files) * nbytes / dt / 1e9:.2f} GB/s decoded")

    def pil_decode(f):
        return np.asarray(Image.open(io.BytesIO(f))).shape
    threads = os.cpu_count()
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(pil_decode, files[:64]))
        t0 = time.perf_counter()
        list(ex.map(pil_decode, files))
        dtc = time.perf_counter() - t0
    print(f"   Pillow on {threads} host threads: {len(files) / dtc:.0f} images/s")

"""TEST INFRASTRUCTURE (oracle side) -- synthetic UCF101-shaped video store, numpy restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The pixel function is a pure-uint32 counter hash, identical to `synth_fill_kernel` in
video_analytics_b200/csrc/va_small_kernels.cu, so the CPU oracle and the GPU path see byte-identical frames
(SURVEY.md section 8d).  Shapes follow the reference's data: RGB frames 240x320x3 (UCF101 native; the reference
never resizes, utils.py:116-120) and single-channel flow images 256x340 (TSN tvl1 convention for the directory
named at parameters.py:27 -- a documented assumption, kept a parameter).
"""
from __future__ import annotations

import numpy as np

RGB_SHAPE = (240, 320, 3)
FLOW_SHAPE = (256, 340, 1)
STORE_SEED = 1234


def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def synth_image(seed: int, image_id: int, shape) -> np.ndarray:
    """u8 image [H, W, C] for a global image id; same arithmetic as the CUDA generator."""
    H, W, Cn = shape
    with np.errstate(over="ignore"):
        idx = np.arange(H * W * Cn, dtype=np.uint32)
        c = idx % np.uint32(Cn)
        x = (idx // np.uint32(Cn)) % np.uint32(W)
        y = idx // np.uint32(Cn * W)
        idv = np.uint32(image_id & 0xFFFFFFFF)
        k = _mix32(np.array([idv * np.uint32(0x9E3779B1) + np.uint32(0x85EBCA6B)], dtype=np.uint32))[0]
        h = _mix32(np.uint32(seed & 0xFFFFFFFF) ^ k ^ (idx * np.uint32(0xC2B2AE35)))
        smooth = (x * np.uint32(3) + y * np.uint32(5) + c * np.uint32(41) + idv * np.uint32(29)) >> np.uint32(1)
        val = (smooth + (h & np.uint32(63))) & np.uint32(255)
    return val.astype(np.uint8).reshape(H, W, Cn)


def build_store_numpy(lay):
    """`lay` is a video_analytics_b200.store.StoreLayout (duck-typed: seed, rgb_shape, flow_shape, n_rgb_images,
    n_flow_images).  Returns (rgb u8 [n_rgb, H, W, 3], flow u8 [n_flow, Hf, Wf, 1]) -- CPU copy of the device store.
    RGB image ids hash with `seed`, flow image ids with `seed + 1` (both id-spaces start at 0)."""
    rgb = np.stack([synth_image(lay.seed, i, lay.rgb_shape) for i in range(lay.n_rgb_images)]) if lay.n_rgb_images else \
        np.zeros((0,) + tuple(lay.rgb_shape), np.uint8)
    flow = np.stack([synth_image(lay.seed + 1, i, lay.flow_shape) for i in range(lay.n_flow_images)]) if lay.n_flow_images else \
        np.zeros((0,) + tuple(lay.flow_shape), np.uint8)
    return rgb, flow

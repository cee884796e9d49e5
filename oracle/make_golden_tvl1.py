"""Writes tests/golden/tvl1_small.npz: inputs and oracle outputs (oracle/tvl1.py) of two small frame pairs, plus the
pinned sub-steps that CAN be checked against a real implementation in this container: cv2.cvtColor grey values.
Run from the repo root: python oracle/make_golden_tvl1.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tvl1  # noqa: E402
from video_analytics_b200.flow import synthetic_clip  # noqa: E402


def main():
    out = {}
    clip = synthetic_clip(3, 48, 64, seed=11, velocity=(1.2, -0.6), object_velocity=(-1.5, 0.8))
    out["clip_a"] = clip
    out["gray_a_cv2"] = np.stack([cv2.cvtColor(f, cv2.COLOR_RGB2GRAY) for f in clip])
    p = tvl1.TVL1Params()
    for k in range(2):
        u1, u2, st = tvl1.tvl1_flow(tvl1.gray_from_rgb(clip[k]), tvl1.gray_from_rgb(clip[k + 1]), p, return_stats=True)
        out[f"a{k}_u1"], out[f"a{k}_u2"], out[f"a{k}_iters"] = u1, u2, np.array(st, np.int32)
        out[f"a{k}_x"], out[f"a{k}_y"] = tvl1.flow_to_u8(u1, p.bound), tvl1.flow_to_u8(u2, p.bound)
    clip = synthetic_clip(2, 77, 100, seed=12, channels=1, velocity=(-2.6, 1.9), object_velocity=(3.0, 0.5), noise=1)
    out["clip_b"] = clip
    u1, u2, st = tvl1.tvl1_flow(clip[0, :, :, 0], clip[1, :, :, 0], p, return_stats=True)
    out["b0_u1"], out["b0_u2"], out["b0_iters"] = u1, u2, np.array(st, np.int32)
    out["b0_x"], out["b0_y"] = tvl1.flow_to_u8(u1, p.bound), tvl1.flow_to_u8(u2, p.bound)
    path = os.path.join(ROOT, "tests", "golden", "tvl1_small.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()

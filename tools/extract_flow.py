"""dense_flow-style command line on va_tvl1_flow: one video, or a UCF101-style list, to flow_x_/flow_y_ JPEG directories.

  python tools/extract_flow.py --video v_Archery_g01_c01.avi --out flow/Archery/v_Archery_g01_c01
  python tools/extract_flow.py --root UCF101 --list demoTrain.txt --save mini-ucf101_flow_img_tvl1_gpu [--mode train]

Defaults follow TSN's tool: frames resized to 340 x 256, TV-L1 with OpenCV's CUDA defaults, bound 20, step 1."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analytics_b200 import flow  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--video")
    ap.add_argument("--out")
    ap.add_argument("--root")
    ap.add_argument("--list")
    ap.add_argument("--save")
    ap.add_argument("--mode", default="train")
    ap.add_argument("--bound", type=float, default=20.0)
    ap.add_argument("--new-width", type=int, default=340)
    ap.add_argument("--new-height", type=int, default=256)
    ap.add_argument("--step", type=int, default=1)
    a = ap.parse_args()
    size = (a.new_width, a.new_height) if a.new_width > 0 and a.new_height > 0 else None
    p = flow.TVL1Params(bound=a.bound, new_size=size)
    t0 = time.time()
    if a.video:
        fx, _ = flow.extract_video_flow(a.video, a.out, params=p, step=a.step)
        print(f"{a.video}: {fx.shape[0]} flow pairs of {tuple(fx.shape[1:])} in {time.time() - t0:.2f} s -> {a.out}")
    else:
        n = flow.convertVideosToFlow(a.root, a.save, a.list, mode=a.mode, params=p)
        print(f"{n} videos in {time.time() - t0:.2f} s -> {a.save}")


if __name__ == "__main__":
    main()

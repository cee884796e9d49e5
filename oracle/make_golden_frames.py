"""Pins the frame-extraction step (second half) against the REFERENCE's own code: imports
/root/reference/Sheet03/utils.py (it parses under Python 3), runs its convertVideosToFrames (utils.py:95-121, every 10th
frame -> <i>.jpg through cv2.VideoCapture / cv2.imwrite) on a small synthetic video, and records the names and SHA-256 of
the files it wrote.  The video itself is committed (tests/golden/frames_video.avi) so that the test can run the package's
own utils.convertVideosToFrames on the same bytes, on any box.  Run from the repo root: python oracle/make_golden_frames.py"""
import hashlib
import importlib.util
import json
import os
import shutil
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/Sheet03"


def load_reference_utils():
    sys.path.insert(0, REF)                      # utils.py does `from parameters import *`
    try:
        spec = importlib.util.spec_from_file_location("ref_utils", os.path.join(REF, "utils.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REF)
        sys.modules.pop("parameters", None)
    return mod


def main():
    from video_analytics_b200.flow import synthetic_clip
    golden = os.path.join(ROOT, "tests", "golden")
    video = os.path.join(golden, "frames_video.avi")
    clip = synthetic_clip(23, 48, 64, seed=91)
    vw = cv2.VideoWriter(video, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (64, 48))
    assert vw.isOpened()
    for f in clip:
        vw.write(np.ascontiguousarray(f[..., ::-1]))
    vw.release()
    ref = load_reference_utils()
    tmp = tempfile.mkdtemp()
    try:
        root = os.path.join(tmp, "videos")
        os.makedirs(os.path.join(root, "Archery"))
        shutil.copy(video, os.path.join(root, "Archery", "v_Archery_g01_c01.avi"))
        lst = os.path.join(tmp, "list.txt")
        with open(lst, "w") as f:
            f.write("Archery/v_Archery_g01_c01.avi 1\n")
        save = os.path.join(tmp, "frames")
        ref.convertVideosToFrames(root, save, lst, ref.VIDEO_FRAME_SAMPLE_RATE, "train")
        out_dir = os.path.join(save, "Archery", "v_Archery_g01_c01")
        files = {n: hashlib.sha256(open(os.path.join(out_dir, n), "rb").read()).hexdigest() for n in sorted(os.listdir(out_dir))}
        kept = ref.extractEveryNthFrame(os.path.join(root, "Archery", "v_Archery_g01_c01.avi"), 7)
        manifest = {"video": "frames_video.avi", "video_sha256": hashlib.sha256(open(video, "rb").read()).hexdigest(),
                    "frames_in_video": 23, "sample_rate": int(ref.VIDEO_FRAME_SAMPLE_RATE), "reference_files": files,
                    "every_7th": {"count": len(kept), "sha256": [hashlib.sha256(np.ascontiguousarray(k).tobytes()).hexdigest() for k in kept]},
                    "cv2_version": cv2.__version__, "source": "Sheet03/utils.py:51-69,95-121 executed from /root/reference"}
    finally:
        shutil.rmtree(tmp)
    with open(os.path.join(golden, "frames_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(json.dumps(manifest, indent=1)[:600])


if __name__ == "__main__":
    main()

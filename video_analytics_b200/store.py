"""Synthetic UCF101-shaped video store resident in HBM (bench / test data; there is no network for datasets).

Replaces the on-disk frame folders the reference reads with os.listdir + PIL.Image.open
(Sheet03/spatialModel.py:72-77, temporalModel.py:76-86): per video, `n_frames` RGB frames named 0.jpg..n-1.jpg
(every 10th video frame, utils.py:65) and `n_flows` x/y flow images flow_x_0001.. / flow_y_0001..
(parameters.py:38-39).  Decoded images live in two flat u8 device buffers addressed by a global image id --
the id is what the preprocess kernel's index table carries.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

VIDEO_FRAME_SAMPLE_RATE = 10  # parameters.py:2: frames are extracted every 10th video frame
RGB_SHAPE = (240, 320, 3)     # UCF101 native; the reference never resizes (utils.py:116-120)
FLOW_SHAPE = (256, 340, 1)    # TSN tvl1 tool convention for parameters.py:27's directory -- an assumption, a parameter
STORE_SEED = 1234


def _mix32(x: int) -> int:
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


@dataclass
class VideoMeta:
    """Where one video's images live in the flat stores (ids are global image indices)."""
    name: str
    category: str
    label: int          # 1-based, as in demoTrain.txt
    n_frames: int
    rgb_first: int      # image id of frame "0.jpg" in the RGB store
    n_flows: int
    flowx_first: int    # image id of flow_x_0001 in the flow store
    flowy_first: int    # image id of flow_y_0001


@dataclass
class StoreLayout:
    """Pool of P distinct synthetic videos; video v of a larger job maps to pool entry v % P."""
    videos: List[VideoMeta] = field(default_factory=list)
    n_rgb_images: int = 0
    n_flow_images: int = 0
    rgb_shape: tuple = RGB_SHAPE
    flow_shape: tuple = FLOW_SHAPE
    seed: int = STORE_SEED

    def video(self, v: int) -> VideoMeta:
        return self.videos[v % len(self.videos)]

    def list_line(self, v: int, mode: str = "train") -> str:
        """A line in the format of demoTrain.txt / demoTest.txt (parsed by utils.videoInfo)."""
        m = self.video(v)
        base = f"{m.category}/{m.name}.avi"
        return f"{base} {m.label}\n" if mode == "train" else base + "\n"


def make_layout(pool: int, *, seed: int = STORE_SEED, min_frames: int = 25, frame_span: int = 19, n_classes: int = 25,
                flows_per_frame: int = VIDEO_FRAME_SAMPLE_RATE, rgb_shape=RGB_SHAPE, flow_shape=FLOW_SHAPE) -> StoreLayout:
    """UCF101-shaped clips: a video of 250..430 frames has n_frames_v = min_frames + hash(seed, v) %
    frame_span STORED frames -- the reference keeps every 10th frame (utils.py:65, parameters.py:2) -- and a flow image
    pair per video frame, n_flows_v = flows_per_frame * n_frames_v.  With >= 25 stored frames the protocol's 25 equally
    spaced snippets are 25 DISTINCT frames and 25 disjoint stacks of 10 flow pairs."""
    lay = StoreLayout(seed=seed, rgb_shape=tuple(rgb_shape), flow_shape=tuple(flow_shape))
    rgb = flow = 0
    for v in range(pool):
        hv = _mix32(seed * 0x9E3779B1 + v * 0x85EBCA6B + 0x27D4EB2F)
        nf = min_frames + hv % frame_span
        nfl = flows_per_frame * nf
        label = 1 + v % n_classes
        cat = f"Class{label:03d}"
        name = f"v_{cat}_g{1 + (v // n_classes) % 25:02d}_c{1 + v % 7:02d}"
        lay.videos.append(VideoMeta(name, cat, label, nf, rgb, nfl, flow, flow + nfl))
        rgb += nf
        flow += 2 * nfl
    lay.n_rgb_images, lay.n_flow_images = rgb, flow
    return lay


class DeviceStore:
    """The two flat u8 image buffers in HBM, generated on the device by va_synth_fill
    (same integer hash as oracle/synth.py, so the CPU oracle can rebuild identical bytes)."""

    def __init__(self, layout: StoreLayout, device=None):
        import torch
        from . import ops

        self.layout = layout
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        h, w, c = layout.rgb_shape
        hf, wf, cf = layout.flow_shape
        self.rgb = torch.empty(max(1, layout.n_rgb_images) * h * w * c, dtype=torch.uint8, device=dev)
        self.flow = torch.empty(max(1, layout.n_flow_images) * hf * wf * cf, dtype=torch.uint8, device=dev)
        if layout.n_rgb_images:
            ops.synth_fill(self.rgb, layout.rgb_shape, layout.n_rgb_images, seed=layout.seed)
        if layout.n_flow_images:
            ops.synth_fill(self.flow, layout.flow_shape, layout.n_flow_images, seed=layout.seed + 1)

    @classmethod
    def from_jpeg_files(cls, layout: StoreLayout, rgb_files, flow_files, device=None):
        """Build the store from the reference's on-disk format: `rgb_files[k]` / `flow_files[k]` are the bytes of the JPEG
        of image id k (frames "<i>.jpg" written by cv2.imwrite, utils.py:116-120; flow_x_/flow_y_ images).  Decoded on the
        GPU by jpeg.decode_into -- the pixels are Pillow's (Image.open, spatialModel.py:76-79), bit for bit."""
        import torch
        from . import jpeg

        self = cls.__new__(cls)
        self.layout = layout
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        nb_rgb = layout.rgb_shape[0] * layout.rgb_shape[1] * layout.rgb_shape[2]
        nb_flow = layout.flow_shape[0] * layout.flow_shape[1] * layout.flow_shape[2]
        self.rgb = torch.empty(max(1, len(rgb_files)) * nb_rgb, dtype=torch.uint8, device=dev)
        self.flow = torch.empty(max(1, len(flow_files)) * nb_flow, dtype=torch.uint8, device=dev)
        for files, buf, nb, shape in ((rgb_files, self.rgb, nb_rgb, layout.rgb_shape), (flow_files, self.flow, nb_flow, layout.flow_shape)):
            if len(files) == 0:
                continue
            fs = jpeg.JpegFileSet(files)
            want = (shape[0], shape[1], shape[2]) if shape[2] != 1 else (shape[0], shape[1], 1)
            bad = [k for k, sz in enumerate(fs.sizes) if tuple(sz) != want]
            if bad:
                raise ValueError("image %d is %r, the store layout expects %r" % (bad[0], fs.sizes[bad[0]], want))
            fs.decode_into(buf, [k * nb for k in range(len(files))])
        return self

    @classmethod
    def from_directories(cls, video_lines, mode: str, frames_root=None, flow_root=None, label_of=None, device=None):
        """Build the store from the reference's ON-DISK TREE, the way its datasets walk it:
        `<frames_root>/<Category>/<video>/<i>.jpg`, i = 0..n-1 with n = len(os.listdir(dir)) (spatialModel.py:72-77; the
        files convertVideosToFrames writes, utils.py:95-121) and `<flow_root>/<Category>/<video>/flow_x_%04d.jpg` /
        `flow_y_%04d.jpg`, 1..len(os.listdir(dir))/2 (temporalModel.py:76-81, parameters.py:38-39).  `video_lines` are the
        lines of demoTrain.txt / demoTest.txt; either root may be None (that stream's store stays empty).  All JPEGs are
        decoded ONCE on the GPU (va_jpeg_decode; pixels = Pillow's) into HBM; every frame of a tree must have the same
        size (UCF101: 320x240 frames, 340x256 flow images) -- a mixed tree raises ValueError.
        `label_of(category) -> int` supplies labels for test-mode lines (the reference's actionLabelDict)."""
        import os

        import torch

        from . import jpeg
        from .utils import videoInfo

        self = cls.__new__(cls)
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        lay = StoreLayout(videos=[], seed=0)
        rgb_files, flow_files, seen = [], [], set()
        for line in video_lines:
            if not line.strip():
                continue
            _, name, label, category, _, _ = videoInfo(line, mode)
            if name in seen:
                continue
            seen.add(name)
            if label is None:
                label = label_of(category) if label_of is not None else 0
            n_frames = n_flows = 0
            rgb_first, fx_first = len(rgb_files), len(flow_files)
            if frames_root is not None:
                d = os.path.join(frames_root, category, name)
                n_frames = len(os.listdir(d))                                     # reference spatialModel.py:74
                for i in range(n_frames):
                    with open(os.path.join(d, "%d.jpg" % i), "rb") as f:
                        rgb_files.append(f.read())
            if flow_root is not None:
                d = os.path.join(flow_root, category, name)
                n_flows = len(os.listdir(d)) // 2                                 # reference temporalModel.py:78
                for prefix in ("flow_x_", "flow_y_"):
                    for i in range(1, n_flows + 1):
                        with open(os.path.join(d, "%s%04d.jpg" % (prefix, i)), "rb") as f:
                            flow_files.append(f.read())
            lay.videos.append(VideoMeta(name, category, int(label), n_frames, rgb_first, n_flows, fx_first, fx_first + n_flows))
        lay.n_rgb_images, lay.n_flow_images = len(rgb_files), len(flow_files)
        self.layout = lay
        self.rgb = torch.empty(1, dtype=torch.uint8, device=dev)
        self.flow = torch.empty(1, dtype=torch.uint8, device=dev)
        for files, attr, comps in ((rgb_files, "rgb", 3), (flow_files, "flow", 1)):
            if not files:
                continue
            fs = jpeg.JpegFileSet(files)
            h, w, c = (int(v) for v in fs.sizes[0])
            if c != comps:
                raise ValueError("%s tree: expected %d-component JPEGs, file 0 has %d" % (attr, comps, c))
            bad = [k for k, sz in enumerate(fs.sizes) if tuple(int(v) for v in sz) != (h, w, c)]
            if bad:
                raise ValueError("%s tree: image %d is %r, image 0 is %r (one size per tree)" % (attr, bad[0], tuple(fs.sizes[bad[0]]), (h, w, c)))
            nb = h * w * c
            buf = torch.empty(len(files) * nb, dtype=torch.uint8, device=dev)
            fs.decode_into(buf, [k * nb for k in range(len(files))])
            setattr(self, attr, buf)
            if attr == "rgb":
                lay.rgb_shape = (h, w, c)
            else:
                lay.flow_shape = (h, w, c)
        torch.cuda.current_stream().synchronize()
        return self

    @classmethod
    def from_videos(cls, video_lines, mode: str, videos_root: str, label_of=None, device=None, flow_params=None,
                    sample_rate: int = VIDEO_FRAME_SAMPLE_RATE):
        """Build BOTH stores straight from the video files, with no image files in between: what the reference prepares
        offline in two steps -- convertVideosToFrames (every `sample_rate`-th frame -> `<i>.jpg`, utils.py:95-121) and the
        third-party TV-L1 tool that writes `flow_x_/flow_y_` (parameters.py:27) -- done here as host decode
        (cv2.VideoCapture, like utils.py:51-69) + va_tvl1_flow on the GPU (flow.py; frames resized to 340 x 256 first, as the
        tool does).  Video v then has ceil(N / sample_rate) stored frames and N - 1 flow pairs, exactly the counts the
        datasets would find in the reference's trees.  All videos must share one frame size."""
        import os

        import torch

        from . import flow as _flow
        from .utils import videoInfo

        self = cls.__new__(cls)
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        fp = flow_params if flow_params is not None else _flow.TVL1Params(new_size=(FLOW_SHAPE[1], FLOW_SHAPE[0]))
        lay = StoreLayout(videos=[], seed=0)
        clips, seen = [], set()
        rgb = flw = 0
        for line in video_lines:
            if not line.strip():
                continue
            loc, name, label, category, _, _ = videoInfo(line, mode)
            if name in seen:
                continue
            seen.add(name)
            if label is None:
                label = label_of(category) if label_of is not None else 0
            frames = _flow.read_video_frames(os.path.join(videos_root, loc))
            if clips and frames.shape[1:] != clips[0].shape[1:]:
                raise ValueError("%s is %r, the first video is %r (one frame size per store)" % (loc, frames.shape[1:], clips[0].shape[1:]))
            n = frames.shape[0]
            n_frames = (n + sample_rate - 1) // sample_rate                       # frames 0, N, 2N, ... (utils.py:65)
            n_flows = max(0, n - 1)
            lay.videos.append(VideoMeta(name, category, int(label), n_frames, rgb, n_flows, flw, flw + n_flows))
            rgb += n_frames
            flw += 2 * n_flows
            clips.append(frames)
        if not clips:
            raise ValueError("from_videos: no videos in the list")
        h, w, c = clips[0].shape[1:]
        oh, ow = fp.out_shape(h, w)
        lay.n_rgb_images, lay.n_flow_images = rgb, flw
        lay.rgb_shape, lay.flow_shape = (h, w, c), (oh, ow, 1)
        self.layout = lay
        self.rgb = torch.empty(max(1, rgb) * h * w * c, dtype=torch.uint8, device=dev)
        self.flow = torch.empty(max(1, flw) * oh * ow, dtype=torch.uint8, device=dev)
        rgb_v = self.rgb.view(-1, h, w, c)
        for k, frames in enumerate(clips):
            m = lay.videos[k]
            fr = torch.from_numpy(frames).to(dev)
            rgb_v[m.rgb_first:m.rgb_first + m.n_frames].copy_(fr[::sample_rate])
            if m.n_flows:
                _flow.fill_flow_store(self, k, fr, params=fp)
        torch.cuda.current_stream().synchronize()
        return self

    @classmethod
    def from_host(cls, layout: StoreLayout, rgb_u8, flow_u8, device=None):
        """Upload host-decoded frames (numpy u8) instead of generating them."""
        import torch

        self = cls.__new__(cls)
        self.layout = layout
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.rgb = torch.from_numpy(rgb_u8).reshape(-1).to(dev)
        self.flow = torch.from_numpy(flow_u8).reshape(-1).to(dev)
        return self

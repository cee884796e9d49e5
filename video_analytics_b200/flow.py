"""TV-L1 optical-flow production on the GPU: the step UPSTREAM of the reference's temporal stream.

The reference opens `<FLOW_DATA_DIR>/<Category>/<video>/flow_x_%04d.jpg` / `flow_y_%04d.jpg` (Sheet03/parameters.py:27,38-39;
temporalModel.py:76-86) and never computes them; the directory name (`..._flow_img_tvl1_gpu`) is that of TSN's
`dense_flow` tool: for consecutive video frames t-1, t (t = 1..N-1), grey -> TV-L1 -> 8-bit images with bound 20, image
number t.  This module is that tool's job on `va_tvl1_flow` (csrc/va_tvl1.cu; arithmetic contract oracle/tvl1.py):

    flow_images(frames)              frames u8 [N, H, W, 3|1] on the device -> (flow_x u8 [N-1, H, W], flow_y ...)
    fill_flow_store(store, ...)      writes a video's flow images into a DeviceStore's flow buffer (the layout the
                                     temporal stream's index tables address), no JPEG round trip
    write_flow_tree(dir, fx, fy)     the reference's on-disk format (flow_x_0001.jpg ...), for interoperability

There is no CPU fallback: without the CUDA library every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import VAError, check, ptr, stream_ptr


class _CParams(C.Structure):
    _fields_ = [("tau", C.c_double), ("lambda_", C.c_double), ("theta", C.c_double), ("epsilon", C.c_double),
                ("scale_step", C.c_double), ("bound", C.c_double), ("nscales", C.c_int), ("warps", C.c_int),
                ("iterations", C.c_int), ("resize_w", C.c_int), ("resize_h", C.c_int), ("reserved", C.c_int)]


@dataclass
class TVL1Params:
    """OpenCV `OpticalFlowDual_TVL1` (CUDA) defaults + dense_flow's bound."""
    tau: float = 0.25
    lambda_: float = 0.15
    theta: float = 0.3
    nscales: int = 5
    warps: int = 5
    epsilon: float = 0.01
    iterations: int = 300
    scale_step: float = 0.8
    bound: float = 20.0
    new_size: Optional[Tuple[int, int]] = None       # (width, height): dense_flow's frame resize before the flow (TSN: (340, 256))

    def _c(self) -> _CParams:
        rw, rh = self.new_size if self.new_size is not None else (0, 0)
        return _CParams(self.tau, self.lambda_, self.theta, self.epsilon, self.scale_step, self.bound, self.nscales,
                        self.warps, self.iterations, int(rw), int(rh), 0)

    def out_shape(self, h: int, w: int) -> Tuple[int, int]:
        return (int(self.new_size[1]), int(self.new_size[0])) if self.new_size is not None else (h, w)

    def levels(self, h: int, w: int) -> int:
        n = 1
        for _ in range(1, self.nscales):
            h, w = int(round(h * self.scale_step)), int(round(w * self.scale_step))
            if h < 16 or w < 16:
                break
            n += 1
        return n


_workspaces = {}


def _workspace(h: int, w: int, cp: _CParams, device) -> torch.Tensor:
    nbytes = int(_lib.load().va_tvl1_workspace_bytes(h, w, C.byref(cp)))
    if nbytes == 0:
        raise VAError("va_tvl1_workspace_bytes: bad image size / parameters")
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def tvl1(images: torch.Tensor, image_shape: Tuple[int, int, int], pair_table: torch.Tensor, out_images: torch.Tensor, *,
         params: Optional[TVL1Params] = None, image_bytes: Optional[int] = None, out_image_bytes: Optional[int] = None,
         return_flow: bool = False, return_iterations: bool = False):
    """va_tvl1_flow on flat stores.  `images` u8 store of [h, w, c] images, `pair_table` int32 [n, 4] (device) = (frame t-1
    id, frame t id, output id of the x image, of the y image), `out_images` u8 store of [h, w] images.  Returns a dict with
    `flow` fp32 [n, 2, h, w] and/or `iterations` int32 [n, levels*warps] when asked."""
    if not (images.is_cuda and pair_table.is_cuda and out_images.is_cuda):
        raise VAError("flow.tvl1: tensors must live on a CUDA device (there is no CPU path)")
    assert images.dtype == torch.uint8 and out_images.dtype == torch.uint8
    assert pair_table.dtype == torch.int32 and pair_table.dim() == 2 and pair_table.shape[1] == 4 and pair_table.is_contiguous()
    h, w, c = image_shape
    p = params or TVL1Params()
    cp = p._c()
    n = int(pair_table.shape[0])
    oh, ow = p.out_shape(h, w)
    ib = image_bytes if image_bytes is not None else h * w * c
    ob = out_image_bytes if out_image_bytes is not None else oh * ow
    res = {}
    flow = torch.empty((n, 2, oh, ow), dtype=torch.float32, device=images.device) if return_flow else None
    its = torch.zeros((n, p.levels(oh, ow) * p.warps), dtype=torch.int32, device=images.device) if return_iterations else None
    wsb = _workspace(h, w, cp, images.device)
    check(_lib.load().va_tvl1_flow(ptr(images), ib, h, w, c, ptr(pair_table), n, C.byref(cp), ptr(out_images), ob, ptr(flow),
                                   ptr(its), ptr(wsb), wsb.numel(), stream_ptr()), "va_tvl1_flow")
    if flow is not None:
        res["flow"] = flow
    if its is not None:
        res["iterations"] = its
    return res


def flow_images(frames: torch.Tensor, *, params: Optional[TVL1Params] = None, step: int = 1, return_flow: bool = False):
    """dense_flow's loop over one video: frames u8 [N, H, W, C] (C = 3 RGB or 1 grey; or [N, H, W]) on the device ->
    (flow_x u8 [N-step, H, W], flow_y u8 [N-step, H, W]); image k is the flow from frame k to frame k+step (the tool
    numbers it k+1: flow_x_0001 = frames 0 -> 1)."""
    if frames.dim() == 3:
        frames = frames.unsqueeze(-1)
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.is_contiguous()
    n, h, w, c = frames.shape
    m = n - step
    if m <= 0:
        raise ValueError("flow_images: need more than `step` frames")
    dev = frames.device
    k = torch.arange(m, dtype=torch.int32, device=dev)
    table = torch.stack([k, k + step, k, k + m], dim=1).contiguous()
    oh, ow = (params or TVL1Params()).out_shape(h, w)
    out = torch.empty((2 * m, oh, ow), dtype=torch.uint8, device=dev)
    res = tvl1(frames, (h, w, c), table, out, params=params, return_flow=return_flow)
    if return_flow:
        return out[:m], out[m:], res["flow"]
    return out[:m], out[m:]


def fill_flow_store(store, video_index: int, frames: torch.Tensor, *, params: Optional[TVL1Params] = None) -> int:
    """Compute the flow images of one video from its consecutive frames and write them where the temporal stream reads
    them: image ids flowx_first + t / flowy_first + t of `store.flow` (store.py::VideoMeta).  The store's flow images
    must have the frames' size.  Returns the number of flow pairs written (min(n_flows, N-1))."""
    lay = store.layout
    m = lay.video(video_index)
    if frames.dim() == 3:
        frames = frames.unsqueeze(-1)
    n, h, w, c = frames.shape
    oh, ow = (params or TVL1Params()).out_shape(h, w)
    if tuple(lay.flow_shape) != (oh, ow, 1):
        raise ValueError(f"fill_flow_store: the flow images come out {oh}x{ow} (frames {h}x{w}, new_size "
                         f"{(params or TVL1Params()).new_size}) but the store's flow images are {lay.flow_shape}")
    cnt = min(m.n_flows, n - 1)
    k = torch.arange(cnt, dtype=torch.int32, device=frames.device)
    table = torch.stack([k, k + 1, k + m.flowx_first, k + m.flowy_first], dim=1).contiguous()
    tvl1(frames.contiguous(), (h, w, c), table, store.flow, params=params)
    return cnt


def write_flow_tree(directory: str, flow_x: torch.Tensor, flow_y: torch.Tensor, *, quality: int = 95) -> None:
    """The reference's on-disk format: `<directory>/flow_x_%04d.jpg`, `flow_y_%04d.jpg`, numbered from 1
    (parameters.py:38-39; TemporalDataset counts len(os.listdir(dir))/2 images, temporalModel.py:76-78)."""
    import cv2

    os.makedirs(directory, exist_ok=True)
    fx, fy = flow_x.cpu().numpy(), flow_y.cpu().numpy()
    for t in range(fx.shape[0]):
        cv2.imwrite(os.path.join(directory, "flow_x_%04d.jpg" % (t + 1)), fx[t], [cv2.IMWRITE_JPEG_QUALITY, quality])
        cv2.imwrite(os.path.join(directory, "flow_y_%04d.jpg" % (t + 1)), fy[t], [cv2.IMWRITE_JPEG_QUALITY, quality])


def synthetic_clip(n_frames: int, h: int, w: int, *, seed: int = 0, channels: int = 3, velocity=(1.7, -0.9),
                   object_velocity=(-2.3, 1.4), noise: int = 2):
    """Bench / test data (numpy, host): a textured background translating by `velocity` pixels per frame with a textured
    disc moving by `object_velocity` over it, plus +-`noise` grey levels of per-frame noise -- analytic functions of the
    pixel coordinate, so sub-pixel motion needs no interpolation.  Returns u8 [n_frames, h, w, channels]."""
    import numpy as np

    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    out = np.empty((n_frames, h, w, channels), np.uint8)
    nk = 24
    freq = rng.uniform(0.02, 0.45, size=(channels, nk, 2)) * rng.choice([-1.0, 1.0], size=(channels, nk, 2))
    amp = rng.uniform(0.3, 1.0, size=(channels, nk)) / np.sqrt(np.abs(freq).sum(-1) + 0.05)
    ph = rng.uniform(0, 2 * np.pi, size=(channels, nk))
    freq_o = rng.uniform(0.1, 0.6, size=(channels, 8, 2))
    ph_o = rng.uniform(0, 2 * np.pi, size=(channels, 8))
    cx0, cy0, rad = w * 0.45, h * 0.5, min(h, w) * 0.18

    def tex(c, x, y, f, a, p):
        v = np.zeros_like(x)
        for k in range(f.shape[1]):
            v += a[c, k] * np.sin(f[c, k, 0] * x + f[c, k, 1] * y + p[c, k])
        return v

    for t in range(n_frames):
        bx, by = xs - velocity[0] * t, ys - velocity[1] * t
        ox, oy = xs - object_velocity[0] * t, ys - object_velocity[1] * t
        inside = (ox - cx0) ** 2 + (oy - cy0) ** 2 < rad * rad
        for c in range(channels):
            bg = tex(c, bx, by, freq, amp, ph)
            ob = tex(c, ox, oy, freq_o, np.full((channels, 8), 0.6), ph_o)
            v = np.where(inside, ob, bg)
            v = 128.0 + 40.0 * v + rng.integers(-noise, noise + 1, size=(h, w))
            out[t, :, :, c] = np.clip(np.rint(v), 0, 255).astype(np.uint8)
    return out


def read_video_frames(video_path: str):
    """All frames of a video as RGB u8 [N, H, W, 3] (host decode with cv2.VideoCapture, like the reference's own frame
    extractor utils.py:51-69; UCF101's .avi files are MPEG-4 ASP, which no NVDEC generation decodes)."""
    import cv2
    import numpy as np

    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise ValueError("Error opening video file %s" % video_path)        # utils.py:56
    frames = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        frames.append(fr[..., ::-1])
    cap.release()
    if not frames:
        raise ValueError("no frames in %s" % video_path)
    return np.ascontiguousarray(np.stack(frames))


def extract_video_flow(video_path: str, out_dir: Optional[str] = None, *, params: Optional[TVL1Params] = None,
                       step: int = 1, chunk: int = 64, quality: int = 95):
    """TSN `dense_flow` (`extract_gpu -f video -x flow_x -y flow_y -b 20 -s 1`, new_size 340 x 256) for one video: every
    frame pair (t - step, t) -> `flow_x_%04d.jpg` / `flow_y_%04d.jpg` numbered from 1 -- the files TemporalDataset lists
    (temporalModel.py:76-81).  Frames are decoded on the host, resized + converted + solved on the GPU in chunks.
    Returns (flow_x, flow_y) u8 device tensors [N - step, H, W]; writes the JPEGs when out_dir is given."""
    p = params if params is not None else TVL1Params(new_size=(340, 256))
    frames = read_video_frames(video_path)
    n = frames.shape[0]
    if n <= step:
        raise ValueError("%s has %d frames, need more than %d" % (video_path, n, step))
    xs, ys = [], []
    for a in range(0, n - step, chunk):
        b = min(n, a + chunk + step)
        fx, fy = flow_images(torch.from_numpy(frames[a:b]).cuda(), params=p, step=step)
        xs.append(fx)
        ys.append(fy)
    fx, fy = torch.cat(xs), torch.cat(ys)
    if out_dir is not None:
        write_flow_tree(out_dir, fx, fy, quality=quality)
    return fx, fy


def convertVideosToFlow(rootDir: str, saveDir: str, videoListLoc: str, mode: str = "train", *,
                        params: Optional[TVL1Params] = None) -> int:
    """The flow counterpart of the reference's convertVideosToFrames (utils.py:95-121): for every line of the video list,
    `<rootDir>/<Category>/<video>.avi` -> `<saveDir>/<Category>/<video>/flow_{x,y}_%04d.jpg` -- the tree FLOW_DATA_DIR
    points at (parameters.py:27).  Returns the number of videos converted."""
    from .utils import videoInfo

    done = 0
    with open(videoListLoc) as f:
        for line in f:
            if not line.strip():
                continue
            loc, name, _, category, _, _ = videoInfo(line, mode)
            extract_video_flow(os.path.join(rootDir, loc), os.path.join(saveDir, category, name), params=params)
            done += 1
    return done

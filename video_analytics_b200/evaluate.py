"""Two-stream video evaluation: the 25 snippets x 10 crops test protocol of notes.txt:113-116 / 225-230 on top of
the reference's per-stream forward, per-video consensus and late fusion.

For every video and stream: 25 equally spaced snippets (frames, or stacks of L flow pairs) x torchvision
`ten_crop` order -> 250 network inputs -> descriptors [250,256] + softmax scores [250,101]; consensus = mean in
AverageMeter order; fusion = [spatial | temporal] descriptor (+ LinearSVC scores) and the weighted class-score
average.  All index arithmetic is done on the host with integers; all pixel/tensor work is in libva_b200.so.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .combinedModel import CombinedModel
from .parameters import (FLOW_NORM_MEAN, FLOW_NORM_STD, N_TEST_CROPS, N_TEST_SNIPPETS, NORM_MEANS_TF, NORM_STDS_TF,
                         VIDEO_INPUT_FLOW_COUNT)
from .store import DeviceStore, VideoMeta
from .utils import ten_crop_params, test_flow_starts, test_frame_indices

SNIPPETS_PER_VIDEO = N_TEST_SNIPPETS * N_TEST_CROPS      # per stream


def spatial_table(meta: VideoMeta, rgb_shape) -> np.ndarray:
    """int32 [250, 1, 4] index table of one video (snippet-major, crop-minor)."""
    crops = ten_crop_params(rgb_shape[0], rgb_shape[1])
    rows = [[meta.rgb_first + f, i, j, fl] for f in test_frame_indices(meta.n_frames) for (i, j, fl) in crops]
    return np.asarray(rows, dtype=np.int32).reshape(-1, 1, 4)


def temporal_table(meta: VideoMeta, flow_shape, L: int = VIDEO_INPUT_FLOW_COUNT) -> np.ndarray:
    """int32 [250, 2L, 4]: one crop per stack, channels x_t, y_t, x_{t+1}, ... (temporalModel.py:80-83)."""
    crops = ten_crop_params(flow_shape[0], flow_shape[1])
    rows = []
    for s in test_flow_starts(meta.n_flows, L):
        for (i, j, fl) in crops:
            for idx in range(s, s + L):
                rows.append([meta.flowx_first + idx - 1, i, j, fl])
                rows.append([meta.flowy_first + idx - 1, i, j, fl])
    return np.asarray(rows, dtype=np.int32).reshape(-1, 2 * L, 4)


class HostStore:
    """The loader side of the end-to-end path: the pool videos' decoded u8 images in PINNED host memory (same flat
    layout and image ids as the DeviceStore they mirror).  `TwoStreamEvaluator.host_pipeline` copies, per video, only the
    images the 25 x 10 protocol reads -- the files the reference would `Image.open` (spatialModel.py:76, temporalModel.py:85)."""

    def __init__(self, layout, rgb: torch.Tensor, flow: torch.Tensor):
        self.layout = layout
        self.rgb = rgb if rgb.is_pinned() else rgb.pin_memory()
        self.flow = flow if flow.is_pinned() else flow.pin_memory()

    @classmethod
    def from_device(cls, store: DeviceStore):
        return cls(store.layout, store.rgb.cpu(), store.flow.cpu())


class _VideoPlan:
    """What one pool video contributes to a staged group: source image runs to copy and index tables whose image ids are
    LOCAL to the video's stage slot (frames 0..24, flow images 0..499)."""

    def __init__(self, meta: VideoMeta, layout, L: int):
        frames = sorted(set(test_frame_indices(meta.n_frames)))
        f_local = {f: i for i, f in enumerate(frames)}
        flows = sorted({s + d for s in test_flow_starts(meta.n_flows, L) for d in range(L)})         # 1-based flow indices
        x_local = {idx: i for i, idx in enumerate(flows)}
        nfl = len(flows)
        self.n_rgb, self.n_flow = len(frames), 2 * nfl
        self.rgb_src = [meta.rgb_first + f for f in frames]                                           # one copy per frame
        runs, start = [], 0                                                                          # consecutive flow indices
        for i in range(1, nfl + 1):
            if i == nfl or flows[i] != flows[i - 1] + 1:
                runs.append((flows[start], start, i - start))                                         # (first index, local, count)
                start = i
        self.flow_runs = [(meta.flowx_first + a - 1, loc, cnt) for a, loc, cnt in runs] + \
                         [(meta.flowy_first + a - 1, nfl + loc, cnt) for a, loc, cnt in runs]
        lm = VideoMeta(meta.name, meta.category, meta.label, meta.n_frames, 0, meta.n_flows, 0, 0)
        ts = spatial_table(lm, layout.rgb_shape)
        ts[:, :, 0] = np.vectorize(f_local.get)(ts[:, :, 0])
        tt = temporal_table(VideoMeta(meta.name, meta.category, meta.label, meta.n_frames, 0, meta.n_flows, 1, 1 + 10 ** 6), layout.flow_shape, L)
        ids = tt[:, :, 0]
        is_y = ids >= 10 ** 6                                                                         # x ids = idx, y ids = 1e6 + idx
        idx = np.where(is_y, ids - 10 ** 6, ids)
        tt[:, :, 0] = np.vectorize(x_local.get)(idx) + np.where(is_y, nfl, 0)
        self.ts, self.tt = ts.astype(np.int32), tt.astype(np.int32)


class TwoStreamEvaluator:
    """Runs groups of videos through preprocess -> both streams -> consensus/fusion on the current device."""

    def __init__(self, spatial: ops.StreamNet, temporal: ops.StreamNet, store: DeviceStore,
                 combined: Optional[CombinedModel] = None, L: int = VIDEO_INPUT_FLOW_COUNT):
        self.spatial, self.temporal, self.store = spatial, temporal, store
        self.combined = combined if combined is not None else CombinedModel()
        self.L = L
        # True (default since the loader/converter version of the kernel): the crops are gathered inside conv1_1
        # (va_forward_store), no preprocessed tensor exists in HBM; False / VA_FUSED_FRONT_END=0: K1 tensor + va_forward
        import os as _os
        self.fused_front_end = _os.environ.get("VA_FUSED_FRONT_END", "1") == "1"
        self._tables: Dict[int, tuple] = {}
        self.mean_s, self.std_s = list(NORM_MEANS_TF), list(NORM_STDS_TF)
        self.mean_t, self.std_t = [FLOW_NORM_MEAN] * (2 * L), [FLOW_NORM_STD] * (2 * L)

    def tables_for(self, v: int):
        """Device index tables of pool video v (cached: the protocol is deterministic per video)."""
        lay = self.store.layout
        k = v % len(lay.videos)
        if k not in self._tables:
            m = lay.videos[k]
            ts = torch.from_numpy(spatial_table(m, lay.rgb_shape)).cuda()
            tt = torch.from_numpy(temporal_table(m, lay.flow_shape, self.L)).cuda()
            self._tables[k] = (ts, tt)
        return self._tables[k]

    def run_videos(self, video_ids: Sequence[int], out: Optional[dict] = None, out_row: int = 0) -> dict:
        """Evaluate a group of videos; returns the fusion result dict for the group (or writes rows
        [out_row, out_row+len) of the preallocated `out` tensors -- e.g. this rank's slice of an all-gather buffer)."""
        V = len(video_ids)
        tabs = [self.tables_for(v) for v in video_ids]
        ts = torch.cat([t[0] for t in tabs]) if V > 1 else tabs[0][0]
        tt = torch.cat([t[1] for t in tabs]) if V > 1 else tabs[0][1]
        return self.run_tables(ts, tt, V, out=out, out_row=out_row)

    def run_tables(self, ts: torch.Tensor, tt: torch.Tensor, V: int, out: Optional[dict] = None, out_row: int = 0,
                   store: Optional[DeviceStore] = None) -> dict:
        """Same, from explicit device index tables ts [V*250,1,4] / tt [V*250,2L,4] (image ids relative to `store`)."""
        store = store if store is not None else self.store
        offs = torch.arange(0, (V + 1) * SNIPPETS_PER_VIDEO, SNIPPETS_PER_VIDEO, dtype=torch.int32, device=ts.device)
        lay = store.layout
        desc_s, prob_s = self._stream(self.spatial, store.rgb, lay.rgb_shape, ts, self.mean_s, self.std_s)
        desc_t, prob_t = self._stream(self.temporal, store.flow, lay.flow_shape, tt, self.mean_t, self.std_t)
        sub = None
        if out is not None:
            sub = {k: t[out_row:out_row + V] for k, t in out.items()}
        return self.combined.fuse(desc_s, desc_t, prob_s, prob_t, offs, out=sub)

    # ---- end-to-end from host memory: double-buffered staging, H2D of group i+1 under the networks of group i
    MAX_RGB_PER_VIDEO = N_TEST_SNIPPETS
    MAX_FLOW_PER_VIDEO = 2 * N_TEST_SNIPPETS * VIDEO_INPUT_FLOW_COUNT

    def host_pipeline(self, host: HostStore, groups: Sequence[Sequence[int]], depth: int = 2, to_host: Optional[Sequence[str]] = None):
        """Generator: evaluates each group of video ids from PINNED HOST images and yields its fusion result dict.
        The images a group touches (25 frames + 2 x 250 flow images per video) and its index tables are copied into one
        of `depth` device stage stores on a copy stream while the previous group's networks run; `self.last_h2d_bytes` is
        the byte count of the most recent group's copies.  Results are stream-ordered on the CURRENT stream.
        to_host = names of result tensors (e.g. ("video_scores", "score_pred")): the pipeline also copies those rows back
        into pinned host buffers and yields HOST tensors, synchronised by an event -- and it queues the NEXT group's
        kernels before it waits for that event, so the GPU never idles while the host handles a group's results (a plain
        stream synchronise per group left it idle for the host's launch time; `self.last_d2h_bytes` = bytes read back)."""
        lay = host.layout
        dev = self.store.rgb.device
        vmax = max((len(g) for g in groups), default=0)
        if vmax == 0:
            return
        rgb_img = lay.rgb_shape[0] * lay.rgb_shape[1] * lay.rgb_shape[2]
        flow_img = lay.flow_shape[0] * lay.flow_shape[1] * lay.flow_shape[2]
        key = (vmax, depth, rgb_img, flow_img)
        if getattr(self, "_stage_key", None) != key:
            self._stages = []
            for _ in range(depth):
                st = DeviceStore.__new__(DeviceStore)
                st.layout = lay
                st.rgb = torch.empty(vmax * self.MAX_RGB_PER_VIDEO * rgb_img, dtype=torch.uint8, device=dev)
                st.flow = torch.empty(vmax * self.MAX_FLOW_PER_VIDEO * flow_img, dtype=torch.uint8, device=dev)
                self._stages.append(st)
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._copied = [torch.cuda.Event() for _ in range(depth)]
            self._consumed = [torch.cuda.Event() for _ in range(depth)]
            self._stage_key = key
            self._plans, self._host_tables = {}, {}
        cur = torch.cuda.current_stream()
        for ev_ in self._consumed:
            ev_.record(cur)
        pending = {}

        def plan_for(k):
            if k not in self._plans:
                self._plans[k] = _VideoPlan(lay.videos[k], lay, self.L)
            return self._plans[k]

        def tables_for_slot(k, pos):
            if (k, pos) not in self._host_tables:
                pl = plan_for(k)
                ts, tt = pl.ts.copy(), pl.tt.copy()
                ts[:, :, 0] += pos * self.MAX_RGB_PER_VIDEO
                tt[:, :, 0] += pos * self.MAX_FLOW_PER_VIDEO
                self._host_tables[(k, pos)] = (torch.from_numpy(ts).pin_memory(), torch.from_numpy(tt).pin_memory())
            return self._host_tables[(k, pos)]

        def issue(gi):
            slot = gi % depth
            st = self._stages[slot]
            nb = 0
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._consumed[slot])
                tabs_s, tabs_t = [], []
                for pos, v in enumerate(groups[gi]):
                    k = v % len(lay.videos)
                    pl = plan_for(k)
                    base = pos * self.MAX_RGB_PER_VIDEO
                    for i, src in enumerate(pl.rgb_src):
                        st.rgb[(base + i) * rgb_img:(base + i + 1) * rgb_img].copy_(host.rgb[src * rgb_img:(src + 1) * rgb_img], non_blocking=True)
                    base = pos * self.MAX_FLOW_PER_VIDEO
                    for src, loc, cnt in pl.flow_runs:
                        st.flow[(base + loc) * flow_img:(base + loc + cnt) * flow_img].copy_(host.flow[src * flow_img:(src + cnt) * flow_img],
                                                                                           non_blocking=True)
                    hs, ht = tables_for_slot(k, pos)
                    tabs_s.append(hs.to(dev, non_blocking=True))
                    tabs_t.append(ht.to(dev, non_blocking=True))
                    nb += pl.n_rgb * rgb_img + pl.n_flow * flow_img + hs.numel() * 4 + ht.numel() * 4
                ts = torch.cat(tabs_s) if len(tabs_s) > 1 else tabs_s[0]
                tt = torch.cat(tabs_t) if len(tabs_t) > 1 else tabs_t[0]
                self._copied[slot].record(self._copy_stream)
            pending[gi] = (ts, tt, nb)

        for gi in range(min(depth - 1, len(groups))):
            issue(gi)
        if to_host is not None and getattr(self, "_host_out_key", None) != (tuple(to_host), vmax):
            self._host_out = [dict() for _ in range(2)]
            self._host_done = [torch.cuda.Event() for _ in range(2)]
            self._host_out_key = (tuple(to_host), vmax)
        prev = None
        for gi in range(len(groups)):
            slot = gi % depth
            ts, tt, nb = pending.pop(gi)
            cur.wait_event(self._copied[slot])
            ts.record_stream(cur)
            tt.record_stream(cur)
            res = self.run_tables(ts, tt, len(groups[gi]), store=self._stages[slot])
            self._consumed[slot].record(cur)
            hres = None
            if to_host is not None:
                hb, nd = self._host_out[gi & 1], 0
                hres = {}
                for name in to_host:
                    t = res[name]
                    if name not in hb or hb[name].shape[1:] != t.shape[1:] or hb[name].dtype != t.dtype:
                        hb[name] = torch.empty((vmax,) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory()
                    hres[name] = hb[name][:t.shape[0]]
                    hres[name].copy_(t, non_blocking=True)
                    nd += t.numel() * t.element_size()
                self._host_done[gi & 1].record(cur)
                self.last_d2h_bytes = nd
            # the next group's ~150 copy calls are issued AFTER this group's kernels are queued: the host spends ~0.5 ms on
            # them, which would otherwise be GPU idle time right after the caller's per-step synchronisation
            if gi + depth - 1 < len(groups):
                issue(gi + depth - 1)                 # its copies run under this group's networks
            self.last_h2d_bytes = nb
            if to_host is None:
                yield res
            else:
                if prev is not None:                  # group gi is queued: now hand out group gi-1
                    prev[1].synchronize()
                    yield prev[0]
                prev = (hres, self._host_done[gi & 1])
        if prev is not None:
            prev[1].synchronize()
            yield prev[0]

    def _stream(self, net: ops.StreamNet, images, shape, table, mean, std):
        """One stream over a table of snippets -> (descriptors, softmax scores).  bf16 handles gather the crops inside the
        first convolution (va_forward_store); fp32-parity handles go through the K1 tensor (va_preprocess + va_forward)."""
        if (self.fused_front_end and net.precision == "bf16" and
                ops.StreamNet.forward_store_supported(images, shape, int(table.shape[1]))):
            desc, _, prob, _ = net.forward_store(images, shape, table, mean, std, want_logits=False, want_pred=False)
        else:
            x = ops.preprocess(images, shape, table, mean, std, c_pad=net.c_pad)
            desc, _, prob, _ = net.forward(x, want_logits=False, want_pred=False)
        return desc, prob

    def alloc_outputs(self, n_rows: int, D: int, C: int, with_svm: bool) -> dict:
        dev = self.store.rgb.device
        out = {"video_desc": torch.zeros((n_rows, 2 * D), dtype=torch.float32, device=dev),
               "video_scores": torch.zeros((n_rows, C), dtype=torch.float32, device=dev),
               "score_pred": torch.full((n_rows,), -1, dtype=torch.int32, device=dev)}
        if with_svm:
            # one column per class the SVM was fitted on (LinearSVC: len(np.unique(y))), not per network class
            c_svm = int(self.combined.coef_.shape[0]) if self.combined.coef_ is not None else C
            out["svm_scores"] = torch.zeros((n_rows, c_svm), dtype=torch.float64, device=dev)
            out["svm_pred"] = torch.full((n_rows,), -1, dtype=torch.int32, device=dev)
        return out

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


class TwoStreamEvaluator:
    """Runs groups of videos through preprocess -> both streams -> consensus/fusion on the current device."""

    def __init__(self, spatial: ops.StreamNet, temporal: ops.StreamNet, store: DeviceStore,
                 combined: Optional[CombinedModel] = None, L: int = VIDEO_INPUT_FLOW_COUNT):
        self.spatial, self.temporal, self.store = spatial, temporal, store
        self.combined = combined if combined is not None else CombinedModel()
        self.L = L
        self.fused_front_end = False      # True: gather the crops inside conv1_1 (va_forward_store) instead of the K1 tensor
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

    def _stream(self, net: ops.StreamNet, images, shape, table, mean, std):
        """One stream over a table of snippets -> (descriptors, softmax scores).  bf16 handles gather the crops inside the
        first convolution (va_forward_store); fp32-parity handles go through the K1 tensor (va_preprocess + va_forward)."""
        if self.fused_front_end and net.precision == "bf16":
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

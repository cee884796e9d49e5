"""Temporal (optical-flow) stream -- API-compatible with the reference's `Sheet03/temporalModel.py`
(`TemporalDataset` :22-92, `TemporalNetwork` :96-312) on the sm_100a path.  A snippet is a stack of
L = flowSampleSize consecutive x/y flow image pairs, channel order x_t, y_t, x_{t+1}, y_{t+1}, ... (reference
:80-83), 2L = 20 channels.
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import VAError
from .parameters import *  # noqa: F401,F403
from .spatialModel import _StreamNetwork, _read_action_labels, swap_classifier
from .utils import check_index_rows, videoInfo


class TemporalDataset(torch.utils.data.Dataset):
    """2L flow images from a random start per video per call (reference temporalModel.py:22-92)."""

    image_channels = 1

    def __init__(self, videoListLoc, rootDir, imageTransforms=None, flowSampleSize=VIDEO_INPUT_FLOW_COUNT, mode="train",
                 actionLabelLoc=None, store=None, perImageCrops=True):
        super().__init__()
        self.rootDir = rootDir if (rootDir is None or rootDir.endswith("/")) else rootDir + "/"
        self.imageTransforms = imageTransforms
        self.flowSampleSize = flowSampleSize
        self.planes = 2 * flowSampleSize
        self.mode = mode
        with open(videoListLoc, "r") as f:
            self.videoList = [line for line in f]
        if actionLabelLoc is None:
            raise ValueError("Action label dictionary required!")          # reference :47-48
        self.actionLabelDict = _read_action_labels(actionLabelLoc)
        if store is None:
            # the reference's own call (:315-324): walk rootDir/<Category>/<video>/flow_{x,y}_%04d.jpg (:76-81) and decode
            # every flow image on the GPU, once
            if self.rootDir is None:
                raise VAError("TemporalDataset needs rootDir (a tree of flow-image folders) or a prebuilt store=DeviceStore")
            from .store import DeviceStore
            store = DeviceStore.from_directories(self.videoList, mode, flow_root=self.rootDir,
                                                 label_of=lambda cat: self.actionLabelDict[cat])
        self.store = store
        self._meta = {m.name: m for m in store.layout.videos}
        # The reference applies its random transform to each of the 2L images separately (:86), so every channel
        # gets its own crop and flip.  perImageCrops=False draws once per stack (the paper's behaviour).
        self.perImageCrops = perImageCrops
        self.last_indices = None

    def __len__(self):
        return len(self.videoList)

    def sample_indices(self, index):
        """Index-table rows [2L, 4] for one item; RNG consumption as reference :67-92."""
        _, videoName, actionLabel, actionCategory, _, _ = videoInfo(self.videoList[index], self.mode)
        if self.mode == "test":
            actionLabel = self.actionLabelDict[actionCategory]
        actionLabel = int(actionLabel)
        if self.imageTransforms is None:
            raise ValueError("imageTransforms is required")
        meta = self._meta[videoName]
        L = self.flowSampleSize
        nFlows = (2 * meta.n_flows) / 2                                     # reference :78 (x and y files)
        iFlowFrame = random.randint(1, int(nFlows - L))                      # reference :79
        h, w, _ = self.store.layout.flow_shape
        rows, crops = [], []
        shared = None
        for idx in range(iFlowFrame, iFlowFrame + L):
            for first in (meta.flowx_first, meta.flowy_first):              # x then y, alternating (:83)
                if self.perImageCrops or shared is None:
                    shared = self.imageTransforms.draw(h, w)
                i, j, flip = shared
                crops.append((i, j, flip))
                rows.append([first + idx - 1, i, j, flip])                  # flow_x_%04d is 1-based
        self.last_indices = dict(start=iFlowFrame, crops=crops)
        return np.array(rows, dtype=np.int32), actionLabel, videoName

    def check_rows(self, rows):
        """Host check of table rows against this dataset's store before upload (utils.check_index_rows)."""
        shape = self.store.layout.flow_shape
        buf = getattr(self.store, "flow", None)           # the bound that matters is the buffer the kernel will read
        n = self.store.layout.n_flow_images if buf is None else int(buf.numel()) // (shape[0] * shape[1] * shape[2])
        check_index_rows(rows, n, shape)

    def preprocess_table(self, table: torch.Tensor, reference_layout: bool = False, c_pad: int = 32):
        mean, std = self.imageTransforms.norm_constants(self.planes, 1)
        return ops.preprocess(self.store.flow, self.store.layout.flow_shape, table, mean, std, c_pad=c_pad,
                              reference_layout=reference_layout)

    def __getitem__(self, index):
        rows, label, name = self.sample_indices(index)
        self.check_rows(rows)
        table = torch.from_numpy(rows[None]).cuda()
        return self.preprocess_table(table, reference_layout=True)[0], label, name


def copy_first_layer(model, flowSampleSize):
    """reference __copyFirstLayer__ (temporalModel.py:149-162): the RGB kernel averaged over its 3 input channels is
    copied into each of the 2L input channels of a new first conv; that conv keeps its own freshly initialised bias."""
    first = model.features[0]
    mean_kernel = torch.zeros_like(first.weight.data[:, 0])
    for c in range(first.in_channels):                                   # the reference's summation order
        mean_kernel = mean_kernel + first.weight.data[:, c]
    mean_kernel = mean_kernel / first.in_channels
    replacement = nn.Conv2d(2 * flowSampleSize, first.out_channels, kernel_size=first.kernel_size, padding=first.padding)
    replacement.weight.data.copy_(mean_kernel.unsqueeze(1).expand_as(replacement.weight.data))
    model.features[0] = replacement


def build_temporal_torch_model(nActionClasses=NACTION_CLASSES, flowSampleSize=VIDEO_INPUT_FLOW_COUNT,
                               descriptorDim=VIDEO_DESCRIPTOR_DIM, pretrained=False, seed=None):
    """Parameter container of the temporal stream (reference temporalModel.py:122-126)."""
    import torchvision.models as models
    if seed is not None:
        torch.manual_seed(seed)
    model = models.vgg16(weights="IMAGENET1K_V1" if pretrained else None)
    copy_first_layer(model, flowSampleSize)
    swap_classifier(model, descriptorDim, nActionClasses)
    return model


class TemporalNetwork(_StreamNetwork):
    """A wrapper for the motion stream (reference temporalModel.py:96-312)."""

    _ckp_file, _best_file = MOTION_CKP_FILE, MOTION_BEST_FILE
    _perf_loc, _train_csv, _test_csv = TEMPORAL_PERFORMANCE_LOC, TEMPORAL_TRAIN_CSV_LOC, TEMPORAL_TEST_CSV_LOC
    _stream_kind = ops.STREAM_TEMPORAL

    def __init__(self, nActionClasses, flowSampleSize, nEpochs, lr, momentumVal, descriptorDim, trainLoader, testLoader,
                 lrMilestones, ckpLoc, gpu=False, pretrained=False, maxBatch=GPU_MAX_BATCH, precision="bf16"):
        super().__init__()
        self.flowSampleSize = flowSampleSize
        self._in_channels = 2 * flowSampleSize
        self._init_common(nActionClasses, nEpochs, lr, momentumVal, descriptorDim, trainLoader, testLoader, lrMilestones,
                          ckpLoc, gpu, pretrained, maxBatch, precision)

    def _build_torch_model(self, pretrained):
        return build_temporal_torch_model(self.nActionClasses, self.flowSampleSize, self.descriptorDim, pretrained)


TemporalModel = TemporalNetwork    # alias named by BASELINE.json's north_star

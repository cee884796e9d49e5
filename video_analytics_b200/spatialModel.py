"""Spatial (RGB) stream -- API-compatible with the reference's `Sheet03/spatialModel.py`
(`SpatialDataset` :21-81, `SpatialNetwork` :85-283), driving the hand-written sm_100a path in libva_b200.so.

Differences a reference user will notice:
  * the JPEG folders under `rootDir` are decoded ONCE, on the GPU, into a device-resident `DeviceStore` when the dataset
    is constructed (`DeviceStore.from_directories`); a prebuilt store can be passed with `store=` instead;
  * `__getitem__` still returns the reference's fp32 [3,224,224] tensor (bit-exact), produced by the CUDA
    preprocess kernel; batched use goes through `utils.getDataLoader`, which yields NHWC bf16 batches;
  * the forward pass inlined in the reference's train()/validate() is the explicit method `forward(ip)`;
  * `train()` runs the K5 kernels (training.py): no torch autograd, no cuDNN.
"""
from __future__ import annotations

import os
import random
import time
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import VAError
from .parameters import *  # noqa: F401,F403
from .utils import (AverageMeter, DeviceVideoDict, SnippetBatch, checkAndMakeDirectories, check_index_rows,
                    makeCheckpoint, saveVideoDescriptors, savePerformance, videoInfo)


def _read_action_labels(actionLabelLoc):
    """`<id> <ClassName>` lines -> {ClassName: id} (reference spatialModel.py:47-53)."""
    table = {}
    with open(actionLabelLoc, "r") as f:
        for line in f:
            val, key = line.split(" ")
            table[key.strip()] = int(val)
    return table


class SpatialDataset(torch.utils.data.Dataset):
    """One random frame per video per call (reference spatialModel.py:21-81)."""

    planes = 1            # one RGB image -> 3 channels
    image_channels = 3

    def __init__(self, videoListLoc, rootDir, imageTransforms=None, frameSampleSize=VIDEO_INPUT_FRAME_COUNT, mode="train",
                 actionLabelLoc=None, store=None):
        super().__init__()
        self.rootDir = rootDir if (rootDir is None or rootDir.endswith("/")) else rootDir + "/"
        self.imageTransforms = imageTransforms
        self.frameSampleSize = frameSampleSize      # stored, unused -- as in the reference (:41)
        self.mode = mode
        with open(videoListLoc, "r") as f:
            self.videoList = [line for line in f]
        if actionLabelLoc is None:
            raise ValueError("Action label dictionary required!")          # reference :45-46
        self.actionLabelDict = _read_action_labels(actionLabelLoc)
        if store is None:
            # the reference's own call: SpatialDataset(videoListLoc, rootDir, transforms, actionLabelLoc=...) (:286-298).
            # Walk rootDir/<Category>/<video>/<i>.jpg like __getitem__ :72-77 does and decode every frame on the GPU.
            if self.rootDir is None:
                raise VAError("SpatialDataset needs rootDir (a tree of JPEG frame folders) or a prebuilt store=DeviceStore")
            from .store import DeviceStore
            store = DeviceStore.from_directories(self.videoList, mode, frames_root=self.rootDir,
                                                 label_of=lambda cat: self.actionLabelDict[cat])
        self.store = store
        self._meta = {m.name: m for m in store.layout.videos}
        self.last_indices = None

    def __len__(self):
        return len(self.videoList)

    # -- host side: the reference's RNG draws, nothing else
    def sample_indices(self, index):
        """Index-table rows [planes, 4] = (image id, crop top, crop left, flip) for one item, consuming Python's
        `random` (frame) and the torch RNG (crop, flip, jitter order) exactly like reference :64-81."""
        _, videoName, actionLabel, actionCategory, _, _ = videoInfo(self.videoList[index], self.mode)
        if self.mode == "test":
            actionLabel = self.actionLabelDict[actionCategory]
        if self.imageTransforms is None:
            raise ValueError("imageTransforms is required (the reference's no-transform branch is not runnable either)")
        meta = self._meta[videoName]
        frameName = random.randint(0, meta.n_frames - 1)                     # reference :75
        h, w, _ = self.store.layout.rgb_shape
        i, j, flip = self.imageTransforms.draw(h, w)
        self.last_indices = dict(frame=frameName, crops=[(i, j, flip)])
        rows = np.array([[meta.rgb_first + frameName, i, j, flip]], dtype=np.int32)
        return rows, int(actionLabel), videoName

    def check_rows(self, rows):
        """Host check of table rows against this dataset's store before upload (utils.check_index_rows)."""
        shape = self.store.layout.rgb_shape
        buf = getattr(self.store, "rgb", None)           # the bound that matters is the buffer the kernel will read
        n = self.store.layout.n_rgb_images if buf is None else int(buf.numel()) // (shape[0] * shape[1] * shape[2])
        check_index_rows(rows, n, shape)

    # -- device side
    def preprocess_table(self, table: torch.Tensor, reference_layout: bool = False, c_pad: int = 16):
        mean, std = self.imageTransforms.norm_constants(3, 3)
        return ops.preprocess(self.store.rgb, self.store.layout.rgb_shape, table, mean, std, c_pad=c_pad,
                              reference_layout=reference_layout)

    def __getitem__(self, index):
        rows, label, name = self.sample_indices(index)
        self.check_rows(rows)
        table = torch.from_numpy(rows[None]).cuda()
        return self.preprocess_table(table, reference_layout=True)[0], label, name


class _StreamNetwork(object):
    """Shared body of SpatialNetwork / TemporalNetwork (the reference duplicates it line for line)."""

    _ckp_file, _best_file = SPATIAL_CKP_FILE, SPATIAL_BEST_FILE
    _perf_loc, _train_csv, _test_csv = SPATIAL_PERFORMANCE_LOC, SPATIAL_TRAIN_CSV_LOC, SPATIAL_TEST_CSV_LOC
    _stream_kind, _in_channels = ops.STREAM_SPATIAL, 3

    def _build_torch_model(self, pretrained):
        raise NotImplementedError

    def _init_common(self, nActionClasses, nEpochs, lr, momentumVal, descriptorDim, trainLoader, testLoader, lrMilestones,
                     ckpLoc, gpu, pretrained, maxBatch, precision="bf16"):
        import torch.optim.lr_scheduler as slr
        self.nActionClasses, self.nEpochs, self.lr = nActionClasses, nEpochs, lr
        self.trainLoader, self.testLoader = trainLoader, testLoader
        self.totalTrain = len(trainLoader.dataset) if trainLoader is not None else 0
        self.totalTest = len(testLoader.dataset) if testLoader is not None else 0
        self.lrMilestones, self.descriptorDim, self.gpu = lrMilestones, descriptorDim, gpu
        if not self.gpu or not torch.cuda.is_available():
            raise VAError("this implementation runs on a B200 only: construct with gpu=True on a CUDA machine "
                          "(the reference's CPU branch is served by the reference itself)")
        print("GPU available!")                                            # reference :105-108
        self.model = self._build_torch_model(pretrained)                  # parameter container (init-time only)
        self.criterion = nn.CrossEntropyLoss().cuda()
        self.optimizer = torch.optim.SGD(self.model.parameters(), self.lr, momentum=momentumVal)
        self.startEpoch = 0
        self.scheduler = slr.MultiStepLR(self.optimizer, lrMilestones, gamma=0.1, last_epoch=-1)
        self.highestPrecision, self.isBest = 0.0, False
        if not ckpLoc.endswith("/"):
            ckpLoc += "/"
        checkAndMakeDirectories(ckpLoc)
        self.ckpLoc = ckpLoc
        self.resumeLoc = self.ckpLoc + self._ckp_file
        self.features = self.model.features
        self.classifierList = list(self.model.classifier)
        self.classifierLen = len(self.classifierList)
        self.trainDict = DeviceVideoDict(descriptorDim)
        self.testDict = DeviceVideoDict(descriptorDim)
        self.model = nn.DataParallel(self.model)                           # reference :133 -- gives "module." keys
        self.epoch = 0
        self.net = ops.StreamNet(self._stream_kind, self._in_channels, nActionClasses, descriptorDim, max_batch=maxBatch,
                                 precision=precision)
        self.sync_weights()

    def sync_weights(self):
        """(Re)pack the torch parameters into the CUDA handle -- after construction, resume() or external edits."""
        self.net.load_state_dict(self.model.state_dict())

    # ---- the forward pass the reference inlines in train()/validate() (:171-177, :212-218)
    def forward(self, ip):
        """ip: SnippetBatch (bf16 NHWC from the loader) or a reference-layout float tensor [B,C,224,224].
        Returns (featureVectors [B,descriptorDim], logits [B,nActionClasses]) as CUDA fp32 tensors."""
        desc, logits, _, _ = self._forward_full(ip)
        return desc, logits

    def _forward_full(self, ip):
        if isinstance(ip, SnippetBatch):
            if self.net.precision != "bf16":
                raise VAError("loader batches are bf16 snippets; feed the fp32 mode reference-layout float tensors "
                              "(dataset[i][0] / torch.stack) so that no input bits are lost")
            x = ip.nhwc
        else:
            x = self.net.pack_input(ip)
        return self.net.forward(x)

    def _get_trainer(self):
        if getattr(self, "_trainer", None) is None:
            from .training import StreamTrainer
            group = None
            if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                    torch.distributed.get_world_size() > 1:
                group = torch.distributed.group.WORLD
            self._trainer = StreamTrainer(self.model.module, self.optimizer, c_pad=self.net.c_pad, process_group=group)
        return self._trainer

    def train(self):
        """reference :161-194 -- one epoch of SGD-momentum steps on the K5 kernels (training.py); there is no torch
        autograd on this path.  The reference's clip_grad_norm_ after the loop (:190) acts on the gradients of the last
        batch AFTER their step was applied and before the next zero_grad(): it cannot change a parameter and is not
        reproduced."""
        if self.net.precision != "bf16":
            raise VAError("training runs in the bf16 mode only")
        self.model.train()
        startTime = time.time()
        trainer = self._get_trainer()
        for iBatch, (data, labels, videoNames) in enumerate(self.trainLoader):
            x = data.nhwc if isinstance(data, SnippetBatch) else self.net.pack_input(data)
            loss, featureVectors, op = trainer.step(x, labels)
            self.trainDict.update_batch(videoNames, labels, featureVectors)
        self.sync_weights()                                                # evaluation handle <- updated parameters
        duration = time.time() - startTime
        print("Epoch %d completed in %lf seconds" % (self.epoch, duration))
        self.save()

    def validate(self):
        """reference :197-231 -- eval-mode pass over the test loader; returns (precision, summed CE loss)."""
        self.model.eval()
        correct = 0
        loss = 0
        for iBatch, (data, labels, videoNames) in enumerate(self.testLoader):
            labelVar = labels.cuda(non_blocking=True)
            featureVectors, op, _, pred = self._forward_full(data)
            loss += self.criterion(op, labelVar)                            # scalar bookkeeping on the logits
            correct += int((pred.to(torch.int64) == labelVar).sum().item())
            self.testDict.update_batch(videoNames, labels, featureVectors)
        total = self.totalTest
        world, _ = self._world_rank()
        if world > 1:
            # each rank validated its shard of the test list (utils.RankShardSampler): sum the counts and the losses
            acc = torch.tensor([float(correct), float(loss)], dtype=torch.float64, device="cuda")
            torch.distributed.all_reduce(acc)
            correct, loss = int(acc[0].item()), acc[1].to(torch.float32)
            total = world * ((self.totalTest + world - 1) // world)       # shards are padded to equal length
        print("Validation for epoch %d: total = %d, correct = %d, loss = %f" % (self.epoch, total, correct, float(loss)))
        return (correct / total), loss

    def resume(self):
        """reference :234-249."""
        import torch.optim.lr_scheduler as slr
        if not (self.resumeLoc and os.path.isfile(self.resumeLoc)):
            print("No checkpoints found; starting from scratch!")
            return False
        print("Resuming training from checkpoint file: %s" % (self.resumeLoc))
        checkpoint = torch.load(self.resumeLoc, weights_only=False)
        self.startEpoch = checkpoint["epoch"] + 1
        self.highestPrecision = checkpoint["highestPrecision"]
        self.model.load_state_dict(checkpoint["model"])
        self.optimizer.load_state_dict(checkpoint["optimizer"])
        self.scheduler = slr.MultiStepLR(self.optimizer, self.lrMilestones, gamma=0.1, last_epoch=self.startEpoch)
        if getattr(self, "_trainer", None) is not None:
            self._trainer._adopt_optimizer_state()
        self.sync_weights()
        print("Loaded checkpoint: starting from epoch: %d" % (self.startEpoch))
        return True

    @staticmethod
    def _world_rank():
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_world_size(), torch.distributed.get_rank()
        return 1, 0

    def save(self):
        """reference :252-262 -- same checkpoint dict, same file names.  One process per GPU: replicas are identical
        (StreamTrainer.sync_replicas + all-reduced gradients), rank 0 alone writes."""
        if self._world_rank()[1] != 0:
            return
        makeCheckpoint({"epoch": self.epoch, "model": self.model.state_dict(), "highestPrecision": self.highestPrecision,
                        "optimizer": self.optimizer.state_dict()}, self.isBest, self.ckpLoc + self._ckp_file,
                       self.ckpLoc + self._best_file)

    def execute(self, evalOnly=False):
        """reference :265-283.  With evalOnly=True the training step is skipped and each epoch is one
        validation pass -- that is the reference's descriptor-accumulation loop (25 epochs = 25 random snippets)."""
        self.resume()
        for self.epoch in range(self.startEpoch, self.nEpochs):
            if not evalOnly:
                self.train()
            precision, loss = self.validate()
            if precision > self.highestPrecision:
                self.highestPrecision = precision
                self.isBest = True
            # reference :278 `self.scheduler.step(loss)`: the summed validation loss is passed where MultiStepLR expects
            # the epoch index, so the learning rate is lr * 0.1 ** bisect_right(milestones, loss) -- a function of the
            # loss VALUE.  Reproduced as written; torch's closed form gives the same number.
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.scheduler.step(float(loss))
            self.save()
            world, rank = self._world_rank()
            if world == 1:
                savePerformance(precision, float(loss), self._perf_loc)
                saveVideoDescriptors(self.trainDict, self._train_csv, self.gpu)
                saveVideoDescriptors(self.testDict, self._test_csv, self.gpu)
            else:
                # every rank holds the descriptors of ITS shard of the videos: each writes a part file, rank 0 joins them
                # into the reference's CSVs (one writer per file: no write race)
                for d, path in ((self.trainDict, self._train_csv), (self.testDict, self._test_csv)):
                    saveVideoDescriptors(d, path + ".rank%d" % rank, self.gpu)
                torch.distributed.barrier()
                if rank == 0:
                    savePerformance(precision, float(loss), self._perf_loc)
                    for path in (self._train_csv, self._test_csv):
                        with open(path, "w") as out:
                            for r in range(world):
                                with open(path + ".rank%d" % r) as part:
                                    out.write(part.read())
                                os.remove(path + ".rank%d" % r)
                torch.distributed.barrier()


def swap_classifier(model, descriptorDim, nActionClasses):
    """reference __swapClassifier__ (spatialModel.py:136-152): 25088-4096-4096-descriptorDim-nActionClasses."""
    model.classifier = nn.Sequential(
        nn.Linear(512 * 7 * 7, 4096), nn.ReLU(True), nn.Dropout(),
        nn.Linear(4096, 4096), nn.ReLU(True), nn.Dropout(),
        nn.Linear(4096, descriptorDim), nn.ReLU(True), nn.Dropout(),
        nn.Linear(descriptorDim, nActionClasses))


def build_spatial_torch_model(nActionClasses=NACTION_CLASSES, descriptorDim=VIDEO_DESCRIPTOR_DIM, pretrained=False,
                              seed=None):
    """The parameter container of the spatial stream (reference spatialModel.py:110-113): torchvision VGG16 with the
    swapped classifier.  The reference loads ImageNet weights (:110); offline this is random init unless a local
    torchvision cache provides them (pretrained=True)."""
    import torchvision.models as models
    if seed is not None:
        torch.manual_seed(seed)
    model = models.vgg16(weights="IMAGENET1K_V1" if pretrained else None)
    swap_classifier(model, descriptorDim, nActionClasses)
    return model


class SpatialNetwork(_StreamNetwork):
    """A wrapper for the spatial stream (reference spatialModel.py:85-283)."""

    def __init__(self, nActionClasses, nEpochs, lr, momentumVal, descriptorDim, trainLoader, testLoader, lrMilestones,
                 ckpLoc, gpu=False, pretrained=False, maxBatch=GPU_MAX_BATCH, precision="bf16"):
        super().__init__()
        self._init_common(nActionClasses, nEpochs, lr, momentumVal, descriptorDim, trainLoader, testLoader, lrMilestones,
                          ckpLoc, gpu, pretrained, maxBatch, precision)

    def _build_torch_model(self, pretrained):
        return build_spatial_torch_model(self.nActionClasses, self.descriptorDim, pretrained)


SpatialModel = SpatialNetwork      # alias named by BASELINE.json's north_star


def main():
    """reference :286-298, on the synthetic device store (no dataset is reachable offline)."""
    from .store import DeviceStore, make_layout
    from .utils import getDataLoader, getTransforms
    import tempfile
    lay = make_layout(8)
    store = DeviceStore(lay)
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "list.txt"), "w") as f:
        f.writelines(lay.list_line(v, "train") for v in range(8))
    with open(os.path.join(tmp, "classInd.txt"), "w") as f:
        f.writelines(f"{m.label} {m.category}\n" for m in lay.videos)
    tr = getTransforms()
    ds = SpatialDataset(os.path.join(tmp, "list.txt"), None, tr, actionLabelLoc=os.path.join(tmp, "classInd.txt"), store=store)
    loader = getDataLoader(ds, batchSize=SPATIAL_BATCH_SIZE)
    net = SpatialNetwork(NACTION_CLASSES, 1, INITIAL_LR, MOMENTUM_VAL, VIDEO_DESCRIPTOR_DIM, loader, loader, MILESTONES_LR,
                         os.path.join(tmp, "ckp"), gpu=True)
    t0 = time.time()
    print(net.validate(), time.time() - t0)


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE -- CPU oracle: a Python-3 restatement of the reference's Sheet03 two-stream path on stock
PyTorch fp32.  NOT product code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it.  Every function cites the reference lines (under /root/reference/Sheet03) it follows.

Pinning status: the reference ships NO tests, golden vectors or fixtures for this path.  The
restatement is pinned instead against the reference's own importable modules run in the build container
(`utils.py`, `parameters.py`: transforms bit-for-bit, videoInfo on all 3360 list lines, AverageMeter) by
oracle/make_golden.py, whose outputs are committed under tests/golden/.  The three model scripts are Python-2 only
(print statements, `.cuda(async=True)`, `it.next()`) and cannot be imported, so the model/forward/fusion part is a
line-by-line restatement on the installed torch 2.11 / torchvision 0.26 / sklearn -- "parity unpinned" by the
reference itself for those rows.

What is replaced relative to the reference: `os.listdir` / `PIL.Image.open` on JPEG folders become lookups in an
in-memory synthetic frame store (oracle/synth.py); nothing else.
"""
from __future__ import annotations

import csv
import io
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

# parameters.py:4-5,10-14,19
VIDEO_INPUT_FLOW_COUNT = 10
CROP_SIZE_TF = 224
NORM_MEANS_TF = [0.485, 0.456, 0.406]
NORM_STDS_TF = [0.229, 0.224, 0.225]
NACTION_CLASSES = 101
VIDEO_DESCRIPTOR_DIM = 256
# 25 x 10 test protocol, notes.txt:113-116 and 225-230
N_TEST_SNIPPETS = 25
N_TEST_CROPS = 10


# ----------------------------------------------------------------------------------------------- S1 utils.py:73-91
def videoInfo(line: str, mode: str):
    actionLabel = None
    if mode == "train":
        videoLoc, actionLabel = line.split(" ")
        actionLabel = actionLabel.strip()
    else:
        videoLoc = line
    videoLoc = videoLoc.strip()
    actionCategory, videoName = videoLoc.split("/")
    actionCategory = actionCategory.strip()
    videoName = videoName[:videoName.rfind(".")]
    _, _, ngroup, nclip = videoName.split("_")
    return videoLoc, videoName, actionLabel, actionCategory, ngroup, nclip


# ----------------------------------------------------------------------------------------------- C1 utils.py:154-171
class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


# ----------------------------------------------------------------------------------------------- S4 utils.py:137-151
def draw_transform_params(h: int, w: int, crop: int = CROP_SIZE_TF) -> Tuple[int, int, int]:
    """RNG draws of Compose[RandomCrop(224), RandomHorizontalFlip(), ColorJitter(0,0,0,0), ...] for one image, in
    the order the installed torchvision 0.26 makes them: randint(i), randint(j) [skipped when
    the image is exactly crop x crop], rand(1) < 0.5, randperm(4) (ColorJitter draws its op order even when it is
    the identity)."""
    if w == crop and h == crop:
        i, j = 0, 0
    else:
        i = int(torch.randint(0, h - crop + 1, size=(1,)).item())
        j = int(torch.randint(0, w - crop + 1, size=(1,)).item())
    flip = int(bool(torch.rand(1) < 0.5))
    torch.randperm(4)
    return i, j, flip


def apply_transform(img_hwc_u8: np.ndarray, i: int, j: int, flip: int, means: Sequence[float], stds: Sequence[float],
                    crop: int = CROP_SIZE_TF) -> torch.Tensor:
    """crop -> hflip -> ToTensor (u8 -> f32 / 255, HWC -> CHW) -> Normalize (x.sub_(mean).div_(std)), fp32.
    For 1-channel flow images the 2018 torchvision Normalize zipped (tensor, mean, std) and so used only mean[0],
    std[0]; pass means=[0.485], stds=[0.229] for that case."""
    patch = img_hwc_u8[i:i + crop, j:j + crop, :]
    if flip:
        patch = patch[:, ::-1, :]
    x = torch.from_numpy(np.ascontiguousarray(patch)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    mean = torch.as_tensor(list(means), dtype=torch.float32).view(-1, 1, 1)
    std = torch.as_tensor(list(stds), dtype=torch.float32).view(-1, 1, 1)
    return x.sub_(mean).div_(std)


def flow_norm_constants(n_planes: int):
    return [NORM_MEANS_TF[0]] * n_planes, [NORM_STDS_TF[0]] * n_planes


# ----------------------------------------------------------------------------------------------- frame store
class OracleStore:
    """In-memory replacement for the frame / flow JPEG folders."""

    def __init__(self, layout, rgb: np.ndarray, flow: np.ndarray):
        self.layout, self.rgb, self.flow = layout, rgb, flow
        self.by_name = {m.name: m for m in layout.videos}

    def n_frame_files(self, name):      # len(os.listdir(frameDir)), spatialModel.py:74
        return self.by_name[name].n_frames

    def frame(self, name, k):           # Image.open(frameDir + str(k) + ".jpg"), spatialModel.py:77
        return self.rgb[self.by_name[name].rgb_first + k]

    def n_flow_files(self, name):       # len(os.listdir(flowDir)), temporalModel.py:78
        return 2 * self.by_name[name].n_flows

    def flow_x(self, name, idx):        # flow_x_%04d.jpg, 1-based (temporalModel.py:80, parameters.py:38)
        return self.flow[self.by_name[name].flowx_first + idx - 1]

    def flow_y(self, name, idx):
        return self.flow[self.by_name[name].flowy_first + idx - 1]


# ----------------------------------------------------------------------------------------------- S2 spatialModel.py:64-81
class SpatialDataset:
    def __init__(self, videoList: List[str], store: OracleStore, mode="train", actionLabelDict: Optional[Dict] = None):
        if actionLabelDict is None:
            raise ValueError("Action label dictionary required!")      # spatialModel.py:45-46
        self.videoList, self.store, self.mode, self.actionLabelDict = videoList, store, mode, actionLabelDict
        self.last_indices = None

    def __len__(self):
        return len(self.videoList)

    def __getitem__(self, index):
        _, videoName, actionLabel, actionCategory, _, _ = videoInfo(self.videoList[index], self.mode)
        if self.mode == "test":
            actionLabel = self.actionLabelDict[actionCategory]
        nFrames = self.store.n_frame_files(videoName)
        frameName = random.randint(0, nFrames - 1)                        # :75
        img = self.store.frame(videoName, frameName)
        i, j, flip = draw_transform_params(img.shape[0], img.shape[1])
        self.last_indices = dict(frame=frameName, crops=[(i, j, flip)])
        loadedFrame = apply_transform(img, i, j, flip, NORM_MEANS_TF, NORM_STDS_TF)
        return loadedFrame, int(actionLabel), videoName


# ----------------------------------------------------------------------------------------------- S3 temporalModel.py:67-92
class TemporalDataset:
    def __init__(self, videoList: List[str], store: OracleStore, flowSampleSize=VIDEO_INPUT_FLOW_COUNT, mode="train",
                 actionLabelDict: Optional[Dict] = None):
        if actionLabelDict is None:
            raise ValueError("Action label dictionary required!")
        self.videoList, self.store, self.mode, self.actionLabelDict = videoList, store, mode, actionLabelDict
        self.flowSampleSize = flowSampleSize
        self.last_indices = None

    def __len__(self):
        return len(self.videoList)

    def __getitem__(self, index):
        _, videoName, actionLabel, actionCategory, _, _ = videoInfo(self.videoList[index], self.mode)
        if self.mode == "test":
            actionLabel = self.actionLabelDict[actionCategory]
        actionLabel = int(actionLabel)
        nFlows = self.store.n_flow_files(videoName) / 2                   # :78 (true division -> float)
        iFlowFrame = random.randint(1, int(nFlows - self.flowSampleSize))  # :79 (py2 randint took the integral float)
        L = self.flowSampleSize
        xs = [("x", idx) for idx in range(iFlowFrame, iFlowFrame + L)]    # :80
        ys = [("y", idx) for idx in range(iFlowFrame, iFlowFrame + L)]    # :81
        flowFrames = [f for pair in zip(xs, ys) for f in pair]            # :83 x,y alternating
        mean, std = flow_norm_constants(1)
        loaded, crops = [], []
        for kind, idx in flowFrames:                                      # :86 transform applied PER IMAGE
            img = self.store.flow_x(videoName, idx) if kind == "x" else self.store.flow_y(videoName, idx)
            i, j, flip = draw_transform_params(img.shape[0], img.shape[1])
            crops.append((i, j, flip))
            loaded.append(apply_transform(img, i, j, flip, mean, std))
        self.last_indices = dict(start=iFlowFrame, crops=crops)
        flowVolume = torch.squeeze(torch.stack(loaded, dim=0))            # :90 -> [2L,224,224]
        return flowVolume, actionLabel, videoName


# ----------------------------------------------------------------------------------------------- N1 model surgery
def _swap_classifier(model, descriptorDim, nActionClasses):              # spatialModel.py:136-152
    model.classifier = nn.Sequential(
        nn.Linear(512 * 7 * 7, 4096), nn.ReLU(True), nn.Dropout(),
        nn.Linear(4096, 4096), nn.ReLU(True), nn.Dropout(),
        nn.Linear(4096, descriptorDim), nn.ReLU(True), nn.Dropout(),
        nn.Linear(descriptorDim, nActionClasses))


def build_spatial_model(nActionClasses=NACTION_CLASSES, descriptorDim=VIDEO_DESCRIPTOR_DIM, seed=0):
    """spatialModel.py:110-113 with weights=None (BASELINE: random init; no network for the ImageNet weights)."""
    import torchvision.models as models
    torch.manual_seed(seed)
    model = models.vgg16(weights=None)
    _swap_classifier(model, descriptorDim, nActionClasses)
    return model


def build_temporal_model(nActionClasses=NACTION_CLASSES, flowSampleSize=VIDEO_INPUT_FLOW_COUNT,
                         descriptorDim=VIDEO_DESCRIPTOR_DIM, seed=0):
    """temporalModel.py:122-126 + __copyFirstLayer__ :149-162 (channel-mean of the 3-channel kernel copied into all
    2L input channels; the new Conv2d keeps its own fresh default-init bias) + __swapClassifier__ :165-181."""
    import torchvision.models as models
    torch.manual_seed(seed)
    model = models.vgg16(weights=None)
    layerOne = model.features[0]
    avg = 0
    for inChannel in range(layerOne.in_channels):
        avg = avg + layerOne.weight[:, inChannel, :, :]
    avg = avg / layerOne.in_channels
    newLayerOne = nn.Conv2d(flowSampleSize * 2, layerOne.out_channels, kernel_size=layerOne.kernel_size,
                            padding=layerOne.padding)
    for inChannel in range(2 * flowSampleSize):
        newLayerOne.weight.data[:, inChannel, :, :] = avg.data
    model.features[0] = newLayerOne
    _swap_classifier(model, descriptorDim, nActionClasses)
    return model


# ----------------------------------------------------------------------------------------------- N2-N4 forward
@torch.no_grad()
def forward_eval(model, ip: torch.Tensor):
    """validate()'s inlined forward (spatialModel.py:212-221): returns (featureVectors [B,256], logits [B,C],
    pred [B]) in eval mode (Dropout = identity)."""
    model.eval()
    classifierList = list(model.classifier)
    classifierLen = len(classifierList)
    op = model.features(ip)
    op = op.view(op.size(0), -1)                                           # NCHW flatten c*49+h*7+w
    for cl in classifierList[:(classifierLen - 1)]:
        op = cl(op)
    featureVectors = op
    for cl in classifierList[(classifierLen - 1):]:
        op = cl(op)
    pred = op.max(1, keepdim=True)[1].view(-1)                              # :220
    return featureVectors, op, pred


# ----------------------------------------------------------------------------------------------- N5 training step
def draw_dropout_masks(shapes, p=0.5):
    """The keep-masks nn.Dropout(p) draws in train mode on CPU, in call order: torch's dropout is
    input.new_empty(shape).bernoulli_(1 - p) (then .div_(1 - p) and a multiply).  Drawing them here with the same global
    RNG state gives the masks the reference's classifier would use -- tests/test_oracle_golden.py pins that."""
    return [torch.empty(s).bernoulli_(1 - p) for s in shapes]


def train_step(model, optimizer, criterion, ip: torch.Tensor, labels: torch.Tensor, dropout_masks):
    """Body of the train() loop (spatialModel.py:171-181 == temporalModel.py same lines) for one batch, with the three
    Dropout keep-masks supplied by the caller instead of drawn inside nn.Dropout (same arithmetic:
    x * mask / (1 - p)).  Returns (loss, featureVectors, logits); `optimizer` has stepped."""
    model.train()                                                           # :165
    classifierList = list(model.classifier)
    classifierLen = len(classifierList)
    masks = iter(dropout_masks)

    def run(cl, op):
        if isinstance(cl, nn.Dropout):
            return op * next(masks).to(op.dtype) / (1.0 - cl.p)
        return cl(op)

    op = model.features(ip)                                                 # :171
    op = op.view(op.size(0), -1)                                            # :172
    for cl in classifierList[:(classifierLen - 1)]:                         # :173-174
        op = run(cl, op)
    featureVectors = op                                                     # :175
    for cl in classifierList[(classifierLen - 1):]:                         # :176-177
        op = run(cl, op)
    loss = criterion(op, labels)                                            # :178
    optimizer.zero_grad()                                                   # :179
    loss.backward()                                                         # :180
    optimizer.step()                                                        # :181
    return loss.detach(), featureVectors.detach(), op.detach()


# ----------------------------------------------------------------------------------------------- C1 consensus
def update_video_dict(videoDict: dict, videoNames, labels, featureVectors):
    """spatialModel.py:223-228."""
    for i in range(len(featureVectors)):
        if videoNames[i] in videoDict:
            videoDict[videoNames[i]][0].update(featureVectors[i])
        else:
            videoDict[videoNames[i]] = (AverageMeter(), labels[i])
            videoDict[videoNames[i]][0].update(featureVectors[i])


# ----------------------------------------------------------------------------------------------- C2 utils.py:174-195
def saveVideoDescriptors(videoDescDict, csvFile: io.TextIOBase):
    writer = csv.writer(csvFile, delimiter=",")
    for videoName in videoDescDict.keys():
        videoLabel = np.asarray(videoDescDict[videoName][1])
        csvFile.write(videoName + "," + str(videoLabel) + ",")
        videoDesc = videoDescDict[videoName][0].avg.detach().cpu().numpy().astype(float)
        writer.writerow(videoDesc)


# ----------------------------------------------------------------------------------------------- F1 combinedModel.py:9-26
def combineDescriptors(spatialCsv, temporalCsv, descriptorDim=VIDEO_DESCRIPTOR_DIM):
    import pandas as pd
    headers = ["vidname", "label"]
    for dim in range(descriptorDim):
        headers.append("dim" + str(dim))
    dfSpatial = pd.read_csv(spatialCsv, names=headers)
    dfTemporal = pd.read_csv(temporalCsv, names=headers)
    dfMerged = pd.merge(dfSpatial, dfTemporal, on="vidname", how="inner", suffixes=("_s", "_t"))
    spatialHeaders = [headers[i] + "_s" for i in range(2, len(headers))]
    temporalHeaders = [headers[i] + "_t" for i in range(2, len(headers))]
    spatialHeaders.extend(temporalHeaders)
    descriptors = dfMerged[spatialHeaders].values
    labels = dfMerged["label_s"].values
    return descriptors, labels


# ----------------------------------------------------------------------------------------------- F2 combinedModel.py:38
def svm_decision(X: np.ndarray, coef: np.ndarray, intercept: np.ndarray):
    """LinearSVC.predict == classes_[argmax(X @ coef.T + intercept)] in fp64 (one-vs-rest)."""
    scores = np.asarray(X, dtype=np.float64) @ np.asarray(coef, dtype=np.float64).T + np.asarray(intercept, np.float64)
    return scores, scores.argmax(axis=1)


# ----------------------------------------------------------------------------------------------- X1 25 x 10 protocol
def test_frame_indices(nFrames: int, n=N_TEST_SNIPPETS):
    """25 frames, equal temporal spacing (notes.txt:113-114), inside the reference's range [0, nFrames-1]
    (spatialModel.py:75).  Integer-exact: f_k = (k*(nFrames-1)) // (n-1)."""
    return [(k * (nFrames - 1)) // (n - 1) for k in range(n)]


def test_flow_starts(nFlows: int, L=VIDEO_INPUT_FLOW_COUNT, n=N_TEST_SNIPPETS):
    """25 stack starts, equally spaced over the reference's range [1, nFlows-L] (temporalModel.py:79)."""
    return [1 + (k * (nFlows - L - 1)) // (n - 1) for k in range(n)]


def ten_crop_params(h: int, w: int, crop=CROP_SIZE_TF):
    """(top, left, flip) of torchvision ten_crop order: tl, tr, bl, br, center, then the same five crops of the
    h-flipped image (notes.txt:114-115; torchvision/transforms/functional.py five_crop/ten_crop).  A crop at
    (i, j') of the flipped image covers original columns [w-crop-j', w-j') mirrored."""
    ci, cj = int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))
    five = [(0, 0), (0, w - crop), (h - crop, 0), (h - crop, w - crop), (ci, cj)]
    return [(i, j, 0) for (i, j) in five] + [(i, w - crop - j, 1) for (i, j) in five]


def video_snippets_spatial(store: OracleStore, name: str):
    """[250,3,224,224] tensor + index records (snippet-major, crop-minor)."""
    out, recs = [], []
    for f in test_frame_indices(store.n_frame_files(name)):
        img = store.frame(name, f)
        for (i, j, flip) in ten_crop_params(img.shape[0], img.shape[1]):
            out.append(apply_transform(img, i, j, flip, NORM_MEANS_TF, NORM_STDS_TF))
            recs.append((f, i, j, flip))
    return torch.stack(out), recs


def video_snippets_temporal(store: OracleStore, name: str, L=VIDEO_INPUT_FLOW_COUNT):
    """[250,2L,224,224]: one crop/flip per STACK (the 25x10 protocol crops the flow volume, notes.txt:107-108)."""
    out, recs = [], []
    mean, std = flow_norm_constants(1)
    nFlows = store.n_flow_files(name) // 2
    for s in test_flow_starts(nFlows, L):
        imgs = []
        for idx in range(s, s + L):
            imgs.append(store.flow_x(name, idx))
            imgs.append(store.flow_y(name, idx))
        for (i, j, flip) in ten_crop_params(imgs[0].shape[0], imgs[0].shape[1]):
            out.append(torch.cat([apply_transform(im, i, j, flip, mean, std) for im in imgs], dim=0))
            recs.append((s, i, j, flip))
    return torch.stack(out), recs


@torch.no_grad()
def video_consensus(model, snippets: torch.Tensor, batch=10):
    """Per-video consensus over the 250 snippets of one stream: descriptor mean in AverageMeter order
    (sequential fp32 sum / count, utils.py:167-171) and mean of softmax class scores (notes.txt:116)."""
    descMeter, scoreMeter = AverageMeter(), AverageMeter()
    all_desc, all_logits = [], []
    for b in range(0, snippets.shape[0], batch):
        fv, logits, _ = forward_eval(model, snippets[b:b + batch])
        probs = torch.softmax(logits, dim=1)
        for i in range(fv.shape[0]):
            descMeter.update(fv[i])
            scoreMeter.update(probs[i])
        all_desc.append(fv)
        all_logits.append(logits)
    return descMeter.avg, scoreMeter.avg, torch.cat(all_desc), torch.cat(all_logits)


def fuse_scores(score_s: torch.Tensor, score_t: torch.Tensor, w_s=1.0, w_t=1.0):
    """Weighted average of the two streams' class scores (notes.txt:121-124, 229-230)."""
    w_s = torch.tensor(w_s, dtype=torch.float32)
    w_t = torch.tensor(w_t, dtype=torch.float32)
    return (w_s * score_s + w_t * score_t) / (w_s + w_t)

"""Host-side helpers of the two-stream path, API-compatible with the reference's `Sheet03/utils.py` (same function
names, argument order, return values and error behaviour; line citations per function).  What differs is where
the pixels are touched: `getTransforms()` returns a *parameter sampler* that makes the reference's RNG draws on
the host, and the crop/flip/normalise arithmetic itself runs in the fused CUDA preprocess kernel
(`va_preprocess`).  Nothing here falls back to CPU image math.
"""
from __future__ import annotations

import csv
import os
import shutil
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .parameters import *  # noqa: F401,F403  (the reference star-imports its constants the same way, utils.py:10)
from . import parameters as _P


# ------------------------------------------------------------------------------------------------ filesystem helpers
def checkAndMakeDirectories(*args):
    """reference utils.py:14-26 -- create missing directories; returns, per argument, whether it already existed."""
    existed = []
    for path in args:
        present = os.path.exists(path)
        if not present:
            os.makedirs(path)
        existed.append(present)
    return existed


def makeCheckpoint(modelState, isBest, ckpLoc, bestModel):
    """reference utils.py:29-35 -- torch.save the state dict; copy to `bestModel` when it is the best so far."""
    torch.save(modelState, ckpLoc)
    if isBest:
        shutil.copyfile(ckpLoc, bestModel)


def getOneHot(label, nClasses):
    """reference utils.py:38-45 (unused by the reference; kept for API completeness). Labels are 1-based."""
    assert label < nClasses
    onehot = np.zeros((1, nClasses), dtype=np.float32)
    onehot[0, label - 1] = 1.0
    return torch.from_numpy(onehot)


def extractEveryNthFrame(videoLoc, N):
    """reference utils.py:51-69 -- decode a video with OpenCV and keep frames 0, N, 2N, ...  Host-side input preparation,
    as in the reference (UCF101's .avi files are MPEG-4 ASP, which NVDEC does not decode); pinned against the reference's
    own function on a committed video (tests/test_host_logic.py, oracle/make_golden_frames.py).  store.DeviceStore.from_videos
    uses the same decode and keeps the frames (and the TV-L1 flow of all pairs) in HBM without writing JPEGs."""
    import cv2
    if not (os.path.exists(videoLoc) and os.path.isfile(videoLoc)):
        raise ValueError("Video does not exist: %s" % (videoLoc))
    capture = cv2.VideoCapture(videoLoc)
    kept, k = [], 0
    while capture.isOpened():
        ok, frame = capture.read()
        if not ok:
            break
        if k % N == 0:
            kept.append(frame)
        k += 1
    capture.release()
    return kept


def videoInfo(line, mode):
    """reference utils.py:73-91 -- parse `Class/v_Class_gNN_cNN.avi[ label]`.
    Returns (video location, video name, action label str|None, action category, group, clip)."""
    label = None
    if mode == "train":
        loc, label = line.split(" ")
        label = label.strip()
    else:
        loc = line            # the test list carries no numeric label
    loc = loc.strip()
    category, fname = loc.split("/")
    category = category.strip()
    name = fname[:fname.rfind(".")]
    _, _, group, clip = name.split("_")
    return loc, name, label, category, group, clip


def convertVideosToFrames(rootDir, saveDir, videoListLoc, sampleRate=VIDEO_FRAME_SAMPLE_RATE, mode="train"):
    """reference utils.py:95-121 -- write every `sampleRate`-th frame of each listed video as <i>.jpg."""
    import cv2
    rootDir = rootDir if rootDir.endswith("/") else rootDir + "/"
    saveDir = saveDir if saveDir.endswith("/") else saveDir + "/"
    with open(videoListLoc, "r") as listing:
        for line in listing:
            loc, name, _, category, _, _ = videoInfo(line, mode)
            frameDir = saveDir + category + "/" + name
            if all(checkAndMakeDirectories(frameDir)):
                continue          # already converted
            for i, frame in enumerate(extractEveryNthFrame(rootDir + loc, sampleRate)):
                cv2.imwrite(frameDir + "/" + str(i) + ".jpg", frame)
    return


# ------------------------------------------------------------------------------------------------ transforms
class SnippetTransform:
    """What `getTransforms()` returns here: the reference's Compose[RandomCrop(224), RandomHorizontalFlip,
    ColorJitter(0,0,0,0), ToTensor, Normalize] reduced to (a) its RNG draws, made on the host in the exact order
    the installed torchvision makes them, and (b) the constants the CUDA kernel needs."""

    def __init__(self, crop: bool, flip: bool, means: Optional[Sequence[float]], stds: Optional[Sequence[float]],
                 jitter: Optional[Sequence[float]]):
        self.crop_size = 224 if crop else 0          # the reference hard-codes 224 whatever cropSize is (utils.py:143)
        self.flip = bool(flip)
        self.jitter = list(jitter) if jitter else None
        if self.jitter and any(float(v) != 0.0 for v in self.jitter):
            raise NotImplementedError("non-zero ColorJitter is not built (the reference uses COLOR_JITTERS=[0,0,0,0])")
        if bool(means) != bool(stds):
            means = stds = None                         # reference normalises only when both are given (utils.py:149)
        self.means = list(means) if means else None
        self.stds = list(stds) if stds else None

    def draw(self, h: int, w: int) -> Tuple[int, int, int]:
        """(crop top i, crop left j, flip) for one h x w image; consumes the global torch RNG like torchvision."""
        i = j = 0
        if self.crop_size:
            if h < self.crop_size or w < self.crop_size:
                raise ValueError(f"Required crop size {(self.crop_size, self.crop_size)} is larger than input image size {(h, w)}")
            if not (h == self.crop_size and w == self.crop_size):
                i = int(torch.randint(0, h - self.crop_size + 1, size=(1,)).item())
                j = int(torch.randint(0, w - self.crop_size + 1, size=(1,)).item())
        flip = 0
        if self.flip:
            flip = int(bool(torch.rand(1) < 0.5))
        if self.jitter is not None:
            torch.randperm(4)                           # ColorJitter draws its op order even as the identity
        return i, j, flip

    def norm_constants(self, n_channels: int, per_image_channels: int):
        """mean/std per stacked output channel.  3-channel images use the three constants; 1-channel (flow) images
        see only mean[0]/std[0] -- the zip semantics of the torchvision the reference was written for."""
        if self.means is None:
            return [0.0] * n_channels, [1.0] * n_channels
        if per_image_channels == 1:
            return [self.means[0]] * n_channels, [self.stds[0]] * n_channels
        reps = n_channels // per_image_channels
        return list(self.means[:per_image_channels]) * reps, list(self.stds[:per_image_channels]) * reps


def getTransforms(cropSize=CROP_SIZE_TF, hortizontalFlip=HORIZONTAL_FLIP_TF, normMeans=NORM_MEANS_TF,
                  normStds=NORM_STDS_TF, jitter=COLOR_JITTERS):
    """reference utils.py:137-151 (argument names, including the `hortizontalFlip` spelling, preserved)."""
    return SnippetTransform(bool(cropSize), bool(hortizontalFlip), normMeans, normStds, jitter)


def ten_crop_params(h: int, w: int, crop: int = CROP_SIZE_TF) -> List[Tuple[int, int, int]]:
    """(top, left, flip) for the 10 test crops in torchvision `ten_crop` order: tl, tr, bl, br, centre, then the
    same five of the h-flipped image (notes.txt:114-115).  A crop at column j' of the flipped image is the
    original's columns [w-crop-j', w-j') mirrored, which is how the preprocess kernel's flip flag reads them."""
    ci, cj = int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))
    five = [(0, 0), (0, w - crop), (h - crop, 0), (h - crop, w - crop), (ci, cj)]
    return [(i, j, 0) for i, j in five] + [(i, w - crop - j, 1) for i, j in five]


def test_frame_indices(nFrames: int, n: int = _P.N_TEST_SNIPPETS) -> List[int]:
    """25 equally spaced frame ids in the reference's range [0, nFrames-1] (spatialModel.py:75; notes.txt:113)."""
    return [(k * (nFrames - 1)) // (n - 1) for k in range(n)]


def test_flow_starts(nFlows: int, L: int = VIDEO_INPUT_FLOW_COUNT, n: int = _P.N_TEST_SNIPPETS) -> List[int]:
    """25 equally spaced stack starts in the reference's range [1, nFlows-L] (temporalModel.py:79).  A video with fewer
    than L+1 flow pairs has no valid start: the reference's `random.randint(1, nFlows - L)` raises ValueError there, and a
    start < 1 would address an image before the video's first flow image in the store."""
    if nFlows - L < 1:
        raise ValueError("video has %d flow pairs; a stack of %d needs at least %d" % (nFlows, L, L + 1))
    return [1 + (k * (nFlows - L - 1)) // (n - 1) for k in range(n)]


# ------------------------------------------------------------------------------------------------ loader
def check_index_rows(rows, n_images: int, image_shape, crop: int = 224) -> None:
    """Host-side validation of index-table rows (image id, crop top, crop left, flip) before they are uploaded: the
    preprocess kernels address `images + id * image_bytes` unchecked (va_preprocess is not told the store's size), so an
    id outside [0, n_images) or a crop window outside the image would read foreign memory.  The reference fails in the
    same situations with an IOError from Image.open / a ValueError from RandomCrop."""
    r = np.asarray(rows).reshape(-1, 4)
    if r.size == 0:
        return
    h, w = int(image_shape[0]), int(image_shape[1])
    if r[:, 0].min() < 0 or r[:, 0].max() >= n_images:
        raise IndexError("index table: image id range [%d, %d] outside the store's %d images"
                         % (r[:, 0].min(), r[:, 0].max(), n_images))
    if r[:, 1].min() < 0 or r[:, 1].max() + crop > h or r[:, 2].min() < 0 or r[:, 2].max() + crop > w:
        raise ValueError("index table: a %dx%d crop window leaves the %dx%d image" % (crop, crop, h, w))
    if ((r[:, 3] != 0) & (r[:, 3] != 1)).any():
        raise ValueError("index table: flip must be 0 or 1")


class SnippetBatch:
    """A preprocessed batch resident in HBM: bf16 NHWC [B,224,224,c_pad] plus the index table that produced it."""

    def __init__(self, nhwc: torch.Tensor, table: torch.Tensor):
        self.nhwc, self.table = nhwc, table

    def size(self, dim=None):
        return self.nhwc.shape[0] if dim == 0 else self.nhwc.shape

    def __len__(self):
        return self.nhwc.shape[0]


def getDataLoader(dataset, batchSize=TEMPORAL_BATCH_SIZE, nWorkers=NWORKERS_LOADER, shuffle=SHUFFLE_LOADER):
    """reference utils.py:125-133.  Returns a torch DataLoader whose sampler and per-item RNG draws are the
    reference's (so shuffles and crops replay bit-for-bit under the same seeds) but whose items are index-table
    rows; the collate step uploads the table and runs the fused CUDA preprocess kernel, yielding
    (SnippetBatch, LongTensor labels, tuple names).  `nWorkers` is accepted for compatibility and ignored: the
    reference's workers parallelise JPEG decode + PIL transforms, which here is one kernel launch."""
    from torch.utils.data import DataLoader

    class _IndexView(torch.utils.data.Dataset):
        def __len__(self):
            return len(dataset)

        def __getitem__(self, index):
            return dataset.sample_indices(index)

    def _collate(items):
        rows = np.stack([it[0] for it in items]).astype(np.int32)
        dataset.check_rows(rows)
        labels = torch.tensor([it[1] for it in items], dtype=torch.int64)
        names = tuple(it[2] for it in items)
        table = torch.from_numpy(rows).pin_memory().cuda(non_blocking=True)
        return SnippetBatch(dataset.preprocess_table(table), table), labels, names

    world, rank = _dist_world_rank()
    if world > 1:
        # one process per GPU: every rank walks its own 1/world of each epoch's permutation (the reference is
        # single-process; its DataParallel wrapper would split each batch across devices, spatialModel.py:133)
        sampler = RankShardSampler(len(dataset), shuffle, rank, world)
        loader = DataLoader(dataset=_IndexView(), batch_size=batchSize, sampler=sampler, num_workers=0, collate_fn=_collate)
    else:
        loader = DataLoader(dataset=_IndexView(), batch_size=batchSize, shuffle=shuffle, num_workers=0, collate_fn=_collate)
    loader.snippet_dataset = dataset
    return loader


def _dist_world_rank():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


class RankShardSampler(torch.utils.data.Sampler):
    """Rank r of R takes positions r, r+R, r+2R, ... of the epoch's permutation (padded by wrapping so every rank sees
    the same number of batches: the gradient all-reduce is a collective).  The permutation is drawn by rank 0 from the
    global torch RNG -- one draw per epoch, like torch's RandomSampler -- and broadcast, so ranks seeded differently
    (different crops and dropout masks, as they should be) still partition the data instead of repeating it."""

    def __init__(self, n_items: int, shuffle: bool, rank: int, world: int):
        self.n, self.shuffle, self.rank, self.world = n_items, bool(shuffle), rank, world
        self.per_rank = (n_items + world - 1) // world

    def __len__(self):
        return self.per_rank

    def __iter__(self):
        import torch.distributed as dist
        if self.shuffle:
            perm = torch.randperm(self.n) if self.rank == 0 else torch.empty(self.n, dtype=torch.int64)
            if dist.get_backend() == "nccl":
                dev_perm = perm.cuda()
                dist.broadcast(dev_perm, src=0)
                perm = dev_perm.cpu()
            else:
                dist.broadcast(perm, src=0)
        else:
            perm = torch.arange(self.n)
        total = self.per_rank * self.world
        if total > self.n:
            perm = torch.cat([perm, perm[:total - self.n]])
        return iter(perm[self.rank:total:self.world].tolist())


# ------------------------------------------------------------------------------------------------ consensus
class AverageMeter(object):
    """reference utils.py:154-171 -- running sum / count / avg (works on floats and tensors alike)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.avg, self.sum, self.count = 0, 0, 0, 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class DeviceVideoDict:
    """Device-resident form of the reference's `trainDict` / `testDict` (videoName -> (AverageMeter, label),
    spatialModel.py:131-132,183-188): one fp32 sum row and one count per video in HBM, updated by
    `va_consensus_update` in the reference's summation order.  Behaves like the dict for the reference's readers:
    `in`, `keys()`, `len()`, and `d[name] -> (meter, label)` where `meter.avg` is the mean descriptor."""

    def __init__(self, descriptorDim: int, capacity: int = 1024):
        self.dim = descriptorDim
        self.slot: Dict[str, int] = {}
        self.labels: Dict[str, torch.Tensor] = {}
        self.sum = torch.zeros((capacity, descriptorDim), dtype=torch.float32, device="cuda")
        self.count = torch.zeros((capacity,), dtype=torch.int32, device="cuda")

    def _grow(self, need: int):
        cap = self.sum.shape[0]
        if need <= cap:
            return
        new_cap = max(need, 2 * cap)
        s = torch.zeros((new_cap, self.dim), dtype=torch.float32, device="cuda")
        c = torch.zeros((new_cap,), dtype=torch.int32, device="cuda")
        s[:cap].copy_(self.sum)
        c[:cap].copy_(self.count)
        self.sum, self.count = s, c

    def update_batch(self, videoNames: Sequence[str], labels, featureVectors: torch.Tensor):
        from . import _lib
        ids = []
        for i, name in enumerate(videoNames):
            if name not in self.slot:
                self.slot[name] = len(self.slot)
                self.labels[name] = labels[i]
            ids.append(self.slot[name])
        self._grow(len(self.slot))
        vid = torch.tensor(ids, dtype=torch.int32).pin_memory().cuda(non_blocking=True)
        fv = featureVectors.contiguous()
        _lib.check(_lib.load().va_consensus_update(_lib.ptr(self.sum), _lib.ptr(self.count), _lib.ptr(vid), _lib.ptr(fv),
                                                   len(ids), self.dim, _lib.stream_ptr()), "va_consensus_update")

    def __contains__(self, name):
        return name in self.slot

    def __len__(self):
        return len(self.slot)

    def keys(self):
        return self.slot.keys()

    def averages(self) -> torch.Tensor:
        """[n_videos, dim] mean descriptors (sum / count, the AverageMeter definition) in insertion order."""
        n = len(self.slot)
        return self.sum[:n] / self.count[:n].to(torch.float32).unsqueeze(1)

    def __getitem__(self, name):
        k = self.slot[name]
        meter = AverageMeter()
        meter.sum = self.sum[k]
        meter.count = int(self.count[k].item())
        meter.avg = meter.sum / meter.count
        return meter, self.labels[name]


def saveVideoDescriptors(videoDescDict, csvLoc, gpu=False):
    """reference utils.py:174-195 -- one CSV row per video: name,label,<descriptor floats>.  Byte-compatible with the
    reference writer (csv module float repr, '\\r\\n' row terminator), so `combineDescriptors` reads either."""
    try:
        os.remove(csvLoc)
    except OSError:
        pass
    if isinstance(videoDescDict, DeviceVideoDict):
        avgs = videoDescDict.averages().cpu().numpy().astype(float)      # one D2H for the whole table
        rows = ((name, videoDescDict.labels[name], avgs[k]) for name, k in videoDescDict.slot.items())
    else:
        rows = ((name, pair[1], pair[0].avg.detach().cpu().numpy().astype(float)) for name, pair in videoDescDict.items())
    with open(csvLoc, "a") as csvFile:
        writer = csv.writer(csvFile, delimiter=",")
        for name, label, desc in rows:
            label_np = label.detach().cpu().numpy() if isinstance(label, torch.Tensor) else np.asarray(label)
            csvFile.write(name + "," + str(label_np) + ",")
            writer.writerow(desc)


def savePerformance(precision, loss, csvLoc):
    """reference utils.py:198-205 -- append `precision,loss` for the epoch."""
    with open(csvLoc, "a") as csvFile:
        csvFile.write(str(precision) + "," + str(loss) + "\n")
    return

"""Python call layer over the C-ABI (include/va_b200.h).  torch is used only to own device memory and streams.

Every function launches hand-written sm_100a kernels from libva_b200.so on the current CUDA stream and raises
`VAError` on failure -- nothing here computes with torch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import VAError, check, ptr, stream_ptr

STREAM_SPATIAL, STREAM_TEMPORAL = 0, 1
CROP = 224


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise VAError("libva_b200 operates on CUDA tensors only (no CPU fallback)")
        if t is not None and not t.is_contiguous():
            raise VAError("libva_b200 needs contiguous tensors")


def device_info():
    lib = _lib.load()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(lib.va_device_info(C.byref(a), C.byref(b), C.byref(c)), "va_device_info")
    return {"sm_count": a.value, "cc": (b.value, c.value)}


def conv2d_nhwc(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, *, relu=True, pool=False, force_bn=0,
                force_r=0) -> torch.Tensor:
    """x bf16 [n,H,W,cin_pad]; w fp32 OIHW [cout,cin,ks,ks]; -> bf16 [n,H',W',cout] (H'=H/2 when pool)."""
    _need_cuda(x, w, bias)
    assert x.dtype == torch.bfloat16 and w.dtype == torch.float32 and bias.dtype == torch.float32
    n, H, W, cin_pad = x.shape
    cout, cin, ks, _ = w.shape
    sh = 1 if pool else 0
    y = torch.empty((n, H >> sh, W >> sh, cout), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().va_conv2d_nhwc(ptr(x), n, H, W, cin, cin_pad, ptr(w), ptr(bias), cout, ks, int(relu), int(pool),
                                     ptr(y), force_bn, force_r, stream_ptr()), "va_conv2d_nhwc")
    return y


def linear(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, *, relu=True, out_f32=False, force_bn=0) -> torch.Tensor:
    """x bf16 [n,in]; w fp32 [out,in] -> bf16 (or fp32) [n,out]."""
    _need_cuda(x, w, bias)
    assert x.dtype == torch.bfloat16 and w.dtype == torch.float32
    n, fin = x.shape
    fout = w.shape[0]
    y = torch.empty((n, fout), dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    check(_lib.load().va_linear(ptr(x), n, fin, ptr(w), ptr(bias), fout, int(relu), ptr(None if out_f32 else y),
                                ptr(y if out_f32 else None), force_bn, stream_ptr()), "va_linear")
    return y


def preprocess(images: torch.Tensor, image_shape: Sequence[int], index_table: torch.Tensor, mean: Sequence[float],
               std: Sequence[float], *, c_pad: int = 16, reference_layout: bool = False, crop: int = CROP,
               image_bytes: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K1.  images: u8 store (flat); image_shape (h, w, c); index_table int32 [n, planes, 4]
    = (image id, crop top, crop left, flip).  Returns bf16 NHWC [n,crop,crop,c_pad], or with
    reference_layout=True the reference's fp32 NCHW tensor [n, planes*c, crop, crop] (bit-exact)."""
    _need_cuda(images, index_table)
    assert images.dtype == torch.uint8 and index_table.dtype == torch.int32 and index_table.dim() == 3
    h, w, c = image_shape
    n, planes, four = index_table.shape
    assert four == 4
    nch = planes * c
    assert len(mean) == nch and len(std) == nch, "one mean/std per stacked channel"
    if image_bytes is None:
        image_bytes = h * w * c
    shape = (n, nch, crop, crop) if reference_layout else (n, crop, crop, c_pad)
    dtype = torch.float32 if reference_layout else torch.bfloat16
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=images.device)
    else:                                   # caller-owned buffer (e.g. reused across steps on a side stream)
        _need_cuda(out)
        assert out.dtype == dtype and out.numel() >= int(torch.Size(shape).numel()), (out.shape, shape)
        out = out.view(-1)[:int(torch.Size(shape).numel())].view(shape)
    fm = (C.c_float * nch)(*mean)
    fs = (C.c_float * nch)(*std)
    check(_lib.load().va_preprocess(ptr(images), image_bytes, h, w, c, ptr(index_table), n, planes, crop, fm, fs, c_pad,
                                    1 if reference_layout else 0, ptr(out), stream_ptr()), "va_preprocess")
    return out


def conv1_fused(images: torch.Tensor, image_shape: Sequence[int], index_table: torch.Tensor, mean: Sequence[float],
                std: Sequence[float], w: torch.Tensor, bias: torch.Tensor, *, image_bytes: Optional[int] = None) -> torch.Tensor:
    """K1+conv1_1 fused (csrc/va_conv1_fused.cu): store + index table -> bf16 NHWC [n,224,224,64] =
    ReLU(conv3x3(normalised crop, w) + bias), w fp32 OIHW [64, planes*c, 3, 3]."""
    _need_cuda(images, index_table, w, bias)
    assert images.dtype == torch.uint8 and index_table.dtype == torch.int32 and index_table.dim() == 3
    h, wd, c = image_shape
    n, planes, four = index_table.shape
    assert four == 4 and tuple(w.shape) == (64, planes * c, 3, 3) and w.dtype == torch.float32
    if image_bytes is None:
        image_bytes = h * wd * c
    nch = planes * c
    fm = (C.c_float * nch)(*mean)
    fs = (C.c_float * nch)(*std)
    y = torch.empty((n, CROP, CROP, 64), dtype=torch.bfloat16, device=images.device)
    check(_lib.load().va_conv1_fused(ptr(images), image_bytes, h, wd, c, ptr(index_table), n, planes, fm, fs, ptr(w),
                                     ptr(bias), ptr(y), stream_ptr()), "va_conv1_fused")
    return y


def svm_fit(X: torch.Tensor, class_index: torch.Tensor, n_classes: int, *, C_reg: float = 1.0, bias: float = 1.0,
            tol: float = 1e-4, max_iter: int = 1000):
    """LinearSVC().fit (reference combinedModel.py:34-35) on the device: one-vs-rest L2-regularised squared-hinge SVM by
    dual coordinate descent in fp64.  X fp64 [V, F] (device), class_index int32 [V] in [0, n_classes).  Returns
    (coef [P, F], intercept [P], epochs [P]) device tensors, P = 1 for two classes else n_classes."""
    _need_cuda(X, class_index)
    assert X.dtype == torch.float64 and X.dim() == 2 and X.is_contiguous()
    assert class_index.dtype == torch.int32 and class_index.numel() == X.shape[0]
    V, F = X.shape
    P = 1 if n_classes == 2 else n_classes
    dev = X.device
    coef = torch.empty((P, F), dtype=torch.float64, device=dev)
    intercept = torch.empty((P,), dtype=torch.float64, device=dev)
    epochs = torch.empty((P,), dtype=torch.int32, device=dev)
    work = torch.empty(((P + 1) * max(V, 1) + (P * max(V, 1) + 1) // 2,), dtype=torch.float64, device=dev)
    check(_lib.load().va_svm_fit(ptr(X), ptr(class_index), V, F, n_classes, C_reg, bias, tol, max_iter, ptr(coef),
                                 ptr(intercept), ptr(epochs), ptr(work), stream_ptr()), "va_svm_fit")
    return coef, intercept, epochs


def fuse(desc_s, desc_t, score_s, score_t, video_offsets: torch.Tensor, *, svm_w=None, svm_b=None, w_s=1.0, w_t=1.0,
         out: Optional[dict] = None) -> dict:
    """K4.  Per-video consensus (sequential-sum means) + late fusion.  Returns dict with video_desc [V,2D],
    video_scores [V,C], score_pred [V], svm_scores [V,C] f64, svm_pred [V] (those that apply)."""
    _need_cuda(desc_s, desc_t, score_s, score_t, video_offsets, svm_w, svm_b)
    assert video_offsets.dtype == torch.int32
    V = video_offsets.numel() - 1
    ref = desc_s if desc_s is not None else score_s
    dev = ref.device
    D = desc_s.shape[1] if desc_s is not None else 1
    Cn = score_s.shape[1] if score_s is not None else 1
    Csvm = 0
    if svm_w is not None:
        # the SVM has one row per class it was FITTED on (LinearSVC: len(np.unique(y))), independent of the score width
        if desc_s is None or desc_t is None:
            raise VAError("ops.fuse: SVM scoring needs both descriptor arrays")
        Csvm = int(svm_w.shape[0])
        if svm_w.dim() != 2 or svm_w.shape[1] != 2 * D or svm_b is None or svm_b.numel() != Csvm:
            raise VAError(f"ops.fuse: svm_w {tuple(svm_w.shape)} / svm_b {None if svm_b is None else tuple(svm_b.shape)} "
                          f"do not match descriptors of width 2*{D}")
    res = out if out is not None else {}
    if desc_s is not None and "video_desc" not in res:
        res["video_desc"] = torch.empty((V, 2 * D), dtype=torch.float32, device=dev)
    if score_s is not None and "video_scores" not in res:
        res["video_scores"] = torch.empty((V, Cn), dtype=torch.float32, device=dev)
        res["score_pred"] = torch.empty((V,), dtype=torch.int32, device=dev)
    if svm_w is not None and "svm_scores" not in res:
        assert svm_w.dtype == torch.float64 and svm_b.dtype == torch.float64
        res["svm_scores"] = torch.empty((V, Csvm), dtype=torch.float64, device=dev)
        res["svm_pred"] = torch.empty((V,), dtype=torch.int32, device=dev)
    if svm_w is not None and res.get("svm_scores") is not None and tuple(res["svm_scores"].shape[1:]) != (Csvm,):
        raise VAError(f"ops.fuse: svm_scores buffer {tuple(res['svm_scores'].shape)} does not have {Csvm} SVM classes")
    check(_lib.load().va_fuse(ptr(desc_s), ptr(desc_t), ptr(score_s), ptr(score_t), ptr(video_offsets), V, D, Cn, Csvm,
                              ptr(svm_w), ptr(svm_b), float(w_s), float(w_t), ptr(res.get("video_desc")),
                              ptr(res.get("video_scores")), ptr(res.get("score_pred")), ptr(res.get("svm_scores")),
                              ptr(res.get("svm_pred")), stream_ptr()), "va_fuse")
    return res


def svm_decision(X: torch.Tensor, W: torch.Tensor, b: torch.Tensor):
    """LinearSVC decision function + predict on fp64 rows: X [V,F] f64, W [P,F] f64, b [P] f64 -> (scores [V,P] f64,
    pred [V] i32 = first maximum)."""
    _need_cuda(X, W, b)
    assert X.dtype == torch.float64 and W.dtype == torch.float64 and b.dtype == torch.float64
    V, F = X.shape
    P = W.shape[0]
    if W.shape[1] != F or b.numel() != P:
        raise VAError(f"svm_decision: X {tuple(X.shape)}, W {tuple(W.shape)}, b {tuple(b.shape)} do not match")
    scores = torch.empty((V, P), dtype=torch.float64, device=X.device)
    pred = torch.empty((V,), dtype=torch.int32, device=X.device)
    check(_lib.load().va_svm_decision(ptr(X), V, F, ptr(W), ptr(b), P, ptr(scores), ptr(pred), stream_ptr()), "va_svm_decision")
    return scores, pred


def synth_fill(images: torch.Tensor, image_shape: Sequence[int], n_images: int, *, seed: int, first_id: int = 0,
               image_bytes: Optional[int] = None) -> torch.Tensor:
    _need_cuda(images)
    h, w, c = image_shape
    if image_bytes is None:
        image_bytes = h * w * c
    assert images.numel() >= n_images * image_bytes
    check(_lib.load().va_synth_fill(ptr(images), image_bytes, n_images, h, w, c, seed & 0xFFFFFFFF,
                                    first_id & 0xFFFFFFFF, stream_ptr()), "va_synth_fill")
    return images


# reference state_dict order (torchvision vgg16 features + the swapped classifier, spatialModel.py:141-152)
STATE_DICT_KEYS = [f"features.{i}.{p}" for i in (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28) for p in ("weight", "bias")] + \
                  [f"classifier.{i}.{p}" for i in (0, 3, 6, 9) for p in ("weight", "bias")]


class StreamNet:
    """One VGG16-D stream (spatial: 3 input channels, temporal: 2L=20) resident on the current CUDA device."""

    def __init__(self, stream_kind: int, in_channels: int, n_classes: int = 101, desc_dim: int = 256, max_batch: int = 64,
                 precision: str = "bf16"):
        """precision: "bf16" (throughput path: bf16 storage, fp32 accumulate) or "fp32" (parity mode: bf16x3 slices,
        six cross terms per layer; ~6x the FLOPs and activation memory -- use a small max_batch)."""
        if precision not in ("bf16", "fp32"):
            raise VAError(f"precision {precision!r}: expected 'bf16' or 'fp32'")
        lib = _lib.load()
        h = C.c_void_p()
        check(lib.va_create_ex(C.byref(h), stream_kind, in_channels, n_classes, desc_dim, max_batch,
                               1 if precision == "fp32" else 0), "va_create_ex")
        self._h = h
        self.precision = precision
        self.stream_kind, self.in_channels = stream_kind, in_channels
        self.n_classes, self.desc_dim, self.max_batch = n_classes, desc_dim, max_batch
        self.c_pad = lib.va_input_channels_padded(h)
        self._keep = None

    def load_state_dict(self, state_dict) -> None:
        """Accepts the reference checkpoint's "model" dict (keys may carry the DataParallel "module." prefix,
        spatialModel.py:133,256-261)."""
        sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in state_dict.items()}
        missing = [k for k in STATE_DICT_KEYS if k not in sd]
        if missing:
            raise VAError(f"state_dict is missing {missing[:3]}...")
        dev = torch.device("cuda", torch.cuda.current_device())
        tensors = [sd[k].detach().to(device=dev, dtype=torch.float32).contiguous() for k in STATE_DICT_KEYS]
        arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        check(_lib.load().va_load_weights(self._h, arr, len(tensors), stream_ptr()), "va_load_weights")
        torch.cuda.current_stream().synchronize()   # packing done; fp32 staging copies may be freed

    def pack_input(self, ip: torch.Tensor) -> torch.Tensor:
        """Reference-layout snippets (float [n,C,H,W], what the reference feeds as `ip`) -> this handle's network
        input layout (bf16 NHWC, zero-padded channels; six slice blocks per channel for precision "fp32")."""
        ip = ip.detach().to(device="cuda", dtype=torch.float32).contiguous()
        n, c, hh, ww = ip.shape
        x = torch.empty((n, hh, ww, self.c_pad), dtype=torch.bfloat16, device="cuda")
        lib = _lib.load()
        if self.precision == "fp32":
            check(lib.va_pack_input_nchw_split6(ptr(ip), n, c, hh, ww, self.c_pad, ptr(x), stream_ptr()),
                  "va_pack_input_nchw_split6")
        else:
            check(lib.va_pack_input_nchw(ptr(ip), n, c, hh, ww, self.c_pad, ptr(x), stream_ptr()), "va_pack_input_nchw")
        return x

    def forward(self, x_nhwc: torch.Tensor, *, want_logits=True, want_probs=True, want_pred=True):
        """x bf16 [n,224,224,c_pad] -> (descriptors [n,D] f32, logits [n,C] f32, probs [n,C] f32, pred [n] i32)."""
        _need_cuda(x_nhwc)
        assert x_nhwc.dtype == torch.bfloat16 and tuple(x_nhwc.shape[1:]) == (CROP, CROP, self.c_pad), x_nhwc.shape
        n = x_nhwc.shape[0]
        dev = x_nhwc.device
        desc = torch.empty((n, self.desc_dim), dtype=torch.float32, device=dev)
        logits = torch.empty((n, self.n_classes), dtype=torch.float32, device=dev) if want_logits else None
        probs = torch.empty((n, self.n_classes), dtype=torch.float32, device=dev) if want_probs else None
        pred = torch.empty((n,), dtype=torch.int32, device=dev) if want_pred else None
        check(_lib.load().va_forward(self._h, ptr(x_nhwc), n, ptr(desc), ptr(logits), ptr(probs), ptr(pred),
                                     stream_ptr()), "va_forward")
        return desc, logits, probs, pred

    def forward_store(self, images: torch.Tensor, image_shape: Sequence[int], index_table: torch.Tensor,
                      mean: Sequence[float], std: Sequence[float], *, image_bytes: Optional[int] = None, want_logits=True,
                      want_probs=True, want_pred=True):
        """The forward fed from the image store: the snippet transform is gathered straight into the first
        convolution (va_forward_store), no preprocessed tensor is materialised.  Same returns as forward()."""
        _need_cuda(images, index_table)
        assert images.dtype == torch.uint8 and index_table.dtype == torch.int32 and index_table.dim() == 3
        h, wd, c = image_shape
        n, planes, four = index_table.shape
        assert four == 4
        nch = planes * c
        assert len(mean) == nch and len(std) == nch, "one mean/std per stacked channel"
        if image_bytes is None:
            image_bytes = h * wd * c
        dev = images.device
        desc = torch.empty((n, self.desc_dim), dtype=torch.float32, device=dev)
        logits = torch.empty((n, self.n_classes), dtype=torch.float32, device=dev) if want_logits else None
        probs = torch.empty((n, self.n_classes), dtype=torch.float32, device=dev) if want_probs else None
        pred = torch.empty((n,), dtype=torch.int32, device=dev) if want_pred else None
        fm = (C.c_float * nch)(*mean)
        fs = (C.c_float * nch)(*std)
        check(_lib.load().va_forward_store(self._h, ptr(images), image_bytes, h, wd, c, ptr(index_table), n, planes, fm, fs,
                                           ptr(desc), ptr(logits), ptr(probs), ptr(pred), stream_ptr()), "va_forward_store")
        return desc, logits, probs, pred

    @staticmethod
    def forward_store_supported(images: torch.Tensor, image_shape: Sequence[int], planes: int) -> bool:
        """What the fused gather + conv1_1 kernel needs from the store (csrc/va_conv1_fused.cu): one RGB image or a stack
        of 1-channel images per snippet, at least 224 x 224, 16-byte aligned with a 16-byte multiple per image (the loader
        copies 16-byte blocks), and a plane count whose strips fit in shared memory."""
        h, wd, c = image_shape
        if not ((c == 3 and planes == 1) or (c == 1 and 1 <= planes <= 23)):
            return False
        return h >= CROP and wd >= CROP and (h * wd * c) % 16 == 0 and images.data_ptr() % 16 == 0

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().va_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

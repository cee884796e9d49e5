"""JPEG files -> device image store, decoded by libva_b200's CUDA kernels.

Replaces `Image.open(frameDir + str(frameName) + FRAME_EXTN)` / `Image.open(frame)` of the reference datasets
(Sheet03/spatialModel.py:76-79, temporalModel.py:85-88) for the files `cv2.imwrite` writes (utils.py:116-120).  The host
only walks the marker segments in front of the scan (SOI, DQT, DHT, SOF0, DRI, SOS -- a few hundred bytes per file) and
derives the canonical Huffman decoding tables; entropy decoding, IDCT, chroma upsampling and colour conversion run on the
GPU (csrc/va_jpeg.cu) and reproduce Pillow's pixels bit for bit.  There is no CPU decode fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import VAError, check, ptr, stream_ptr

_ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
                    21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53,
                    60, 61, 54, 47, 55, 62, 63], dtype=np.int64)

# mirrors va_jpeg_image / va_jpeg_huff of include/va_b200.h
IMAGE_DTYPE = np.dtype([("scan_offset", "<u8"), ("out_offset", "<u8"), ("scan_bytes", "<u4"), ("restart_interval", "<u4"),
                        ("width", "<u2"), ("height", "<u2"), ("n_comp", "u1"), ("sampling", "u1"), ("qt", "u1", (3,)),
                        ("dc", "u1", (3,)), ("ac", "u1", (3,)), ("pad", "u1", (9,))])
HUFF_DTYPE = np.dtype([("maxcode", "<i4", (18,)), ("valoffset", "<i4", (17,)), ("huffval", "u1", (256,)), ("pad", "<i4")])
assert IMAGE_DTYPE.itemsize == 48 and HUFF_DTYPE.itemsize == 400


MAX_HUFF_TABLES_PER_CALL = 8     # kJpegMaxTables of csrc/va_jpeg.cu: tables live in shared memory


class JpegFormatError(VAError):
    """The file is not an 8-bit baseline sequential Huffman JPEG with 1 component or YCbCr 4:2:0 / 4:4:4."""


def _derive_huffman(bits: Sequence[int], huffval: bytes) -> np.ndarray:
    """Canonical code assignment (ITU-T T.81 Annex C) -> decoding table: maxcode[l], valoffset[l], symbols."""
    t = np.zeros((), dtype=HUFF_DTYPE)
    t["maxcode"][:] = -1
    code, p = 0, 0
    for l in range(1, 17):
        n = bits[l - 1]
        if n:
            t["valoffset"][l] = p - code
            p += n
            code += n
            t["maxcode"][l] = code - 1
        code <<= 1
    t["maxcode"][17] = 0xFFFFF
    if p > 256 or p != len(huffval):
        raise JpegFormatError("corrupt Huffman table")
    t["huffval"][:p] = np.frombuffer(huffval, dtype=np.uint8)
    return t


class ParsedJpeg:
    __slots__ = ("width", "height", "n_comp", "sampling", "qtables", "htables", "comp_tables", "restart_interval",
                 "scan_offset")


_HEADER_CACHE: Dict[tuple, tuple] = {}
_LAST_HEADER: List[Optional[tuple]] = [None]


def parse_header(data: bytes) -> ParsedJpeg:
    """Walk the marker segments up to the first scan.  Files of one encoder / size / quality share every table, so the
    table work is memoised on the raw DQT / DHT / SOF / SOS segments (a loader parses the same header thousands of times)."""
    if len(data) < 4 or data[0] != 0xFF or data[1] != 0xD8:
        raise JpegFormatError("not a JPEG file (no SOI marker)")
    last = _LAST_HEADER[0]
    if last is not None and data.startswith(last[0]):       # same bytes up to the scan as the previous file: same header
        return last[1]
    segs = []
    restart_interval = 0
    i, n = 2, len(data)
    while i + 4 <= n:
        if data[i] != 0xFF:
            raise JpegFormatError("marker expected at byte %d" % i)
        while i + 1 < n and data[i + 1] == 0xFF:
            i += 1
        m = data[i + 1]
        i += 2
        if m == 0x01 or 0xD0 <= m <= 0xD8:
            continue
        L = (data[i] << 8) | data[i + 1]
        if m in (0xDB, 0xC4, 0xC0, 0xC1, 0xDA):
            segs.append((m, bytes(data[i + 2:i + L])))
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise JpegFormatError("SOF%d: progressive / lossless / arithmetic JPEGs are not what cv2.imwrite writes" % (m - 0xC0))
        elif m == 0xDD:
            restart_interval = (data[i + 2] << 8) | data[i + 3]
        if m == 0xDA:
            key = tuple(segs)
            hit = _HEADER_CACHE.get(key)
            if hit is None:
                hit = _interpret_segments(segs)
                if len(_HEADER_CACHE) < 4096:
                    _HEADER_CACHE[key] = hit
            out = ParsedJpeg()
            out.width, out.height, out.n_comp, out.sampling, out.qtables, out.htables = hit
            out.restart_interval = restart_interval
            out.scan_offset = i + L
            _LAST_HEADER[0] = (bytes(data[:out.scan_offset]), out)
            return out
        i += L
    raise JpegFormatError("no scan (SOS) found")


def _interpret_segments(segs) -> tuple:
    qt: Dict[int, bytes] = {}
    ht: Dict[Tuple[int, int], bytes] = {}
    comps: List[Tuple[int, int, int, int]] = []
    width = height = 0
    for m, seg in segs:
        if m == 0xDB:
            j = 0
            while j < len(seg):
                pq, tq = seg[j] >> 4, seg[j] & 15
                size = 128 if pq else 64
                vals = np.frombuffer(seg[j + 1:j + 1 + size], dtype=">u2" if pq else np.uint8).astype(np.uint16)
                nat = np.zeros(64, dtype=np.uint16)
                nat[_ZIGZAG] = vals
                qt[tq] = nat.tobytes()
                j += 1 + size
        elif m == 0xC4:
            j = 0
            while j < len(seg):
                tc, th = seg[j] >> 4, seg[j] & 15
                bits = seg[j + 1:j + 17]
                cnt = sum(bits)
                ht[(tc, th)] = bytes(bits) + bytes(seg[j + 17:j + 17 + cnt])
                j += 17 + cnt
        elif m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise JpegFormatError("sample precision %d (8 expected)" % seg[0])
            height, width = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            comps = [(seg[6 + 3 * k], seg[7 + 3 * k] >> 4, seg[7 + 3 * k] & 15, seg[8 + 3 * k]) for k in range(seg[5])]
        elif m == 0xDA:
            if not comps:
                raise JpegFormatError("SOS before SOF")
            if seg[0] != len(comps):
                raise JpegFormatError("multi-scan (non-interleaved) files are not supported")
            sel = {seg[1 + 2 * k]: seg[2 + 2 * k] for k in range(seg[0])}
            samp = [(c[1], c[2]) for c in comps]
            if len(comps) == 1:
                n_comp, sampling = 1, 0
            elif len(comps) == 3 and samp == [(2, 2), (1, 1), (1, 1)]:
                n_comp, sampling = 3, 2
            elif len(comps) == 3 and samp == [(1, 1), (1, 1), (1, 1)]:
                n_comp, sampling = 3, 1
            else:
                raise JpegFormatError("component layout %r is not supported (1 component, 4:4:4 or 4:2:0)" % (samp,))
            try:
                qtables = [qt[c[3]] for c in comps]
                htables = [(ht[(0, sel[c[0]] >> 4)], ht[(1, sel[c[0]] & 15)]) for c in comps]
            except KeyError as e:
                raise JpegFormatError("missing table %r" % (e.args[0],)) from None
            return width, height, n_comp, sampling, qtables, htables
    raise JpegFormatError("no scan (SOS) found")


class JpegBatch:
    """Headers of a batch of files, with the distinct quantisation / Huffman tables pooled."""

    def __init__(self, files: Sequence[bytes], out_offsets: Sequence[int]):
        assert len(files) == len(out_offsets)
        self.n = len(files)
        q_index: Dict[bytes, int] = {}
        h_index: Dict[bytes, int] = {}
        h_tabs: List[np.ndarray] = []
        table_rows: Dict[int, tuple] = {}          # id(header cache entry) -> per-component table indices
        rows = []
        pos = 0
        self.sizes = []
        for f, off in zip(files, out_offsets):
            hd = parse_header(f)
            key = (id(hd.qtables), id(hd.htables))
            idx = table_rows.get(key)
            if idx is None:
                qi, di, ai = [0, 0, 0], [0, 0, 0], [0, 0, 0]
                for c in range(hd.n_comp):
                    qi[c] = q_index.setdefault(hd.qtables[c], len(q_index))
                    for dst, raw in ((di, hd.htables[c][0]), (ai, hd.htables[c][1])):
                        if raw not in h_index:
                            h_index[raw] = len(h_tabs)
                            h_tabs.append(_derive_huffman(raw[:16], raw[16:]))
                        dst[c] = h_index[raw]
                if len(q_index) > 255 or len(h_tabs) > MAX_HUFF_TABLES_PER_CALL:
                    raise JpegFormatError("too many distinct tables in one batch (decode_into splits batches for you)")
                idx = table_rows[key] = (hd.qtables, hd.htables, tuple(qi), tuple(di), tuple(ai))   # keeps the ids alive
            rows.append((pos + hd.scan_offset, off, len(f) - hd.scan_offset, hd.restart_interval, hd.width, hd.height,
                         hd.n_comp, hd.sampling, idx[2], idx[3], idx[4]))
            self.sizes.append((hd.height, hd.width, hd.n_comp))
            pos += len(f)
        self.total_bytes = pos
        self.images = np.zeros(self.n, dtype=IMAGE_DTYPE)
        if rows:
            cols = list(zip(*rows))
            for name, col in zip(("scan_offset", "out_offset", "scan_bytes", "restart_interval", "width", "height", "n_comp",
                                  "sampling", "qt", "dc", "ac"), cols):
                self.images[name] = np.asarray(col)
        self.qtables = np.frombuffer(b"".join(q_index.keys()), dtype=np.uint16).reshape(-1, 64).copy()
        self.htables = np.stack(h_tabs) if h_tabs else np.zeros(0, dtype=HUFF_DTYPE)


class JpegFileSet:
    """A set of files staged for decoding: bytes concatenated in pinned host memory, headers parsed, tables pooled --
    split into groups of at most 8 distinct Huffman tables (files with optimised tables carry their own).  Reusable:
    a loader stages a video's files once and decodes them into any store slot."""

    def __init__(self, files: Sequence[bytes]):
        self.n = len(files)
        self.groups: List[Tuple[int, int, JpegBatch]] = []          # (first file, end file, batch)
        distinct, start = set(), 0
        for k, f in enumerate(files):
            hd = parse_header(f)
            mine = {t for pair in hd.htables for t in pair}
            if len(distinct | mine) > MAX_HUFF_TABLES_PER_CALL and k > start:
                self.groups.append((start, k, JpegBatch(files[start:k], [0] * (k - start))))
                distinct, start = set(), k
            distinct |= mine
        if self.n > start:
            self.groups.append((start, self.n, JpegBatch(files[start:], [0] * (self.n - start))))
        self.sizes = [s for _, _, b in self.groups for s in b.sizes]
        total = sum(len(f) for f in files)
        self.host = torch.empty(total + 16, dtype=torch.uint8).pin_memory()
        hv = self.host.numpy()
        pos = 0
        self.group_bytes = []
        for a, b_, batch in self.groups:
            g0 = pos
            for f in files[a:b_]:
                hv[pos:pos + len(f)] = np.frombuffer(f, dtype=np.uint8)
                pos += len(f)
            self.group_bytes.append((g0, pos))
        hv[pos:] = 0

    def upload(self, device, stream=None):
        """H2D copy of the compressed files only (pinned host memory -> device), on `stream` (default: current).  Returns the
        per-group device buffers + the event to wait for; pass them to decode_into(staged=...).  Lets a loader run the
        copies of the NEXT step on a copy stream while the decode kernels stay on the compute stream."""
        st = stream if stream is not None else torch.cuda.current_stream()
        bufs = []
        with torch.cuda.stream(st):
            for (g0, g1) in self.group_bytes:
                bufs.append(self.host[g0:g1 + 16].to(device, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(st)
        return bufs, ev

    def decode_into(self, out: torch.Tensor, out_offsets: Sequence[int], staged=None) -> None:
        """Image k goes to byte offset out_offsets[k] of the flat uint8 device tensor `out` as [H][W] (one component) or
        [H][W][3] RGB.  Asynchronous on the current CUDA stream.  staged = upload()'s result: the files are already on the
        device (the current stream waits for that copy)."""
        if not out.is_cuda or out.dtype != torch.uint8 or not out.is_contiguous():
            raise VAError("jpeg decode: `out` must be a contiguous uint8 CUDA tensor (no CPU decode fallback)")
        if len(out_offsets) != self.n:
            raise VAError("jpeg decode: %d offsets for %d files" % (len(out_offsets), self.n))
        for (h, w, c), off in zip(self.sizes, out_offsets):
            if off < 0 or off + h * w * c > out.numel():
                raise VAError("jpeg decode: image of %dx%dx%d at offset %d does not fit in `out`" % (h, w, c, off))
        lib = _lib.load()
        cur = torch.cuda.current_stream()
        if staged is not None:
            cur.wait_event(staged[1])
        for gi, ((a, b_, batch), (g0, g1)) in enumerate(zip(self.groups, self.group_bytes)):
            dev = staged[0][gi] if staged is not None else self.host[g0:g1 + 16].to(out.device, non_blocking=True)
            batch.images["out_offset"] = np.asarray(out_offsets[a:b_], dtype=np.uint64)
            images, q, h = batch.images, batch.qtables, batch.htables
            check(lib.va_jpeg_decode(ptr(dev), images.ctypes.data_as(C.c_void_p), batch.n, q.ctypes.data_as(C.c_void_p),
                                     q.shape[0], h.ctypes.data_as(C.c_void_p), h.shape[0], ptr(out), stream_ptr()),
                  "va_jpeg_decode")
            dev.record_stream(cur)


def decode_into(files: Sequence[bytes], out: torch.Tensor, out_offsets: Sequence[int]) -> None:
    """Decode `files` into the flat uint8 device tensor `out`, image k at byte offset out_offsets[k]."""
    if not out.is_cuda or out.dtype != torch.uint8 or not out.is_contiguous():
        raise VAError("jpeg.decode_into: `out` must be a contiguous uint8 CUDA tensor (no CPU decode fallback)")
    if len(files) == 0:
        return
    JpegFileSet(files).decode_into(out, out_offsets)


def decode(files: Sequence[bytes], device: Optional[torch.device] = None) -> List[torch.Tensor]:
    """Decode a list of files; returns one uint8 CUDA tensor per file ([H,W] or [H,W,3] RGB), views of one buffer."""
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    sizes = []
    for f in files:
        hd = parse_header(f)
        sizes.append((hd.height, hd.width, hd.n_comp))
    offs, pos = [], 0
    for (h, w, c) in sizes:
        offs.append(pos)
        pos += (h * w * c + 15) & ~15
    out = torch.empty(max(pos, 16), dtype=torch.uint8, device=device)
    decode_into(files, out, offs)
    return [out[o:o + h * w * c].view((h, w) if c == 1 else (h, w, c)) for (h, w, c), o in zip(sizes, offs)]

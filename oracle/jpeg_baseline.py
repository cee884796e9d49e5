"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of baseline-JPEG decoding as the reference performs it.

The reference reads every frame / flow image with `Image.open(path)` (Sheet03/spatialModel.py:76-79,
temporalModel.py:85-88) on files written by `cv2.imwrite(frameLoc, frame)` (utils.py:116-120): baseline sequential
JPEG (SOF0), 8-bit, YCbCr 4:2:0 for colour frames and one component for the flow images, standard Huffman tables, no
restart markers.  The arithmetic lives in a third-party dependency that is not under /root/reference: Pillow's bundled
libjpeg-turbo (here Pillow 12.2 / libjpeg-turbo API level 6.2; `PIL.features.version("jpg")`).  What is restated below is
that library's default decompression path, by its published algorithm:
  * jdhuff.c    decode_mcu: canonical Huffman decoding (maxcode / valptr), RECEIVE + EXTEND, DC prediction, ZRL / EOB
  * jdcoefct.c  MCU / block order of an interleaved scan, edge blocks
  * jidctint.c  jpeg_idct_islow: 8x8 integer IDCT, CONST_BITS = 13, PASS1_BITS = 2, range limiting
  * jdsample.c  h2v2_fancy_upsample (triangle filter, bias 8 / 7) with the context rows of jdmainct.c
  * jdcolor.c   ycc_rgb_convert: 16-bit fixed-point tables
Pinning: tests/test_oracle_golden.py::test_jpeg_restatement_vs_pillow decodes files written by cv2.imwrite / Pillow with
THIS module and with Pillow itself and requires identical bytes (colour 4:2:0 and 4:4:4, grayscale, sizes that are not
multiples of the MCU, several qualities, restart intervals).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60,
                   61, 54, 47, 55, 62, 63], dtype=np.int32)       # jpeg_natural_order


class JpegUnsupported(ValueError):
    pass


@dataclass
class Component:
    cid: int
    h: int
    v: int
    tq: int
    td: int = 0
    ta: int = 0


@dataclass
class JpegHeader:
    """Everything in front of the entropy-coded segment (what libjpeg's jdmarker.c collects)."""
    height: int = 0
    width: int = 0
    comps: List[Component] = field(default_factory=list)
    qt: Dict[int, np.ndarray] = field(default_factory=dict)          # table id -> 64 entries in NATURAL order
    huff: Dict[Tuple[int, int], Tuple[np.ndarray, np.ndarray]] = field(default_factory=dict)   # (class, id) -> (bits[17], huffval)
    restart_interval: int = 0
    scan_offset: int = 0                                               # first byte of the entropy-coded data
    adobe_transform: int = -1


def parse_header(data: bytes) -> JpegHeader:
    if data[0:2] != b"\xff\xd8":
        raise JpegUnsupported("not a JPEG (no SOI)")
    hd = JpegHeader()
    i = 2
    n = len(data)
    while i < n:
        if data[i] != 0xFF:
            raise JpegUnsupported("marker expected at byte %d" % i)
        while data[i + 1] == 0xFF:            # fill bytes
            i += 1
        m = data[i + 1]
        i += 2
        if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
            continue
        L = (data[i] << 8) | data[i + 1]
        seg = data[i + 2:i + L]
        if m == 0xDB:                          # DQT
            j = 0
            while j < len(seg):
                pq, tq = seg[j] >> 4, seg[j] & 15
                j += 1
                if pq:
                    vals = np.frombuffer(seg[j:j + 128], dtype=">u2").astype(np.int32)
                    j += 128
                else:
                    vals = np.frombuffer(seg[j:j + 64], dtype=np.uint8).astype(np.int32)
                    j += 64
                nat = np.zeros(64, dtype=np.int32)
                nat[ZIGZAG] = vals                                     # stored in zigzag order
                hd.qt[tq] = nat
        elif m == 0xC4:                        # DHT
            j = 0
            while j < len(seg):
                tc, th = seg[j] >> 4, seg[j] & 15
                bits = np.zeros(17, dtype=np.int32)
                bits[1:] = np.frombuffer(seg[j + 1:j + 17], dtype=np.uint8)
                cnt = int(bits.sum())
                hd.huff[(tc, th)] = (bits, np.frombuffer(seg[j + 17:j + 17 + cnt], dtype=np.uint8).astype(np.int32))
                j += 17 + cnt
        elif m == 0xC0 or m == 0xC1:           # SOF0 / SOF1 (8-bit sequential Huffman)
            if seg[0] != 8:
                raise JpegUnsupported("sample precision %d" % seg[0])
            hd.height, hd.width = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            for k in range(seg[5]):
                c = seg[6 + 3 * k:9 + 3 * k]
                hd.comps.append(Component(c[0], c[1] >> 4, c[1] & 15, c[2]))
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise JpegUnsupported("SOF%d (progressive / lossless / arithmetic) is not the reference's format" % (m - 0xC0))
        elif m == 0xDD:
            hd.restart_interval = (seg[0] << 8) | seg[1]
        elif m == 0xEE and seg[0:5] == b"Adobe":
            hd.adobe_transform = seg[11]
        elif m == 0xDA:                        # SOS
            ns = seg[0]
            if ns != len(hd.comps):
                raise JpegUnsupported("non-interleaved multi-scan files are not the reference's format")
            for k in range(ns):
                cs, t = seg[1 + 2 * k], seg[2 + 2 * k]
                comp = next(c for c in hd.comps if c.cid == cs)
                comp.td, comp.ta = t >> 4, t & 15
            hd.scan_offset = i + L
            return hd
        i += L
    raise JpegUnsupported("no SOS marker")


def derive_table(bits: np.ndarray, huffval: np.ndarray):
    """jdhuff.c jpeg_make_d_derived_tbl: canonical codes -> (maxcode[18], valoffset[17]) for the bit-serial decoder."""
    huffsize = []
    for l in range(1, 17):
        huffsize += [l] * int(bits[l])
    huffcode = []
    code, si = 0, huffsize[0] if huffsize else 0
    p = 0
    while p < len(huffsize):
        while p < len(huffsize) and huffsize[p] == si:
            huffcode.append(code)
            code += 1
            p += 1
        code <<= 1
        si += 1
    maxcode = np.full(18, -1, dtype=np.int64)
    valoffset = np.zeros(17, dtype=np.int64)
    p = 0
    for l in range(1, 17):
        if bits[l]:
            valoffset[l] = p - huffcode[p]
            p += int(bits[l])
            maxcode[l] = huffcode[p - 1]
    maxcode[17] = 0xFFFFF
    return maxcode, valoffset


class BitReader:
    """Entropy-coded segment reader: 0xFF00 -> 0xFF, stops at markers (jdhuff.c jpeg_fill_bit_buffer)."""

    def __init__(self, data: bytes, pos: int):
        self.d, self.pos, self.buf, self.nbits = data, pos, 0, 0

    def _fill(self):
        while self.nbits <= 24:
            if self.pos < len(self.d):
                c = self.d[self.pos]
                if c == 0xFF:
                    c2 = self.d[self.pos + 1] if self.pos + 1 < len(self.d) else 0xD9
                    if c2 == 0:
                        self.pos += 2
                    else:
                        c = 0                   # marker: feed zeros (libjpeg does the same once it hits a marker)
                else:
                    self.pos += 1
            else:
                c = 0
            self.buf = ((self.buf << 8) | c) & 0xFFFFFFFFFFFF
            self.nbits += 8

    def get(self, n: int) -> int:
        if n == 0:
            return 0
        if self.nbits < n:
            self._fill()
        self.nbits -= n
        return (self.buf >> self.nbits) & ((1 << n) - 1)

    def decode(self, maxcode, valoffset, huffval) -> int:
        code, l = self.get(1), 1
        while code > maxcode[l]:
            code = (code << 1) | self.get(1)
            l += 1
        if l > 16:
            return 0
        return int(huffval[int(code + valoffset[l])])

    def restart(self):
        """Discard partial byte, skip the RSTn marker."""
        self.nbits, self.buf = 0, 0
        while self.pos + 1 < len(self.d) and not (self.d[self.pos] == 0xFF and 0xD0 <= self.d[self.pos + 1] <= 0xD7):
            self.pos += 1
        self.pos += 2


def _extend(v: int, s: int) -> int:
    return v if v >= (1 << (s - 1)) else v - (1 << s) + 1


def decode_coefficients(data: bytes, hd: JpegHeader):
    """-> per component int32 [blocks_h, blocks_w, 64] DEQUANTISED coefficients in natural order (padded to whole MCUs)."""
    hmax, vmax = max(c.h for c in hd.comps), max(c.v for c in hd.comps)
    mcux, mcuy = -(-hd.width // (8 * hmax)), -(-hd.height // (8 * vmax))
    if len(hd.comps) == 1:                      # a single-component scan is never interleaved: MCU = one block
        c = hd.comps[0]
        mcux, mcuy = -(-hd.width // 8), -(-hd.height // 8)
        shape = [(mcuy, mcux)]
        per = [(1, 1)]
    else:
        shape = [(mcuy * c.v, mcux * c.h) for c in hd.comps]
        per = [(c.v, c.h) for c in hd.comps]
    coefs = [np.zeros((s[0], s[1], 64), dtype=np.int32) for s in shape]
    tabs = []
    for c in hd.comps:
        dc = hd.huff[(0, c.td)]
        ac = hd.huff[(1, c.ta)]
        tabs.append((derive_table(*dc) + (dc[1],), derive_table(*ac) + (ac[1],), hd.qt[c.tq]))
    br = BitReader(data, hd.scan_offset)
    pred = [0] * len(hd.comps)
    todo = hd.restart_interval
    for my in range(mcuy):
        for mx in range(mcux):
            if hd.restart_interval and todo == 0:
                br.restart()
                pred = [0] * len(hd.comps)
                todo = hd.restart_interval
            for ci in range(len(hd.comps)):
                (dmax, dval, dsym), (amax, aval, asym), q = tabs[ci]
                for by in range(per[ci][0]):
                    for bx in range(per[ci][1]):
                        blk = coefs[ci][my * per[ci][0] + by, mx * per[ci][1] + bx]
                        s = br.decode(dmax, dval, dsym)
                        diff = _extend(br.get(s), s) if s else 0
                        pred[ci] += diff
                        blk[0] = pred[ci] * q[0]
                        k = 1
                        while k < 64:
                            rs = br.decode(amax, aval, asym)
                            r, s = rs >> 4, rs & 15
                            if s:
                                k += r
                                blk[ZIGZAG[k]] = _extend(br.get(s), s) * q[ZIGZAG[k]]
                                k += 1
                            elif r == 15:
                                k += 16
                            else:
                                break
            todo -= 1
    return coefs


# jidctint.c constants (CONST_BITS = 13)
F_0_298631336, F_0_390180644, F_0_541196100, F_0_765366865 = 2446, 3196, 4433, 6270
F_0_899976223, F_1_175875602, F_1_501321110, F_1_847759065 = 7373, 9633, 12299, 15137
F_1_961570560, F_2_053119869, F_2_562915447, F_3_072711026 = 16069, 16819, 20995, 25172


def _idct_1d(d, shift):
    """One pass of jpeg_idct_islow over the LAST axis of int64 array d[..., 8]; result descaled by `shift`."""
    z2, z3 = d[..., 2], d[..., 6]
    z1 = (z2 + z3) * F_0_541196100
    tmp2 = z1 + z3 * (-F_1_847759065)
    tmp3 = z1 + z2 * F_0_765366865
    tmp0 = (d[..., 0] + d[..., 4]) << 13
    tmp1 = (d[..., 0] - d[..., 4]) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = d[..., 7], d[..., 5], d[..., 3], d[..., 1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * F_1_175875602
    t0, t1, t2, t3 = t0 * F_0_298631336, t1 * F_2_053119869, t2 * F_3_072711026, t3 * F_1_501321110
    z1, z2, z3, z4 = z1 * (-F_0_899976223), z2 * (-F_2_562915447), z3 * (-F_1_961570560) + z5, z4 * (-F_0_390180644) + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    out = np.stack([tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2, tmp10 - t3], axis=-1)
    return (out + (1 << (shift - 1))) >> shift


def idct_islow(coefs: np.ndarray) -> np.ndarray:
    """[..., 64] dequantised coefficients (natural order, row-major 8x8) -> [..., 8, 8] uint8 samples."""
    blk = coefs.astype(np.int64).reshape(coefs.shape[:-1] + (8, 8))
    ws = _idct_1d(np.swapaxes(blk, -1, -2), 13 - 2)                 # pass 1: columns -> workspace (transposed view)
    ws = np.swapaxes(ws, -1, -2)
    out = _idct_1d(ws, 13 + 2 + 3)                                   # pass 2: rows
    out = ((out + 512) & 1023) - 512                                 # range_limit[(x) & RANGE_MASK] ...
    return np.clip(out + 128, 0, 255).astype(np.uint8)               # ... is a clamp of x + CENTERJSAMPLE on that window


def _planes(coefs, hd):
    planes = []
    for ci, cf in enumerate(coefs):
        px = idct_islow(cf)                                          # [bh, bw, 8, 8]
        bh, bw = px.shape[:2]
        planes.append(px.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8))
    return planes


def h2v2_fancy_upsample(plane: np.ndarray, ds_h: int, ds_w: int) -> np.ndarray:
    """jdsample.c h2v2_fancy_upsample on the real (downsampled_height x downsampled_width) samples; context rows above
    the first / below the last row are copies of that row (jdmainct.c).  -> [2*ds_h, 2*ds_w] uint8."""
    p = plane[:ds_h, :ds_w].astype(np.int32)
    if ds_w <= 2:                                                     # jinit_upsampler: fancy only if downsampled_width > 2
        return np.repeat(np.repeat(p, 2, axis=0), 2, axis=1).astype(np.uint8)
    up = np.concatenate([p[:1], p[:-1]], axis=0)                      # row above (replicated at the top)
    dn = np.concatenate([p[1:], p[-1:]], axis=0)                      # row below (replicated at the bottom)
    out = np.zeros((2 * ds_h, 2 * ds_w), dtype=np.int32)
    for v, far in ((0, up), (1, dn)):
        colsum = 3 * p + far                                          # thiscolsum per column
        last = np.concatenate([colsum[:, :1], colsum[:, :-1]], axis=1)
        nxt = np.concatenate([colsum[:, 1:], colsum[:, -1:]], axis=1)
        even = (colsum * 3 + last + 8) >> 4
        odd = (colsum * 3 + nxt + 7) >> 4
        even[:, 0] = (colsum[:, 0] * 4 + 8) >> 4                      # special cases for the first / last column
        odd[:, -1] = (colsum[:, -1] * 4 + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out.astype(np.uint8)


def _fix(x):
    return int(x * 65536 + 0.5)


_X = np.arange(256, dtype=np.int64) - 128
CR_R = (_fix(1.40200) * _X + 32768) >> 16
CB_B = (_fix(1.77200) * _X + 32768) >> 16
CR_G = -_fix(0.71414) * _X
CB_G = -_fix(0.34414) * _X + 32768


def ycc_to_rgb(y, cb, cr):
    y = y.astype(np.int64)
    r = y + CR_R[cr]
    g = y + ((CB_G[cb] + CR_G[cr]) >> 16)
    b = y + CB_B[cb]
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def decode(data: bytes) -> np.ndarray:
    """JPEG file bytes -> uint8 [H, W] (one component) or [H, W, 3] RGB, equal to np.asarray(PIL.Image.open(...))."""
    hd = parse_header(data)
    coefs = decode_coefficients(data, hd)
    planes = _planes(coefs, hd)
    H, W = hd.height, hd.width
    if len(hd.comps) == 1:
        return planes[0][:H, :W].copy()
    if len(hd.comps) != 3:
        raise JpegUnsupported("%d components" % len(hd.comps))
    samp = [(c.h, c.v) for c in hd.comps]
    if samp == [(2, 2), (1, 1), (1, 1)]:
        dh, dw = -(-H // 2), -(-W // 2)
        cb = h2v2_fancy_upsample(planes[1], dh, dw)[:H, :W]
        cr = h2v2_fancy_upsample(planes[2], dh, dw)[:H, :W]
    elif samp == [(1, 1), (1, 1), (1, 1)]:
        cb, cr = planes[1][:H, :W], planes[2][:H, :W]
    else:
        raise JpegUnsupported("sampling factors %r (the reference's files are 4:2:0 or one component)" % (samp,))
    return ycc_to_rgb(planes[0][:H, :W], cb, cr)

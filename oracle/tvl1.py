"""CPU oracle (TEST INFRASTRUCTURE ONLY) for SURVEY.md 8f row 4: production of the `flow_x_/flow_y_` images the temporal
stream consumes (Sheet03/parameters.py:27,38-39: directory `mini-ucf101_flow_img_tvl1_gpu`, prefixes `flow_x_`/`flow_y_`;
read by temporalModel.py:76-86).

**parity unpinned.**  The reference never computes optical flow: it reads JPEGs produced by a third-party tool that is
not under /root/reference -- by the directory name, TSN's `dense_flow` (`extract_gpu`: frames -> grey ->
`cv::cuda::OpticalFlowDual_TVL1` -> 8-bit images with bound 20), no version pinned anywhere in the repo.  OpenCV 4.x's
`cudaoptflow` module is NOT in this image (cv2 4.13 here has neither `cv2.optflow` nor `cv2.cuda`), so nothing here can be
checked against the real tool.  This file restates the PUBLISHED algorithm -- Zach, Pock, Bischof, "A Duality Based Approach
for Realtime TV-L1 Optical Flow" (DAGM 2007), in the discretisation of Sanchez, Meinhardt-Llopis, Facciolo, "TV-L1 Optical
Flow Estimation" (IPOL 2013), with the structure and defaults of OpenCV's CUDA implementation (tau 0.25, lambda 0.15,
theta 0.3, 5 scales, 5 warps, epsilon 0.01, 300 iterations, scale step 0.8, gamma 0; bicubic backward warp; bilinear
pyramid; error sum every other iteration with the "less frequent sums" rule) -- and defines every floating-point
expression as SEPARATELY ROUNDED IEEE fp32 operations in the order written below (no fused multiply-add, hypot restated as
sqrt(x*x + y*y)).  The CUDA kernel (csrc/va_tvl1.cu) uses the same operations through __fmul_rn/__fadd_rn/..., so GPU
and oracle agree BIT FOR BIT on the u8 images and on the fp32 flow; what is unpinned is only how close this restatement
is to the third-party tool's own rounding (its nvcc build contracts multiply-adds).

What IS pinned: the frame resize and the grey conversion (`cv2.resize` INTER_LINEAR, `cv2.cvtColor(BGR2GRAY)`, run in this
container; tests/test_tvl1_oracle.py, tests/golden/tvl1_*.npz) and the
8-bit conversion rule (dense_flow's CAST macro: round-half-even of 255*(v+bound)/(2*bound), saturated).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

F = np.float32


@dataclass
class TVL1Params:
    tau: float = 0.25
    lambda_: float = 0.15
    theta: float = 0.3
    nscales: int = 5
    warps: int = 5
    epsilon: float = 0.01
    iterations: int = 300
    scale_step: float = 0.8
    bound: float = 20.0          # dense_flow / TSN `--bound`: flow in [-bound, bound] maps to [0, 255]


def gray_from_rgb(rgb: np.ndarray) -> np.ndarray:
    """cv::cvtColor(BGR2GRAY) on u8 (the tool reads frames with cv::VideoCapture, i.e. BGR): fixed point, 15 fractional
    bits, coefficients R 9798, G 19235, B 3735, round to nearest -- equal to cv2 4.13's output on all 2^24 colours
    (tests/test_tvl1_oracle.py).  Input here is RGB order [H][W][3]."""
    r = rgb[..., 0].astype(np.int32)
    g = rgb[..., 1].astype(np.int32)
    b = rgb[..., 2].astype(np.int32)
    return ((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15).astype(np.uint8)


def resize_linear_u8(src: np.ndarray, dh: int, dw: int) -> np.ndarray:
    """cv::resize(src, Size(dw, dh), INTER_LINEAR) on u8 images, the resize TSN's dense_flow applies to every frame before the
    flow (`new_size` 340 x 256).  OpenCV's fixed-point path: source position (d + 0.5) * scale - 0.5 evaluated in double and
    cast to float, 11-bit coefficients round-half-even((1 - f) * 2048) / (f * 2048); horizontally the fraction is zeroed where the
    position leaves the image, vertically only the row indices are clamped (so the first / last rows blend a row with itself
    using both coefficients); horizontal pass in int32, vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
    PINNED: equal to cv2.resize of this container's OpenCV 4.13 on up- and down-scaling cases (tests/test_tvl1_oracle.py)."""
    sh, sw = src.shape[:2]

    def coeffs(dn, sn, clamp_fraction):
        scale = sn / dn
        d = np.arange(dn, dtype=np.float64)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        fr = (f - s.astype(np.float32)).astype(np.float32)
        if clamp_fraction:
            lo = s < 0
            fr[lo] = 0
            s[lo] = 0
            hi = s >= sn - 1
            fr[hi] = 0
            s[hi] = sn - 1
        c0 = np.rint((np.float32(1.0) - fr) * np.float32(2048)).astype(np.int64)
        c1 = np.rint(fr * np.float32(2048)).astype(np.int64)
        return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), c0, c1

    xs, xs1, a0, a1 = coeffs(dw, sw, True)
    ys, ys1, b0, b1 = coeffs(dh, sh, False)
    S = src.astype(np.int64)
    if S.ndim == 2:
        S = S[..., None]
    hp = S[:, xs, :] * a0[None, :, None] + S[:, xs1, :] * a1[None, :, None]
    r0, r1 = hp[ys], hp[ys1]
    out = ((((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2).clip(0, 255).astype(np.uint8)
    return out if src.ndim == 3 else out[..., 0]


def pyramid_sizes(h: int, w: int, p: TVL1Params):
    """[(h, w)] per scale: each level is round-half-even(size * scale_step); stop before a level below 16 pixels."""
    sizes = [(h, w)]
    for _ in range(1, p.nscales):
        ph, pw = sizes[-1]
        nh, nw = int(round(ph * p.scale_step)), int(round(pw * p.scale_step))
        if nw < 16 or nh < 16:
            break
        sizes.append((nh, nw))
    return sizes


def resize_linear(src: np.ndarray, dh: int, dw: int, fx: np.float32, fy: np.float32) -> np.ndarray:
    """Bilinear resize without half-pixel centres: src_x = dst_x * fx, floor, weights (x2 - src_x) / (src_x - x1), the
    +1 neighbour clamped; the four products are added in the order (y1,x1), (y1,x2), (y2,x1), (y2,x2)."""
    sh, sw = src.shape
    dx = np.arange(dw, dtype=F) * F(fx)
    dy = np.arange(dh, dtype=F) * F(fy)
    x1 = np.floor(dx).astype(np.int64)
    y1 = np.floor(dy).astype(np.int64)
    x2r = np.minimum(x1 + 1, sw - 1)
    y2r = np.minimum(y1 + 1, sh - 1)
    x1f, y1f = x1.astype(F), y1.astype(F)
    ax2 = ((x1f + F(1)) - dx)[None, :]          # x2 - src_x
    ax1 = (dx - x1f)[None, :]                   # src_x - x1
    ay2 = ((y1f + F(1)) - dy)[:, None]
    ay1 = (dy - y1f)[:, None]
    out = src[y1][:, x1] * (ax2 * ay2)
    out = out + src[y1][:, x2r] * (ax1 * ay2)
    out = out + src[y2r][:, x1] * (ax2 * ay1)
    out = out + src[y2r][:, x2r] * (ax1 * ay1)
    return out.astype(F)


def centered_gradient(img: np.ndarray):
    h, w = img.shape
    xs = np.arange(w)
    ys = np.arange(h)
    dx = F(0.5) * (img[:, np.minimum(xs + 1, w - 1)] - img[:, np.maximum(xs - 1, 0)])
    dy = F(0.5) * (img[np.minimum(ys + 1, h - 1), :] - img[np.maximum(ys - 1, 0), :])
    return dx.astype(F), dy.astype(F)


def _bicubic(x_: np.ndarray) -> np.ndarray:
    x = np.abs(x_)
    a = (x * x) * (F(1.5) * x - F(2.5)) + F(1.0)
    b = x * (x * (F(-0.5) * x + F(2.5)) - F(4.0)) + F(2.0)
    return np.where(x <= F(1.0), a, np.where(x < F(2.0), b, F(0.0))).astype(F)


def warp_backward(i0, i1, i1x, i1y, u1, u2):
    """Bicubic samples of I1 and its gradient at (x + u1, y + u2) (clamped addressing), normalised by the weight sum;
    returns (I1wx, I1wy, grad, rho_c)."""
    h, w = i0.shape
    xs = np.arange(w, dtype=F)[None, :]
    ys = np.arange(h, dtype=F)[:, None]
    wx = xs + u1
    wy = ys + u2
    xmin, xmax = np.ceil(wx - F(2.0)), np.floor(wx + F(2.0))
    ymin, ymax = np.ceil(wy - F(2.0)), np.floor(wy + F(2.0))
    s = np.zeros((h, w), F)
    sx = np.zeros((h, w), F)
    sy = np.zeros((h, w), F)
    ws = np.zeros((h, w), F)
    for j in range(5):
        cy = ymin + F(j)
        oky = cy <= ymax
        wyv = _bicubic(wy - cy)
        iy = np.clip(cy, 0, h - 1).astype(np.int64)
        for i in range(5):
            cx = xmin + F(i)
            ok = oky & (cx <= xmax)
            wgt = _bicubic(wx - cx) * wyv
            ix = np.clip(cx, 0, w - 1).astype(np.int64)
            s = np.where(ok, s + wgt * i1[iy, ix], s)
            sx = np.where(ok, sx + wgt * i1x[iy, ix], sx)
            sy = np.where(ok, sy + wgt * i1y[iy, ix], sy)
            ws = np.where(ok, ws + wgt, ws)
    coeff = F(1.0) / ws
    i1w = s * coeff
    i1wx = sx * coeff
    i1wy = sy * coeff
    grad = i1wx * i1wx + i1wy * i1wy
    rho_c = ((i1w - i1wx * u1) - i1wy * u2) - i0
    return i1wx.astype(F), i1wy.astype(F), grad.astype(F), rho_c.astype(F)


def _divergence(v1, v2):
    d = v1 + v2                                  # (0,0): v1 + v2
    d[1:, 0] = (v1[1:, 0] + v2[1:, 0]) - v2[:-1, 0]
    d[0, 1:] = (v1[0, 1:] - v1[0, :-1]) + v2[0, 1:]
    d[1:, 1:] = (v1[1:, 1:] - v1[1:, :-1]) + (v2[1:, 1:] - v2[:-1, 1:])
    return d


def estimate_u(i1wx, i1wy, grad, rho_c, p11, p12, p21, p22, u1, u2, l_t: np.float32, theta: np.float32, calc_error: bool):
    rho = rho_c + (i1wx * u1 + i1wy * u2)
    thr = l_t * grad
    lo = rho < -thr
    hi = (~lo) & (rho > thr)
    mid = (~lo) & (~hi) & (grad > np.finfo(F).eps)
    with np.errstate(divide="ignore", invalid="ignore"):
        fi = np.where(mid, (-rho) / grad, F(0)).astype(F)
    d1 = np.where(lo, l_t * i1wx, np.where(hi, -(l_t * i1wx), np.where(mid, fi * i1wx, F(0))))
    d2 = np.where(lo, l_t * i1wy, np.where(hi, -(l_t * i1wy), np.where(mid, fi * i1wy, F(0))))
    v1 = u1 + d1
    v2 = u2 + d2
    u1n = (v1 + theta * _divergence(p11, p12)).astype(F)
    u2n = (v2 + theta * _divergence(p21, p22)).astype(F)
    err = None
    if calc_error:
        e1 = u1n - u1
        e2 = u2n - u2
        err = float(np.sum((e1 * e1 + e2 * e2).astype(np.float64)))
    return u1n, u2n, err


def estimate_dual(u1, u2, p11, p12, p21, p22, taut: np.float32):
    def fwd(u):
        ux = np.empty_like(u)
        uy = np.empty_like(u)
        ux[:, :-1] = u[:, 1:] - u[:, :-1]
        ux[:, -1] = F(0)                          # u(y, min(x+1, w-1)) - u(y, x)
        uy[:-1, :] = u[1:, :] - u[:-1, :]
        uy[-1, :] = F(0)
        return ux, uy
    u1x, u1y = fwd(u1)
    u2x, u2y = fwd(u2)
    g1 = np.sqrt(u1x * u1x + u1y * u1y)
    g2 = np.sqrt(u2x * u2x + u2y * u2y)
    ng1 = F(1.0) + taut * g1
    ng2 = F(1.0) + taut * g2
    return (((p11 + taut * u1x) / ng1).astype(F), ((p12 + taut * u1y) / ng1).astype(F),
            ((p21 + taut * u2x) / ng2).astype(F), ((p22 + taut * u2y) / ng2).astype(F))


def _one_scale(i0, i1, u1, u2, p: TVL1Params, stats):
    h, w = i0.shape
    scaled_eps = p.epsilon * p.epsilon * (h * w)
    i1x, i1y = centered_gradient(i1)
    p11 = np.zeros((h, w), F)
    p12 = np.zeros((h, w), F)
    p21 = np.zeros((h, w), F)
    p22 = np.zeros((h, w), F)
    l_t = F(p.lambda_ * p.theta)
    taut = F(p.tau / p.theta)
    theta = F(p.theta)
    for _ in range(p.warps):
        i1wx, i1wy, grad, rho_c = warp_backward(i0, i1, i1x, i1y, u1, u2)
        error = math.inf
        prev = 0.0
        n = 0
        while error > scaled_eps and n < p.iterations:
            calc = (p.epsilon > 0) and bool(n & 1) and (prev < scaled_eps)
            u1, u2, e = estimate_u(i1wx, i1wy, grad, rho_c, p11, p12, p21, p22, u1, u2, l_t, theta, calc)
            if calc:
                error = e
                prev = e
            else:
                error = math.inf
                prev -= scaled_eps
            p11, p12, p21, p22 = estimate_dual(u1, u2, p11, p12, p21, p22, taut)
            n += 1
        stats.append(n)
    return u1, u2


def tvl1_flow(i0_u8: np.ndarray, i1_u8: np.ndarray, params: TVL1Params | None = None, return_stats: bool = False):
    """fp32 flow (u1 = x displacement, u2 = y displacement) from grey frame i0 to grey frame i1, both u8 [H][W]."""
    p = params or TVL1Params()
    h, w = i0_u8.shape
    sizes = pyramid_sizes(h, w, p)
    i0s = [i0_u8.astype(F)]
    i1s = [i1_u8.astype(F)]
    fpyr = F(1.0 / p.scale_step)
    for s in range(1, len(sizes)):
        i0s.append(resize_linear(i0s[-1], sizes[s][0], sizes[s][1], fpyr, fpyr))
        i1s.append(resize_linear(i1s[-1], sizes[s][0], sizes[s][1], fpyr, fpyr))
    u1 = np.zeros(sizes[-1], F)
    u2 = np.zeros(sizes[-1], F)
    stats = []
    for s in range(len(sizes) - 1, -1, -1):
        u1, u2 = _one_scale(i0s[s], i1s[s], u1, u2, p, stats)
        if s == 0:
            break
        dh, dw = sizes[s - 1]
        sh, sw = sizes[s]
        fx = F(1.0 / (float(dw) / sw))
        fy = F(1.0 / (float(dh) / sh))
        up = F(1.0 / p.scale_step)
        u1 = (resize_linear(u1, dh, dw, fx, fy) * up).astype(F)
        u2 = (resize_linear(u2, dh, dw, fx, fy) * up).astype(F)
    if return_stats:
        return u1, u2, stats
    return u1, u2


def flow_to_u8(u: np.ndarray, bound: float) -> np.ndarray:
    """dense_flow's CAST(v, -bound, bound): v > H -> 255, v < L -> 0, else cvRound(255 * (v - L) / (H - L)) in double."""
    v = u.astype(np.float64)
    q = np.rint(255.0 * (v + bound) / (2.0 * bound))
    q = np.where(v > bound, 255.0, np.where(v < -bound, 0.0, q))
    return q.astype(np.uint8)


def flow_images(frame0_rgb: np.ndarray, frame1_rgb: np.ndarray, params: TVL1Params | None = None, new_size=None):
    """The whole producer for one frame pair: RGB u8 frames -> (flow_x u8, flow_y u8), the images temporalModel.py:80-81
    opens.  new_size = (width, height): dense_flow's resize of every frame before the grey conversion (TSN: 340 x 256)."""
    p = params or TVL1Params()
    if new_size is not None:
        frame0_rgb = resize_linear_u8(frame0_rgb, new_size[1], new_size[0])
        frame1_rgb = resize_linear_u8(frame1_rgb, new_size[1], new_size[0])
    g0 = gray_from_rgb(frame0_rgb) if frame0_rgb.ndim == 3 else frame0_rgb
    g1 = gray_from_rgb(frame1_rgb) if frame1_rgb.ndim == 3 else frame1_rgb
    u1, u2 = tvl1_flow(g0, g1, p)
    return flow_to_u8(u1, p.bound), flow_to_u8(u2, p.bound)

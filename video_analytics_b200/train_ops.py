"""Python call layer over the training-step primitives of the C-ABI (include/va_b200.h, "Training-step primitives").

Same rules as ops.py: torch only owns device memory and the stream; every function launches hand-written sm_100a
kernels and raises `VAError` on failure.  Activations and activation gradients are bf16 NHWC, parameter gradients
fp32 in the reference's own layouts (Conv2d OIHW, Linear [out][in]).
Reference being replaced: loss.backward() / optimizer.step() in train() (Sheet03/spatialModel.py:178-181).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .ops import _need_cuda


def maxpool2x2(x: torch.Tensor, with_codes: bool = False):
    """bf16 NHWC [n,H,W,C] -> [n,H/2,W/2,C]  (MaxPool2d(2,2)).  with_codes=True also returns the gradient-routing codes
    (int32 [n,H/2,W/2,C/8], 4 bits per window and channel) that pool_bwd_codes consumes instead of the activation."""
    _need_cuda(x)
    assert x.dtype == torch.bfloat16
    n, H, W, Cc = x.shape
    y = torch.empty((n, H // 2, W // 2, Cc), dtype=torch.bfloat16, device=x.device)
    codes = torch.empty((n, H // 2, W // 2, Cc // 8), dtype=torch.int32, device=x.device) if with_codes else None
    check(_lib.load().va_maxpool2x2_nhwc(ptr(x), n, H, W, Cc, ptr(y), ptr(codes), stream_ptr()), "va_maxpool2x2_nhwc")
    return (y, codes) if with_codes else y


def pool_bwd_codes(dout: torch.Tensor, codes: torch.Tensor, *, bias_grad_out=None) -> torch.Tensor:
    """Gradient w.r.t. the pre-activation of relu -> maxpool from the pooled gradient dout [n,H/2,W/2,C] and the codes of
    maxpool2x2(..., with_codes=True); same result as relu_pool_bwd(dout, y, pooled=True) without reading y."""
    _need_cuda(dout, codes, bias_grad_out)
    assert dout.dtype == torch.bfloat16 and codes.dtype == torch.int32
    n, Ho, Wo, Cc = dout.shape
    assert tuple(codes.shape) == (n, Ho, Wo, Cc // 8)
    dz = torch.empty((n, 2 * Ho, 2 * Wo, Cc), dtype=torch.bfloat16, device=dout.device)
    check(_lib.load().va_pool_bwd_codes(ptr(dout), ptr(codes), n, 2 * Ho, 2 * Wo, Cc, ptr(dz), ptr(bias_grad_out),
                                        stream_ptr()), "va_pool_bwd_codes")
    return dz


def relu_pool_bwd(dout: torch.Tensor, y: torch.Tensor, *, pooled: bool, bias_grad_out=None) -> torch.Tensor:
    """Gradient w.r.t. the pre-activation z of y = relu(z) [n,H,W,C], given the gradient of pool(y) (pooled=True,
    dout [n,H/2,W/2,C]) or of y itself (pooled=False).  bias_grad_out: fp32 [C] tensor that receives
    sum over pixels of dz (the layer's bias gradient) from the same pass."""
    _need_cuda(dout, y, bias_grad_out)
    assert dout.dtype == torch.bfloat16 and y.dtype == torch.bfloat16
    n, H, W, Cc = y.shape
    if bias_grad_out is not None:
        assert bias_grad_out.dtype == torch.float32 and bias_grad_out.numel() == Cc
    dz = torch.empty_like(y)
    check(_lib.load().va_relu_pool_bwd(ptr(dout), ptr(y), n, H, W, Cc, int(pooled), ptr(dz), ptr(bias_grad_out),
                                       stream_ptr()), "va_relu_pool_bwd")
    return dz


def _out(out, shape, device):
    if out is None:
        return torch.empty(shape, dtype=torch.float32, device=device)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == int(torch.Size(shape).numel()), \
        (tuple(out.shape), tuple(shape))
    return out


def bias_grad(dz: torch.Tensor, out=None) -> torch.Tensor:
    """db[c] = sum over all leading dims of dz[..., c]  (fp32; `out` is overwritten)."""
    _need_cuda(dz)
    assert dz.dtype == torch.bfloat16
    Cc = dz.shape[-1]
    rows = dz.numel() // Cc
    db = _out(out, (Cc,), dz.device)
    check(_lib.load().va_bias_grad(ptr(dz), rows, Cc, ptr(db), stream_ptr()), "va_bias_grad")
    return db


def dropout(x: torch.Tensor, mask: torch.Tensor, p: float = 0.5) -> torch.Tensor:
    """y = x / (1-p) where mask else 0; the same map is its own backward.  mask: uint8, caller-supplied (the parity
    tests share it with the oracle; the training loop draws it with torch's generator)."""
    _need_cuda(x, mask)
    assert mask.dtype == torch.uint8 and mask.numel() == x.numel()
    assert x.dtype in (torch.bfloat16, torch.float32)
    y = torch.empty_like(x)
    check(_lib.load().va_dropout(ptr(x), ptr(mask), x.numel(), 1.0 / (1.0 - p), int(x.dtype == torch.float32), ptr(y),
                                 stream_ptr()), "va_dropout")
    return y


def conv2d_dgrad(dz: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """dX of a 3x3/s1/p1 convolution: dz bf16 [n,H,W,cout], w fp32 OIHW [cout,cin,3,3] -> bf16 [n,H,W,cin]."""
    _need_cuda(dz, w)
    assert dz.dtype == torch.bfloat16 and w.dtype == torch.float32
    n, H, W, cout = dz.shape
    assert w.shape[0] == cout and tuple(w.shape[2:]) == (3, 3)
    cin = w.shape[1]
    dx = torch.empty((n, H, W, cin), dtype=torch.bfloat16, device=dz.device)
    check(_lib.load().va_conv2d_dgrad(ptr(dz), n, H, W, cout, ptr(w), cin, ptr(dx), stream_ptr()), "va_conv2d_dgrad")
    return dx


def linear_dgrad(dy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """dX = dY . W: dy bf16 [n,out], w fp32 [out,in] -> bf16 [n,in]."""
    _need_cuda(dy, w)
    assert dy.dtype == torch.bfloat16 and w.dtype == torch.float32
    n, fout = dy.shape
    fin = w.shape[1]
    dx = torch.empty((n, fin), dtype=torch.bfloat16, device=dy.device)
    check(_lib.load().va_linear_dgrad(ptr(dy), n, fout, ptr(w), fin, ptr(dx), stream_ptr()), "va_linear_dgrad")
    return dx


def conv2d_wgrad(dz: torch.Tensor, x: torch.Tensor, cin: int, ks: int = 3, out=None) -> torch.Tensor:
    """dW fp32 OIHW [cout,cin,ks,ks] from dz bf16 [n,H,W,cout] and the layer input x bf16 [n,H,W,cin_pad>=cin]."""
    _need_cuda(dz, x)
    assert dz.dtype == torch.bfloat16 and x.dtype == torch.bfloat16
    n, H, W, cout = dz.shape
    assert tuple(x.shape[:3]) == (n, H, W)
    dw = _out(out, (cout, cin, ks, ks), dz.device)
    check(_lib.load().va_wgrad(ptr(dz), ptr(x), n, H, W, cout, cin, x.shape[3], ks, ptr(dw), stream_ptr()), "va_wgrad")
    return dw


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, out=None) -> torch.Tensor:
    """dW = dY^T . X fp32 [out,in] from dy bf16 [n,out], x bf16 [n,in]  (the batch is the GEMM's K dimension)."""
    _need_cuda(dy, x)
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16
    n, fout = dy.shape
    fin = x.shape[1]
    dw = _out(out, (fout, fin), dy.device)
    check(_lib.load().va_wgrad(ptr(dy), ptr(x), 1, 1, n, fout, fin, fin, 1, ptr(dw), stream_ptr()), "va_wgrad")
    return dw


def ce_train(x: torch.Tensor, w4: torch.Tensor, b4: torch.Tensor, labels: torch.Tensor, dw4=None, db4=None):
    """fp32 logit layer + mean cross-entropy, forward and backward in one call.
    x fp32 [n,D] (descriptors after dropout), w4 [C,D], b4 [C], labels int64 [n].
    Returns dict(logits, loss (1-element tensor), dlogits, dw4, db4, dx)."""
    _need_cuda(x, w4, b4, labels)
    assert x.dtype == torch.float32 and labels.dtype == torch.int64
    n, D = x.shape
    Cc = w4.shape[0]
    dev = x.device
    out = dict(logits=torch.empty((n, Cc), dtype=torch.float32, device=dev),
               dlogits=torch.empty((n, Cc), dtype=torch.float32, device=dev),
               loss=torch.zeros((1,), dtype=torch.float32, device=dev),
               dw4=_out(dw4, (Cc, D), dev), db4=_out(db4, (Cc,), dev),
               dx=torch.empty((n, D), dtype=torch.float32, device=dev))
    check(_lib.load().va_ce_train(ptr(x), ptr(w4), ptr(b4), ptr(labels), n, D, Cc, ptr(out["logits"]), ptr(out["dlogits"]),
                                  ptr(out["loss"]), ptr(out["dw4"]), ptr(out["db4"]), ptr(out["dx"]), stream_ptr()),
          "va_ce_train")
    return out


def relu_bwd_f32_to_bf16(dy: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    _need_cuda(dy, y)
    assert dy.dtype == torch.float32 and y.dtype == torch.float32
    dz = torch.empty(dy.shape, dtype=torch.bfloat16, device=dy.device)
    check(_lib.load().va_relu_bwd_f32_to_bf16(ptr(dy), ptr(y), dy.numel(), ptr(dz), stream_ptr()), "va_relu_bwd_f32_to_bf16")
    return dz


def transpose_bf16(x: torch.Tensor) -> torch.Tensor:
    """[n,A,B] -> [n,B,A] (bf16)."""
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 3
    n, A, B = x.shape
    y = torch.empty((n, B, A), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().va_transpose_bf16(ptr(x), n, A, B, ptr(y), stream_ptr()), "va_transpose_bf16")
    return y


def f32_to_bf16(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    assert x.dtype == torch.float32
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_lib.load().va_f32_to_bf16(ptr(x), x.numel(), ptr(y), stream_ptr()), "va_f32_to_bf16")
    return y


def sgd_momentum_(param: torch.Tensor, grad: torch.Tensor, buf: torch.Tensor, *, lr: float, momentum: float,
                  first_step: bool, grad_scale: float = 1.0) -> None:
    """In place torch.optim.SGD(lr, momentum) update (dampening 0, no nesterov, no weight decay;
    reference spatialModel.py:116)."""
    _need_cuda(param, grad, buf)
    assert param.dtype == buf.dtype == torch.float32 and param.numel() == grad.numel() == buf.numel()
    if grad.dtype == torch.bfloat16:         # the compressed payload of a data-parallel gradient all-reduce
        check(_lib.load().va_sgd_momentum_bf16g(ptr(param), ptr(grad), ptr(buf), param.numel(), lr, momentum, int(first_step),
                                                grad_scale, stream_ptr()), "va_sgd_momentum_bf16g")
        return
    assert grad.dtype == torch.float32
    check(_lib.load().va_sgd_momentum(ptr(param), ptr(grad), ptr(buf), param.numel(), lr, momentum, int(first_step),
                                      grad_scale, stream_ptr()), "va_sgd_momentum")


def f32_to_bf16_(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """Cast into a caller-owned bf16 buffer of the same numel."""
    _need_cuda(x, out)
    assert x.dtype == torch.float32 and out.dtype == torch.bfloat16 and x.numel() == out.numel()
    check(_lib.load().va_f32_to_bf16(ptr(x), x.numel(), ptr(out), stream_ptr()), "va_f32_to_bf16")
    return out


def reserve_sms(sms: int, launches: int) -> None:
    """va_reserve_sms: the next `launches` layer-kernel launches of this process leave `sms` SMs free for a collective
    that was just launched on another stream (0, 0 switches it off)."""
    check(_lib.load().va_reserve_sms(int(sms), int(launches)), "va_reserve_sms")


def allreduce_bf16_(peer_ptrs, multicast_ptr: int, world: int, rank: int, n_elems: int, n_ctas: int = 8) -> None:
    """va_allreduce_bf16: in-place sum over the ranks' symmetric bf16 buffers (the caller brackets it with the symmetric
    memory barrier).  peer_ptrs: sequence of `world` device addresses (int); multicast_ptr: NVSwitch multicast address or 0."""
    import ctypes as C
    arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
    check(_lib.load().va_allreduce_bf16(arr, C.c_void_p(int(multicast_ptr) or None), int(world), int(rank), C.c_longlong(int(n_elems)),
                                        int(n_ctas), stream_ptr()), "va_allreduce_bf16")

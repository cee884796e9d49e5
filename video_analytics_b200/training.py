"""K5: one SGD-momentum training step of a VGG16 stream on hand-written sm_100a kernels.

Replaces the body of the reference's train() loop (Sheet03/spatialModel.py:171-181 / temporalModel.py same lines):

    op = features(ip); op = classifier[:-1](op)  [train mode: Dropout(0.5) after FC1..FC3];  featureVectors = op
    op = classifier[-1](op); loss = CrossEntropyLoss(op, labels); zero_grad(); loss.backward(); optimizer.step()

Layout of the step on the device:
  * parameters, gradients and momentum buffers live in three flat fp32 arenas (one slice per state_dict tensor, in
    state_dict order); the torch module's parameters and torch.optim.SGD's momentum buffers are re-pointed at views of
    the arenas, so `model.state_dict()` / `optimizer.state_dict()` -- the reference's checkpoint format -- stay valid;
  * forward: the evaluation path's tcgen05 layer kernel (bf16 operands, fp32 accumulate) with the 2x2 max-pool run as
    its own kernel so that the un-pooled activation is kept for the backward pass; Dropout masks are uint8 tensors
    (drawn by torch's generator, or supplied by the caller so the oracle can share them);
  * backward: ReLU/pool routing kernel -> weight-gradient GEMM (tcgen05, split-K fp32 atomics) + bias-gradient
    reduction -> data-gradient through the layer kernel with rotated/transposed filters; gradients between layers bf16;
  * multi-GPU: NCCL all-reduce (sum) over the flat gradient arena in two slices -- the classifier's 476 MB as soon as
    they exist (first in the backward pass, hidden under the conv-stack backward), the conv stack's 59 MB at the end --
    with the 1/world scale folded into the update kernel;
  * update: ONE fused SGD-momentum launch over the whole arena.

There is no torch autograd and no torch compute op on this path; torch owns memory, streams and the process group.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from . import train_ops as T
from ._lib import VAError
from .ops import STATE_DICT_KEYS

# torchvision vgg16 "D": conv index -> is it followed by MaxPool2d(2,2)
POOL_AFTER = (False, True, False, True, False, False, True, False, False, True, False, False, True)
DROPOUT_P = 0.5          # nn.Dropout() default, reference spatialModel.py:143-150


class StreamTrainer:
    """Owns the flat parameter / gradient / momentum arenas of one stream and runs training steps on them."""

    def __init__(self, module: torch.nn.Module, optimizer: Optional[torch.optim.SGD] = None, *, lr: float = 0.1,
                 momentum: float = 0.9, c_pad: int = 16, process_group=None, grad_allreduce_dtype: str = "bf16",
                 overlap_allreduce: Optional[bool] = None, defer_update: Optional[bool] = None,
                 reserve_sms: Optional[int] = None, reserve_launches: Optional[int] = None,
                 allreduce_impl: Optional[str] = None):
        """module: the (unwrapped) torchvision-layout VGG16 with the swapped classifier (parameter container only).
        optimizer: the torch.optim.SGD over module.parameters(); its lr / momentum are read at every step (so a
        MultiStepLR scheduler keeps working) and its momentum buffers are re-pointed at the arena."""
        if not torch.cuda.is_available():
            raise VAError("training runs on a B200 only (no CPU fallback)")
        self.module = module.cuda()
        self.optimizer = optimizer
        self._lr, self._momentum = lr, momentum
        self.c_pad = c_pad
        self.group = process_group
        # Data-parallel gradient payload: "bf16" (default; 270 MB per stream instead of 541 MB -- the layer kernels hold
        # every SM, so NCCL's kernels cannot overlap them and the collective's bytes are exposed time: at 2 GPUs the fp32
        # all-reduce cost 1.7 of 50 ms per step) or "fp32" (bit-exact average of the ranks' fp32 gradients).
        if grad_allreduce_dtype not in ("bf16", "fp32"):
            raise VAError("grad_allreduce_dtype must be 'bf16' or 'fp32'")
        self.grad_allreduce_dtype = grad_allreduce_dtype
        # overlap_allreduce=True launches the classifier slice's collective in the middle of the backward pass.  The layer
        # kernels are PERSISTENT with one CTA per SM: while NCCL's CTAs hold some SMs, a layer kernel's last CTAs start only
        # after others finish and that layer takes up to twice as long -- measured at N = 2: 50.1-50.4 ms per step against
        # 48.1 ms at N = 1, whatever the payload size.  Default (None -> VA_OVERLAP_ALLREDUCE env, else False): ONE collective
        # over the whole arena after the backward pass -- exposed, but only for the time the bytes take.
        if overlap_allreduce is None:
            import os
            overlap_allreduce = os.environ.get("VA_OVERLAP_ALLREDUCE", "0") == "1"
        self.overlap_allreduce = bool(overlap_allreduce)
        # defer_update=True (data parallel only): step() returns as soon as the gradient all-reduce is LAUNCHED; the wait and
        # the SGD update run at the start of this trainer's next step() (or flush()).  The two streams' trainers alternate,
        # so one stream's collective runs under the other stream's forward pass -- same arithmetic, same order of updates.
        # For NCCL's CTAs to find room beside the persistent layer kernels, the next `reserve_launches` layer launches of
        # the process leave `reserve_sms` SMs free (va_reserve_sms); the collective then costs those layers ~sms/148 of
        # their time instead of its own duration.  Measured in DESIGN.md section 5.
        import os as _os
        if defer_update is None:
            defer_update = _os.environ.get("VA_DEFER_UPDATE", "0") == "1"
        self.defer_update = bool(defer_update) and process_group is not None
        self.reserve_sms = int(_os.environ.get("VA_ALLREDUCE_SMS", "8")) if reserve_sms is None else int(reserve_sms)
        self.reserve_launches = int(_os.environ.get("VA_ALLREDUCE_LAUNCHES", "5")) if reserve_launches is None else int(reserve_launches)
        self._deferred = None
        # allreduce_impl: "nccl" (torch.distributed) or "va" -- the library's own two-shot kernel over NVLink peer memory /
        # NVSwitch multicast (csrc/va_allreduce.cu) on a symmetric allocation (torch symmetric memory supplies the peer and
        # multicast addresses and the cross-rank barrier).  bf16 payload only.  None -> VA_ALLREDUCE_IMPL env, else "nccl".
        if allreduce_impl is None:
            allreduce_impl = _os.environ.get("VA_ALLREDUCE_IMPL", "nccl")
        if allreduce_impl not in ("nccl", "va"):
            raise VAError("allreduce_impl must be 'nccl' or 'va'")
        self.allreduce_impl = allreduce_impl if (process_group is not None and grad_allreduce_dtype == "bf16") else "nccl"
        self._symm = None
        sd = dict(self.module.named_parameters())
        missing = [k for k in STATE_DICT_KEYS if k not in sd]
        if missing:
            raise VAError(f"module is missing parameters {missing[:3]}...")
        self.params: List[torch.nn.Parameter] = [sd[k] for k in STATE_DICT_KEYS]
        sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + ((s + 63) // 64) * 64)      # 256-byte aligned slices
        total = self.offsets[-1]
        dev = torch.device("cuda", torch.cuda.current_device())
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_buf = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad_bf16 = None
        if process_group is not None and grad_allreduce_dtype == "bf16":
            if self.allreduce_impl == "va":
                import sys
                import torch.distributed as dist
                ok_flag = torch.ones(1, dtype=torch.int32, device=dev)
                try:
                    import torch.distributed._symmetric_memory as symm_mem
                    self.flat_grad_bf16 = symm_mem.empty(total, dtype=torch.bfloat16, device=dev)
                    self.flat_grad_bf16.zero_()
                    self._symm = symm_mem.rendezvous(self.flat_grad_bf16, process_group)
                    self._symm_world, self._symm_rank = dist.get_world_size(process_group), dist.get_rank(process_group)
                    self._symm_use_mc = (_os.environ.get("VA_ALLREDUCE_MULTICAST", "1") == "1" and
                                         int(self._symm.multicast_ptr or 0) != 0)
                    if self._symm_world not in (1, 2, 4, 8) and not self._symm_use_mc:
                        raise VAError("peer-pointer form needs 1, 2, 4 or 8 ranks")
                    self._comm_stream = torch.cuda.Stream(device=dev)
                except Exception as e:          # no symmetric memory on this box: every rank falls back together, loudly
                    ok_flag.zero_()
                    sys.stderr.write(f"[video_analytics_b200] allreduce_impl='va' unavailable ({e!r}); using NCCL\n")
                dist.all_reduce(ok_flag, op=dist.ReduceOp.MIN, group=process_group)
                if int(ok_flag.item()) == 0:
                    self._symm, self.allreduce_impl = None, "nccl"
                    self.flat_grad_bf16 = torch.zeros(total, dtype=torch.bfloat16, device=dev)
            else:
                self.flat_grad_bf16 = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        self.grads: List[torch.Tensor] = []
        self.bufs: List[torch.Tensor] = []
        for p, off in zip(self.params, self.offsets):
            view = self.flat_param[off:off + p.numel()].view(p.shape)
            view.copy_(p.data.to(dev, torch.float32))                              # one-time arena fill
            p.data = view
            self.grads.append(self.flat_grad[off:off + p.numel()].view(p.shape))
            self.bufs.append(self.flat_buf[off:off + p.numel()].view(p.shape))
        self.steps_done = 0
        self._adopt_optimizer_state()
        self.sync_replicas()

    def sync_replicas(self):
        """Data-parallel replicas must start from the SAME parameters and momentum: the torch module each rank built is
        initialised from its own RNG (and a resumed checkpoint may exist on one rank only), and nothing downstream would
        notice a divergence -- the all-reduce only averages gradients.  Rank 0's arenas are broadcast, as
        nn.DataParallel's replicate() (reference spatialModel.py:133) would copy them from device 0."""
        if self.group is None:
            return
        import torch.distributed as dist
        if dist.get_world_size(self.group) < 2:
            return
        flag = torch.tensor([self.steps_done], device=self.flat_param.device)
        dist.broadcast(flag, src=0, group=self.group)
        dist.broadcast(self.flat_param, src=0, group=self.group)
        dist.broadcast(self.flat_buf, src=0, group=self.group)
        if int(flag.item()) and self.steps_done == 0:
            self._publish_optimizer_state()          # rank 0 resumed with momentum buffers: adopt them here too
        self.steps_done = int(flag.item())

    # ---- optimizer state <-> arena (checkpoint compatibility, reference :240-246,255-260)
    def _adopt_optimizer_state(self):
        """If the optimizer already carries momentum buffers (resume()), copy them into the arena and re-point."""
        if self.optimizer is None:
            return
        have = 0
        for p, buf in zip(self.params, self.bufs):
            st = self.optimizer.state.get(p, {})
            mb = st.get("momentum_buffer")
            if mb is not None:
                buf.copy_(mb.to(buf.device, torch.float32))
                st["momentum_buffer"] = buf
                have += 1
        if have not in (0, len(self.params)):
            raise VAError("optimizer state holds momentum buffers for only some parameters")
        self.steps_done = 1 if have else 0

    def _publish_optimizer_state(self):
        if self.optimizer is None:
            return
        for p, buf in zip(self.params, self.bufs):
            self.optimizer.state[p]["momentum_buffer"] = buf

    def hyper(self):
        if self.optimizer is not None:
            g = self.optimizer.param_groups[0]
            return float(g["lr"]), float(g["momentum"])
        return self._lr, self._momentum

    def param(self, key: str) -> torch.Tensor:
        self.flush()
        return self.params[STATE_DICT_KEYS.index(key)].data

    def grad(self, key: str) -> torch.Tensor:
        return self.grads[STATE_DICT_KEYS.index(key)]

    # ---- masks
    @staticmethod
    def draw_masks(n: int, desc_dim: int, generator: Optional[torch.Generator] = None, device="cuda"):
        """Keep-masks (1 = keep) of the three Dropout layers, Bernoulli(1 - p)."""
        shapes = [(n, 4096), (n, 4096), (n, desc_dim)]
        return [(torch.rand(s, generator=generator, device=generator.device if generator is not None else device)
                 >= DROPOUT_P).to(torch.uint8).to(device) for s in shapes]

    # ---- the step
    def forward_backward(self, x_nhwc: torch.Tensor, labels: torch.Tensor, masks: Optional[Sequence[torch.Tensor]] = None,
                         keep: Optional[dict] = None, on_classifier_grads=None):
        """Fills the gradient arena; returns (loss [1] fp32 tensor, featureVectors [n,D] fp32, logits [n,C] fp32).
        keep: optional dict that receives the saved forward activations (tests check the backward chain against a
        reference backward taken over exactly these activations).  on_classifier_grads: called (no arguments) as soon
        as every classifier gradient is in the arena -- they are produced FIRST in the backward pass and are 88 % of the
        bytes, so their all-reduce can run under the whole conv-stack backward."""
        assert x_nhwc.dtype == torch.bfloat16 and x_nhwc.dim() == 4 and x_nhwc.shape[3] == self.c_pad, x_nhwc.shape
        n = x_nhwc.shape[0]
        W = [p.data for p in self.params]
        G = self.grads
        desc_dim = W[30].shape[0]
        if masks is None:
            masks = self.draw_masks(n, desc_dim)
        if not labels.is_cuda:
            # validated while still on the host (no device sync): torch's CrossEntropyLoss raises "Target out of bounds";
            # the kernel itself answers an out-of-range label with a NaN loss instead of reading past its buffers
            n_classes = W[32].shape[0]
            if labels.numel() and (int(labels.min()) < 0 or int(labels.max()) >= n_classes):
                raise VAError(f"labels must lie in [0, {n_classes}): got [{int(labels.min())}, {int(labels.max())}] "
                              "(the reference's class ids are 1-based and are used as-is, spatialModel.py:178)")
        labels = labels.to(device=x_nhwc.device, dtype=torch.int64).contiguous()

        # ---------------- forward (train mode), activations kept
        # per conv: (input, post-ReLU output, pool codes).  A pooled layer keeps 4 bits per window and channel (which
        # element takes the gradient) instead of its un-pooled output -- 1/16 of the bytes, and the backward pass does
        # not read the activation again; the outputs of the pooled layers are 80 % of the activation volume.
        saved = []
        x = x_nhwc
        for i in range(13):
            y = ops.conv2d_nhwc(x, W[2 * i], W[2 * i + 1], relu=True, pool=False)
            if POOL_AFTER[i]:
                nxt, codes = T.maxpool2x2(y, with_codes=True)
                saved.append((x, y if keep is not None else None, codes))
                x = nxt
            else:
                saved.append((x, y, None))
                x = y
            del y
        hw, ch = x.shape[1] * x.shape[2], x.shape[3]
        flat = T.transpose_bf16(x.view(n, hw, ch)).view(n, ch * hw)          # the reference's NCHW flatten order
        h1 = ops.linear(flat, W[26], W[27], relu=True)
        d1 = T.dropout(h1, masks[0], DROPOUT_P)
        h2 = ops.linear(d1, W[28], W[29], relu=True)
        d2 = T.dropout(h2, masks[1], DROPOUT_P)
        h3 = ops.linear(d2, W[30], W[31], relu=True, out_f32=True)           # descriptor layer: fp32 out
        d3 = T.dropout(h3, masks[2], DROPOUT_P)                              # featureVectors (reference :176)
        ce = T.ce_train(d3, W[32], W[33], labels, dw4=G[32], db4=G[33])
        if keep is not None:
            keep.update(conv=[(a, b) for a, b, _ in saved], flat=flat, h=[h1, h2, h3], d=[d1, d2, d3], dlogits=ce["dlogits"])

        # ---------------- backward
        g = T.dropout(ce["dx"], masks[2], DROPOUT_P)
        dz = T.relu_bwd_f32_to_bf16(g, h3)
        T.linear_wgrad(dz, d2, out=G[30]); T.bias_grad(dz, out=G[31])
        g = T.dropout(T.linear_dgrad(dz, W[30]), masks[1], DROPOUT_P)
        dz = T.relu_pool_bwd(g.view(n, 1, 1, -1), h2.view(n, 1, 1, -1), pooled=False, bias_grad_out=G[29]).view(n, -1)
        T.linear_wgrad(dz, d1, out=G[28])
        g = T.dropout(T.linear_dgrad(dz, W[28]), masks[0], DROPOUT_P)
        dz = T.relu_pool_bwd(g.view(n, 1, 1, -1), h1.view(n, 1, 1, -1), pooled=False, bias_grad_out=G[27]).view(n, -1)
        T.linear_wgrad(dz, flat, out=G[26])
        if on_classifier_grads is not None:
            on_classifier_grads()
        g = T.linear_dgrad(dz, W[26])                                        # [n, ch*hw] in NCHW flatten order
        g = T.transpose_bf16(g.view(n, ch, hw)).view(n, x.shape[1], x.shape[2], ch)
        for i in range(12, -1, -1):
            xin, y, codes = saved[i]
            if codes is not None:
                dz = T.pool_bwd_codes(g, codes, bias_grad_out=G[2 * i + 1])
            else:
                dz = T.relu_pool_bwd(g, y, pooled=False, bias_grad_out=G[2 * i + 1])
            cin = W[2 * i].shape[1]
            T.conv2d_wgrad(dz, xin, cin, out=G[2 * i])
            if i > 0:
                g = T.conv2d_dgrad(dz, W[2 * i])
            saved[i] = None
        return ce["loss"], d3, ce["logits"]

    def _allreduce_async(self, lo: int, hi: int):
        import torch.distributed as dist
        if self.flat_grad_bf16 is not None:
            T.f32_to_bf16_(self.flat_grad[lo:hi], self.flat_grad_bf16[lo:hi])      # one HBM pass; the update reads the bf16 sum
            if self._symm is not None:
                if lo != 0 or hi != self.flat_grad.numel():
                    raise VAError("allreduce_impl='va' reduces the whole arena in one launch (overlap_allreduce is an NCCL option)")
                return self._allreduce_own()
            return dist.all_reduce(self.flat_grad_bf16[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _allreduce_own(self):
        """The whole bf16 arena through va_allreduce_bf16 on the communication stream: barrier (every rank's arena is
        written) -> two-shot kernel over peer / multicast addresses -> barrier (every slice is stored everywhere).  Returns an
        object whose wait() makes the current stream wait for it, like a torch.distributed Work."""
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        done = torch.cuda.Event()
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ready)
            self._symm.barrier(channel=0)
            T.allreduce_bf16_(self._symm.buffer_ptrs, self._symm.multicast_ptr if self._symm_use_mc else 0, self._symm_world,
                              self._symm_rank, self.flat_grad_bf16.numel(), n_ctas=max(1, self.reserve_sms))
            self._symm.barrier(channel=0)
            done.record(self._comm_stream)

        class _Done:
            def wait(self_inner):
                torch.cuda.current_stream().wait_event(done)
        return _Done()

    def apply_update(self, pending=()):
        """Gradient all-reduce (when a process group is given) + the fused SGD-momentum update over the arena.
        pending: work handles of slices whose all-reduce is already in flight."""
        scale = 1.0
        if self.group is not None:
            import torch.distributed as dist
            if not pending:
                pending = [self._allreduce_async(0, self.flat_grad.numel())]
            for w in pending:
                w.wait()                       # makes the current stream wait for the collective
            scale = 1.0 / dist.get_world_size(self.group)
        lr, momentum = self.hyper()
        grad = self.flat_grad_bf16 if (self.group is not None and self.flat_grad_bf16 is not None) else self.flat_grad
        T.sgd_momentum_(self.flat_param, grad, self.flat_buf, lr=lr, momentum=momentum,
                        first_step=(self.steps_done == 0), grad_scale=scale)
        if self.steps_done == 0:
            self._publish_optimizer_state()
        self.steps_done += 1

    def step(self, x_nhwc: torch.Tensor, labels: torch.Tensor, masks=None):
        """Forward + backward + (overlapped) gradient all-reduce + update.  The arena is reduced in two slices: the
        classifier's gradients as soon as they exist (under the conv-stack backward), the conv stack's at the end."""
        pending = []
        split = self.offsets[26]               # first classifier tensor (state_dict order: 13 conv pairs, then 4 FC pairs)
        hook = None
        if self.group is not None and self.overlap_allreduce:
            def hook():
                pending.append(self._allreduce_async(split, self.flat_grad.numel()))
        self.flush()                           # a deferred update of the previous step: wait for its collective, then SGD
        loss, feat, logits = self.forward_backward(x_nhwc, labels, masks, on_classifier_grads=hook)
        if self.group is not None:
            pending.append(self._allreduce_async(0, split if self.overlap_allreduce else self.flat_grad.numel()))
        if self.defer_update:
            self._deferred = pending
            if self.reserve_sms > 0 and self.reserve_launches > 0:
                T.reserve_sms(self.reserve_sms, self.reserve_launches)     # the NEXT layer launches (the other stream's forward)
            return loss, feat, logits
        self.apply_update(pending)
        return loss, feat, logits

    def flush(self):
        """Apply a deferred update (no-op otherwise).  Called by step(); call it before reading parameters."""
        if self._deferred is not None:
            pending, self._deferred = self._deferred, None
            self.apply_update(pending)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        self.flush()
        return {k: p.data for k, p in zip(STATE_DICT_KEYS, self.params)}

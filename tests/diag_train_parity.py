"""Diagnostic (test infrastructure; run from the repo root: python tests/diag_train_parity.py [spatial|temporal] [batch]): per-tensor gradient agreement of the K5 training step with (a) the fp32 oracle and (b) a torch-autograd
emulation that rounds activations to bf16 at the same points as the kernels (so that ReLU masks and pool routing are
identical and only gradient rounding differs).  Test infrastructure only."""
import copy
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import two_stream as ts  # noqa: E402
from video_analytics_b200 import _lib  # noqa: E402
from video_analytics_b200._lib import check, ptr, stream_ptr  # noqa: E402
from video_analytics_b200.ops import STATE_DICT_KEYS  # noqa: E402
from video_analytics_b200.training import POOL_AFTER, StreamTrainer  # noqa: E402


class RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def emulated_step(model, ip, labels, masks):
    """fp32 autograd over bf16-rounded operands/activations (weights rounded per use, straight-through)."""
    r = RoundBF16.apply
    convs = [m for m in model.features if isinstance(m, torch.nn.Conv2d)]
    x = r(ip)
    for i, c in enumerate(convs):
        x = r(torch.relu(F.conv2d(x, r(c.weight), c.bias, padding=1)))
        if POOL_AFTER[i]:
            x = F.max_pool2d(x, 2, 2)
    x = x.flatten(1)
    fcs = [m for m in model.classifier if isinstance(m, torch.nn.Linear)]
    for j in range(3):
        x = torch.relu(F.linear(x, r(fcs[j].weight), fcs[j].bias))
        if j < 2:
            x = r(x)
        x = x * masks[j] * 2.0
    logits = F.linear(x, fcs[3].weight, fcs[3].bias)
    loss = F.cross_entropy(logits, labels)
    model.zero_grad()
    loss.backward()
    return loss.detach()


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "spatial"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cin, c_pad = (3, 16) if kind == "spatial" else (20, 32)
    oracle_model = ts.build_spatial_model(seed=21) if kind == "spatial" else ts.build_temporal_model(seed=21)
    ours_model = copy.deepcopy(oracle_model)
    emu_model = copy.deepcopy(oracle_model).cuda()
    opt_o = torch.optim.SGD(oracle_model.parameters(), 0.1, momentum=0.9)
    trainer = StreamTrainer(ours_model, None, c_pad=c_pad)
    g = torch.Generator().manual_seed(5)
    ip = torch.randn(n, cin, 224, 224, generator=g)
    labels = torch.randint(1, 101, (n,), generator=g)
    torch.manual_seed(100)
    masks = ts.draw_dropout_masks([(n, 4096), (n, 4096), (n, 256)])
    loss_o, fv_o, op_o = ts.train_step(oracle_model, opt_o, torch.nn.CrossEntropyLoss(), ip, labels, masks)
    loss_e = emulated_step(emu_model, ip.cuda(), labels.cuda(), [m.cuda() for m in masks])
    ipc = ip.cuda().float().contiguous()
    x = torch.empty((n, 224, 224, c_pad), dtype=torch.bfloat16, device="cuda")
    check(_lib.load().va_pack_input_nchw(ptr(ipc), n, cin, 224, 224, c_pad, ptr(x), stream_ptr()), "pack")
    loss, fv, logits = trainer.forward_backward(x, labels.cuda(), [m.to(torch.uint8).cuda() for m in masks])
    torch.cuda.synchronize()
    print(f"loss ours {float(loss):.6f} oracle {float(loss_o):.6f} emulated {float(loss_e):.6f}")
    go = dict(oracle_model.named_parameters())
    ge = dict(emu_model.named_parameters())

    def cmp(a, b):
        a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
        return (float((a - b).norm() / b.norm()), float(torch.dot(a, b) / (a.norm() * b.norm())), float(a.norm() / b.norm()))

    print(f"{'tensor':28s} | vs fp32 oracle: rel cos norm-ratio | vs bf16 emulation: rel cos norm-ratio | emu vs oracle rel")
    for k in STATE_DICT_KEYS:
        a = cmp(trainer.grad(k), go[k].grad)
        b = cmp(trainer.grad(k), ge[k].grad)
        c = cmp(ge[k].grad, go[k].grad)
        print(f"{k:28s} | {a[0]:.3e} {a[1]:.5f} {a[2]:.4f} | {b[0]:.3e} {b[1]:.5f} {b[2]:.4f} | {c[0]:.3e}")


main()

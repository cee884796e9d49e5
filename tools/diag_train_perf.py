"""Diagnostic: where the time of one training step goes (CUDA events around each primitive call)."""
import collections
import sys

import torch

sys.path.insert(0, ".")
from video_analytics_b200.spatialModel import build_spatial_torch_model  # noqa: E402
from video_analytics_b200.temporalModel import build_temporal_torch_model  # noqa: E402
from video_analytics_b200 import ops, train_ops as T  # noqa: E402
from video_analytics_b200.training import StreamTrainer  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "spatial"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cin, c_pad = (3, 16) if kind == "spatial" else (20, 32)
model = build_spatial_torch_model(101, 256, seed=1) if kind == "spatial" else build_temporal_torch_model(101, 10, 256, seed=1)
trainer = StreamTrainer(model, None, c_pad=c_pad)
x = torch.randn(n, 224, 224, c_pad, device="cuda").bfloat16()
x[..., cin:] = 0
labels = torch.randint(1, 101, (n,), device="cuda")

spans = []


def wrap(mod, name):
    fn = getattr(mod, name)

    def timed(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        shape = tuple(a[0].shape)
        spans.append((name, shape, e0, e1))
        return r
    setattr(mod, name, timed)


for _ in range(2):
    trainer.step(x, labels)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    trainer.step(x, labels)
e1.record()
torch.cuda.synchronize()
print(f"{kind} n={n}: {e0.elapsed_time(e1) / 3:.2f} ms/step  -> {n * 3 / e0.elapsed_time(e1) * 1e3:.1f} snippets/s; "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
for name in ("conv2d_nhwc", "linear"):
    wrap(ops, name)
for name in ("maxpool2x2", "relu_pool_bwd", "pool_bwd_codes", "bias_grad", "dropout", "conv2d_dgrad", "linear_dgrad", "conv2d_wgrad",
             "linear_wgrad", "ce_train", "relu_bwd_f32_to_bf16", "transpose_bf16", "sgd_momentum_"):
    wrap(T, name)
trainer.step(x, labels)
torch.cuda.synchronize()
tot = collections.OrderedDict()
for name, shape, a, b in spans:
    ms = a.elapsed_time(b)
    tot[name] = tot.get(name, 0.0) + ms
    if name in ("conv2d_wgrad", "conv2d_dgrad", "conv2d_nhwc", "linear_wgrad", "linear_dgrad", "linear", "relu_pool_bwd", "pool_bwd_codes", "maxpool2x2"):
        print(f"  {name:16s} {str(shape):28s} {ms:8.3f} ms")
print({k: round(v, 2) for k, v in tot.items()}, "sum", round(sum(tot.values()), 2))

#!/usr/bin/env python
"""Same-box yardstick (NOT product code): stock PyTorch / cuDNN / cuBLAS on the same B200 for the shapes of the path.

SURVEY.md 2.1 / BASELINE.md section 4 name "cuDNN/cuBLAS via stock PyTorch 2.11 on the same B200" as the implicit bar of
every kernel.  This tool times torchvision VGG16 (+ the reference's swapped classifier) in bf16 channels_last -- the fastest
stock configuration -- per layer and whole stream, forward at the bench's chunk size and forward+backward at the training
batch, with CUDA events, and prints one JSON object next to the tcgen05 numbers of the same run of this library.

    python tools/cudnn_yardstick.py [--batch 250] [--train-batch 256] > profiles/rNN_cudnn_yardstick.json
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F

CFG = [(64, False), (64, True), (128, False), (128, True), (256, False), (256, False), (256, True), (512, False), (512, False),
       (512, True), (512, False), (512, False), (512, True)]
NAMES = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3", "conv5_1",
         "conv5_2", "conv5_3"]


def ev_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=250)
    ap.add_argument("--train-batch", type=int, default=256)
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    dev = "cuda"
    out = {"what": "stock PyTorch %s (cuDNN %s) bf16 channels_last on %s" % (torch.__version__, torch.backends.cudnn.version(),
                                                                           torch.cuda.get_device_name(0)),
           "batch": a.batch, "layers": {}, "ours": {}}
    from video_analytics_b200 import ops
    for cin0, stream in ((3, "spatial"), (20, "temporal")):
        # ---- per layer: cuDNN conv + bias + ReLU (+ max-pool), bf16 channels_last, vs this library's layer kernel
        cin, H = cin0, 224
        rows = []
        for (cout, pool), name in zip(CFG, NAMES):
            if cin0 == 20 and name != "conv1_1":
                break                                          # layers 2..13 are identical for both streams
            conv = nn.Conv2d(cin, cout, 3, padding=1).to(dev).bfloat16().to(memory_format=torch.channels_last)
            x = torch.randn(a.batch, cin, H, H, device=dev).bfloat16().to(memory_format=torch.channels_last)

            def stock():
                y = F.relu(conv(x))
                return F.max_pool2d(y, 2, 2) if pool else y
            with torch.no_grad():
                ms_stock = ev_ms(stock)
            cin_pad = cin if cin % 64 == 0 else (16 if cin <= 16 else 32)
            xn = torch.zeros(a.batch, H, H, cin_pad, device=dev, dtype=torch.bfloat16)
            xn[..., :cin] = x.permute(0, 2, 3, 1)
            w, b = conv.weight.detach().float().contiguous(), conv.bias.detach().float().contiguous()
            ms_ours = ev_ms(lambda: ops.conv2d_nhwc(xn, w, b, relu=True, pool=pool))
            flops = 2.0 * a.batch * H * H * cout * 9 * cin
            rows.append({"layer": name + ("" if cin0 == 3 or name != "conv1_1" else " (20ch)"), "cudnn_ms": ms_stock, "ours_ms": ms_ours,
                         "cudnn_tflops": flops / ms_stock / 1e9, "ours_tflops": flops / ms_ours / 1e9, "speedup": ms_stock / ms_ours})
            del x, xn, conv
            cin = cout
            if pool:
                H //= 2
        out["layers"][stream] = rows
    # ---- whole stream forward (eval) and forward + backward (train), stock
    import torchvision.models as models
    for cin0, stream in ((3, "spatial"), (20, "temporal")):
        m = models.vgg16(weights=None)
        if cin0 != 3:
            m.features[0] = nn.Conv2d(cin0, 64, 3, padding=1)
        m.classifier = nn.Sequential(nn.Linear(25088, 4096), nn.ReLU(True), nn.Dropout(), nn.Linear(4096, 4096), nn.ReLU(True),
                                     nn.Dropout(), nn.Linear(4096, 256), nn.ReLU(True), nn.Dropout(), nn.Linear(256, 101))
        m = m.to(dev).bfloat16().to(memory_format=torch.channels_last)
        x = torch.randn(a.batch, cin0, 224, 224, device=dev).bfloat16().to(memory_format=torch.channels_last)
        m.eval()
        with torch.no_grad():
            ms_f = ev_ms(lambda: m.classifier(torch.flatten(m.features(x), 1)))
        flops = {3: 30_934_485_504, 20: 31_917_132_288}[cin0]
        res = {"forward_ms": ms_f, "forward_snippets_per_s": a.batch / ms_f * 1e3, "forward_tflops": a.batch * flops / ms_f / 1e9}
        del x
        m.train()
        xb = torch.randn(a.train_batch, cin0, 224, 224, device=dev).bfloat16().to(memory_format=torch.channels_last)
        lab = torch.randint(0, 101, (a.train_batch,), device=dev)
        opt = torch.optim.SGD(m.parameters(), 0.001, momentum=0.9)

        def train_step():
            opt.zero_grad(set_to_none=True)
            loss = F.cross_entropy(m.classifier(torch.flatten(m.features(xb), 1)).float(), lab)
            loss.backward()
            opt.step()
        ms_t = ev_ms(train_step, reps=4)
        res.update({"train_step_ms": ms_t, "train_snippets_per_s": a.train_batch / ms_t * 1e3, "train_tflops": 3 * a.train_batch * flops / ms_t / 1e9,
                    "train_note": "bf16 weights and optimizer state (no fp32 masters): cheaper than this library's fp32-master step"})
        out[stream] = res
        del m, xb, opt
        torch.cuda.empty_cache()
    # ---- this library, whole stream forward at the same batch
    from video_analytics_b200.spatialModel import build_spatial_torch_model
    from video_analytics_b200.temporalModel import build_temporal_torch_model
    for kind, cin0, stream, c_pad, build in ((0, 3, "spatial", 16, lambda: build_spatial_torch_model(101, 256, seed=0)),
                                            (1, 20, "temporal", 32, lambda: build_temporal_torch_model(101, 10, 256, seed=0))):
        net = ops.StreamNet(kind, cin0, 101, 256, max_batch=a.batch)
        net.load_state_dict(build().state_dict())
        x = torch.randn(a.batch, 224, 224, c_pad, device=dev).bfloat16()
        x[..., cin0:] = 0
        ms_o = ev_ms(lambda: net.forward(x, want_logits=False, want_pred=False))
        flops = {3: 30_934_485_504, 20: 31_917_132_288}[cin0]
        out["ours"][stream] = {"forward_ms": ms_o, "forward_snippets_per_s": a.batch / ms_o * 1e3, "forward_tflops": a.batch * flops / ms_o / 1e9,
                               "speedup_vs_stock_forward": out[stream]["forward_ms"] / ms_o}
        net.close()
        del x
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

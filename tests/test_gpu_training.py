"""K5: the training step on the CUDA kernels.

(1) vs the oracle's train_step (reference spatialModel.py:171-181 restated, pinned bit-equal to the reference's own
    loop-body lines by oracle/make_golden.py) on identical inputs, random-init weights and SHARED dropout masks.
    The forward runs in bf16 storage / fp32 accumulate, so ReLU masks and max-pool routing differ from the fp32
    reference at the few units whose pre-activation (or window maximum) is decided by the last bits.  A routing change
    moves a gradient entry by its full value, so the L2 distance of a gradient tensor to the fp32 one grows like
    sqrt(fraction of re-routed units) per layer -- tests/diag_train_parity.py shows torch autograd over a bf16-rounded forward
    is exactly as far from the fp32 oracle as these kernels are (features.0.weight: 0.42 vs 0.41).  What is stable, and
    what is asserted against the fp32 oracle:
        loss                      1e-3 relative      (measured 2e-5)
        featureVectors / logits   2e-2 of max |ref|
        last two layers' grads    relative L2 <= 2e-2; FC3's own ReLU gates classifier.6: a unit whose pre-activation is
                                  within bf16 noise of zero may open/close for one sample (one such flip is 1/256 of the
                                  active (sample, unit) pairs = 6e-2 of the norm), so rows of units whose gate differs
                                  from the oracle's are compared loosely, at most 3 of 256, and only where the
                                  activation itself is within the featureVector tolerance of zero
        every gradient tensor     norm within 15% of the oracle's, cosine >= 0.90   (grad norms)
        weights after one step    same bounds on (w_new - w_old) = -lr * grad
        loss of a second step     2e-2 relative (momentum branch, on weights that already differ slightly)
(2) vs a reference backward taken over the SAME saved forward activations (torch fp32 formulas, test-side): masks and
    routing are then identical and only the bf16 storage of the inter-layer gradients differs:
        every gradient tensor     relative L2 <= 3e-2
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu



def _pack(net_c_pad, ip):
    from video_analytics_b200 import _lib
    from video_analytics_b200._lib import check, ptr, stream_ptr
    ip = ip.cuda().float().contiguous()
    n, c, h, w = ip.shape
    x = torch.empty((n, h, w, net_c_pad), dtype=torch.bfloat16, device="cuda")
    check(_lib.load().va_pack_input_nchw(ptr(ip), n, c, h, w, net_c_pad, ptr(x), stream_ptr()), "va_pack_input_nchw")
    return x


def _compare(name, ours, ref, report, *, rel_tol=None, cos_tol=0.90, norm_tol=0.15, keep_rows=None):
    ours, ref = ours.detach().float().cpu(), ref.detach().float().cpu()
    if keep_rows is not None:          # drop the rows (FC3 units) whose ReLU gate differs from the oracle's
        ours, ref = ours[keep_rows], ref[keep_rows]
    ours, ref = ours.flatten(), ref.flatten()
    rel = float((ours - ref).norm() / ref.norm().clamp_min(1e-30))
    cos = float(torch.dot(ours, ref) / (ours.norm() * ref.norm()).clamp_min(1e-30))
    ratio = float(ours.norm() / ref.norm().clamp_min(1e-30))
    report.append((name, rel, cos, ratio))
    assert cos >= cos_tol and abs(ratio - 1.0) <= norm_tol, (name, rel, cos, ratio)
    if rel_tol is not None:
        assert rel <= rel_tol, (name, rel, cos, ratio)


TIGHT = ("classifier.6.weight", "classifier.6.bias", "classifier.9.weight", "classifier.9.bias")


@pytest.mark.parametrize("kind", ["spatial", "temporal"])
def test_train_step_vs_oracle(kind):
    from oracle import two_stream as ts
    from video_analytics_b200.ops import STATE_DICT_KEYS
    from video_analytics_b200.training import StreamTrainer
    n = 4
    cin, c_pad = (3, 16) if kind == "spatial" else (20, 32)
    oracle_model = ts.build_spatial_model(seed=21) if kind == "spatial" else ts.build_temporal_model(seed=21)
    ours_model = copy.deepcopy(oracle_model)
    w_old = {k: v.detach().clone() for k, v in oracle_model.state_dict().items()}
    opt_o = torch.optim.SGD(oracle_model.parameters(), 0.1, momentum=0.9)
    opt_m = torch.optim.SGD(ours_model.parameters(), 0.1, momentum=0.9)
    trainer = StreamTrainer(ours_model, opt_m, c_pad=c_pad)
    g = torch.Generator().manual_seed(5)
    report = []
    for it in range(2):
        ip = torch.randn(n, cin, 224, 224, generator=g)
        labels = torch.randint(1, 101, (n,), generator=g)
        torch.manual_seed(100 + it)
        masks = ts.draw_dropout_masks([(n, 4096), (n, 4096), (n, 256)])
        loss_o, fv_o, op_o = ts.train_step(oracle_model, opt_o, torch.nn.CrossEntropyLoss(), ip, labels, masks)
        loss, fv, logits = trainer.step(_pack(c_pad, ip), labels.cuda(), [m.to(torch.uint8).cuda() for m in masks])
        torch.cuda.synchronize()
        rel_loss = abs(float(loss) - float(loss_o)) / abs(float(loss_o))
        report.append((f"step{it}.loss", rel_loss, 1.0, 1.0))
        if it == 0:
            assert rel_loss < 1e-3, (float(loss), float(loss_o))
            assert float((fv.cpu() - fv_o).abs().max()) <= 2e-2 * float(fv_o.abs().max())
            assert float((logits.cpu() - op_o).abs().max()) <= 2e-2 * float(op_o.abs().max())
            # FC3 units whose ReLU gate differs from the fp32 oracle's for some sample (pre-activation ~ 0)
            gate_diff = (fv.cpu() > 0) != (fv_o > 0)
            flipped = gate_diff.any(0)
            assert int(flipped.sum()) <= 3, int(flipped.sum())
            keep6 = ~flipped
            report.append(("fc3.gate_flips", float(flipped.sum()), 1.0, 1.0))
            # gradients of the first step (the oracle's .grad are still in place after optimizer.step())
            grads_o = dict(oracle_model.named_parameters())
            for k in STATE_DICT_KEYS:
                _compare("grad." + k, trainer.grad(k), grads_o[k].grad, report, rel_tol=2e-2 if k in TIGHT else None,
                         keep_rows=keep6 if k.startswith("classifier.6.") else None)
            # weight delta after one step
            sd_o = oracle_model.state_dict()
            for k in STATE_DICT_KEYS:
                _compare("delta." + k, trainer.param(k).cpu() - w_old[k], sd_o[k] - w_old[k], report,
                         rel_tol=2e-2 if k in TIGHT else None, keep_rows=keep6 if k.startswith("classifier.6.") else None)
            # the momentum buffers are what torch.optim.SGD would checkpoint
            st = opt_m.state_dict()["state"]
            assert len(st) == 34 and all("momentum_buffer" in v for v in st.values())
        else:
            # second step runs the momentum branch on weights that already differ slightly: loss only, loosely
            assert rel_loss < 2e-2, (float(loss), float(loss_o))
    print("\n".join(f"{n_:40s} rel {r:.3e} cos {c:.6f} norm ratio {q:.4f}" for n_, r, c, q in report))


def test_backward_chain_given_forward():
    """Gradients vs torch fp32 backward formulas evaluated on the activations the kernels themselves saved."""
    import torch.nn.functional as F
    from oracle import two_stream as ts
    from video_analytics_b200.ops import STATE_DICT_KEYS
    from video_analytics_b200.training import POOL_AFTER, StreamTrainer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n, cin, c_pad = 3, 3, 16
    model = ts.build_spatial_model(seed=33)
    trainer = StreamTrainer(model, None, c_pad=c_pad)
    g = torch.Generator().manual_seed(6)
    ip = torch.randn(n, cin, 224, 224, generator=g)
    labels = torch.randint(1, 101, (n,), generator=g).cuda()
    torch.manual_seed(7)
    masks = [m.to(torch.uint8).cuda() for m in ts.draw_dropout_masks([(n, 4096), (n, 4096), (n, 256)])]
    keep = {}
    trainer.forward_backward(_pack(c_pad, ip), labels, masks, keep=keep)
    torch.cuda.synchronize()
    W = [p.data for p in trainer.params]
    bfw = lambda t: t.bfloat16().float()                                   # the GEMMs read bf16 weights
    ref = [None] * 34
    dl = keep["dlogits"]
    d1, d2, d3 = keep["d"]
    h1, h2, h3 = keep["h"]
    ref[32], ref[33] = dl.t() @ d3, dl.sum(0)
    gg = (dl @ W[32]) * masks[2].float() * 2.0
    dz = (gg * (h3 > 0)).bfloat16().float()
    ref[30], ref[31] = dz.t() @ d2.float(), dz.sum(0)
    gg = (dz @ bfw(W[30])).bfloat16().float() * masks[1].float() * 2.0
    dz = (gg * (h2.float() > 0)).bfloat16().float()
    ref[28], ref[29] = dz.t() @ d1.float(), dz.sum(0)
    gg = (dz @ bfw(W[28])).bfloat16().float() * masks[0].float() * 2.0
    dz = (gg * (h1.float() > 0)).bfloat16().float()
    ref[26], ref[27] = dz.t() @ keep["flat"].float(), dz.sum(0)
    gg = (dz @ bfw(W[26])).view(n, 512, 7, 7)                               # NCHW from here on, fp32, never rounded again
    for i in range(12, -1, -1):
        xin, y = keep["conv"][i]
        y_ = y.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        out = F.max_pool2d(y_, 2, 2) if POOL_AFTER[i] else y_ * 1.0
        out.backward(gg)
        dz = y_.grad * (y_.detach() > 0)
        x_ = xin.float().permute(0, 3, 1, 2)[:, :W[2 * i].shape[1]].contiguous()
        ref[2 * i] = torch.nn.grad.conv2d_weight(x_, W[2 * i].shape, dz, padding=1)
        ref[2 * i + 1] = dz.sum((0, 2, 3))
        if i > 0:
            gg = torch.nn.grad.conv2d_input(x_.shape, bfw(W[2 * i]), dz, padding=1)
    report = []
    for idx, k in enumerate(STATE_DICT_KEYS):
        _compare("chain." + k, trainer.grads[idx], ref[idx], report, rel_tol=3e-2, cos_tol=0.999, norm_tol=0.02)
    print("\n".join(f"{n_:40s} rel {r:.3e} cos {c:.6f} norm ratio {q:.4f}" for n_, r, c, q in report))


def test_network_train_epoch(tmp_path):
    """SpatialNetwork.train(): one epoch through the reference-facing API (loader batches, running consensus dict,
    checkpoint with torch-format optimizer state), then validate() on the updated weights."""
    import random
    from oracle import synth  # noqa: F401  (store contents come from the CUDA generator; oracle not needed here)
    from video_analytics_b200 import utils as U
    from video_analytics_b200.spatialModel import SpatialDataset, SpatialNetwork
    from video_analytics_b200.store import DeviceStore, make_layout
    lay = make_layout(6)
    store = DeviceStore(lay)
    lst, cls = tmp_path / "list.txt", tmp_path / "classInd.txt"
    lst.write_text("".join(lay.list_line(v, "train") for v in range(6)))
    cls.write_text("".join(f"{m.label} {m.category}\n" for m in lay.videos))
    sd = SpatialDataset(str(lst), None, U.getTransforms(), actionLabelLoc=str(cls), store=store)
    loader = U.getDataLoader(sd, batchSize=3)
    torch.manual_seed(0); random.seed(0)
    net = SpatialNetwork(101, 1, 0.01, 0.9, 256, loader, loader, [10], str(tmp_path / "ckp"), gpu=True, maxBatch=4)
    before = {k: v.detach().clone() for k, v in net.model.state_dict().items()}
    p0, l0 = net.validate()
    net.train()
    after = net.model.state_dict()
    assert all(k.startswith("module.") for k in after)
    moved = [k for k in after if not torch.equal(after[k].cpu(), before[k].cpu())]
    assert len(moved) == 34                                              # every tensor took a step (none frozen, N5)
    assert all(torch.isfinite(v).all() for v in after.values())
    assert len(net.trainDict) == 6
    ck = torch.load(net.resumeLoc, weights_only=False)
    assert len(ck["optimizer"]["state"]) == 34
    p1, l1 = net.validate()                                              # evaluation handle sees the new weights
    assert float(l1) != float(l0)
    assert net.resume()                                                  # optimizer state re-adopted into the arena
    net.train()


def test_data_parallel_gradients_two_gpus():
    """N>1: split batch + NCCL all-reduce of the gradient arena == single-GPU gradient of the whole batch; parameters
    identical on all ranks after the update (tools/check_train_ddp.py under torchrun)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29611", os.path.join(root, "tools", "check_train_ddp.py")],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert "DDP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_labels_out_of_range_are_rejected():
    """torch's CrossEntropyLoss raises on a target outside [0, C) (the reference's 1-based ids reach C = 101 with all
    classes in use); the trainer validates host labels before upload and the kernel answers device labels with NaN."""
    from video_analytics_b200 import train_ops as T
    from video_analytics_b200._lib import VAError
    from video_analytics_b200.spatialModel import build_spatial_torch_model
    from video_analytics_b200.training import StreamTrainer
    tr = StreamTrainer(build_spatial_torch_model(101, 256, seed=1), None, c_pad=16)
    x = torch.zeros(2, 224, 224, 16, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(VAError):
        tr.forward_backward(x, torch.tensor([3, 101]))
    with pytest.raises(VAError):
        tr.forward_backward(x, torch.tensor([-1, 5]))
    g = torch.Generator().manual_seed(2)
    feat = torch.rand(2, 256, generator=g).cuda()
    w4, b4 = torch.randn(101, 256, generator=g).cuda() * 0.05, torch.zeros(101).cuda()
    ce = T.ce_train(feat, w4, b4, torch.tensor([7, 101]).cuda())
    torch.cuda.synchronize()
    assert torch.isnan(ce["loss"]).all()
    assert float(ce["dlogits"][1].abs().max()) == 0.0 and float(ce["dlogits"][0].abs().max()) > 0.0

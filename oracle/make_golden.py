#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- pins the oracle (oracle/two_stream.py) to the reference and writes tests/golden/*.

Runs ONLY in the build container (needs /root/reference, which does not exist on the GPU box):
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py
What is executed from the reference itself (unmodified, imported or exec'd from its source text):
  * Sheet03/parameters.py, Sheet03/utils.py            -> imported as modules (py3-clean)
  * Sheet03/spatialModel.py class SpatialDataset       -> source lines exec'd (the class is py3-clean; the module is
                                                          not: print statements further down)
  * Sheet03/combinedModel.py def combineDescriptors    -> source lines exec'd (same reason)
  * Sheet03/spatialModel.py train() loop body (forward, CrossEntropyLoss, zero_grad/backward/step) -> source lines
    exec'd on a shim `self`, with real nn.Dropout modules
What cannot run (Python-2 only; restated in oracle/two_stream.py, unpinned by reference code):
  TemporalDataset.__getitem__ (`it.next()`, float randint), Spatial/TemporalNetwork (print statements, Variable/
  `.cuda(async=True)`), combinedModel.main (sklearn.externals).
Each check asserts equality between reference output and the restatement, then records compact fixtures
(indices + sha256 of tensor bytes + small arrays) that the CPU and GPU test suites replay without the reference.
"""
import hashlib
import io
import json
import os
import random
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/Sheet03"
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import synth, two_stream as ts  # noqa: E402
from video_analytics_b200.store import make_layout  # noqa: E402  (layout only: plain host bookkeeping)


def sha(t) -> str:
    a = t.detach().cpu().contiguous().numpy() if isinstance(t, torch.Tensor) else np.ascontiguousarray(t)
    return hashlib.sha256(a.tobytes()).hexdigest()


def import_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import parameters as ref_params  # noqa
    import utils as ref_utils  # noqa
    sys.path.remove(REF)
    return ref_params, ref_utils


def ref_source_lines(fname, first, last):
    with open(os.path.join(REF, fname)) as f:
        lines = f.readlines()
    return "".join(lines[first - 1:last])


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref_params, ref_utils = import_reference()
    report = {}

    # ---- (a) constants
    consts = {k: getattr(ref_params, k) for k in dir(ref_params) if k.isupper()}
    json.dump(consts, open(os.path.join(GOLD, "reference_parameters.json"), "w"), indent=1, sort_keys=True)
    assert consts["CROP_SIZE_TF"] == ts.CROP_SIZE_TF and consts["NORM_MEANS_TF"] == ts.NORM_MEANS_TF
    assert consts["NORM_STDS_TF"] == ts.NORM_STDS_TF and consts["VIDEO_INPUT_FLOW_COUNT"] == ts.VIDEO_INPUT_FLOW_COUNT
    assert consts["NACTION_CLASSES"] == ts.NACTION_CLASSES and consts["VIDEO_DESCRIPTOR_DIM"] == ts.VIDEO_DESCRIPTOR_DIM
    report["constants"] = len(consts)

    # ---- (b) videoInfo on every line of the reference's two video lists
    samples, digest, n_lines = {"train": [], "test": []}, hashlib.sha256(), 0
    for mode, fname in (("train", "demoTrain.txt"), ("test", "demoTest.txt")):
        with open(os.path.join(REF, fname)) as f:
            lines = f.readlines()
        for k, line in enumerate(lines):
            a = ref_utils.videoInfo(line, mode)
            b = ts.videoInfo(line, mode)
            assert tuple(a) == tuple(b), (line, a, b)
            digest.update(repr(tuple(a)).encode())
            n_lines += 1
            if k % (len(lines) // 12) == 0:
                samples[mode].append({"line": line, "expect": list(a)})
    json.dump({"n_lines": n_lines, "sha256_all": digest.hexdigest(), "samples": samples},
              open(os.path.join(GOLD, "videoinfo_samples.json"), "w"), indent=1)
    report["videoInfo_lines"] = n_lines

    # ---- (c) transform pipeline, bit-for-bit against the reference's getTransforms() on PIL images
    from PIL import Image
    cases = []
    tr_rgb = ref_utils.getTransforms()
    tr_flow = ref_utils.getTransforms(normMeans=[ts.NORM_MEANS_TF[0]], normStds=[ts.NORM_STDS_TF[0]])  # 2018 zip semantics
    for k in range(12):
        is_flow = k % 2 == 1
        shape = synth.FLOW_SHAPE if is_flow else synth.RGB_SHAPE
        if k >= 10:   # exactly crop-sized image: RandomCrop draws nothing
            shape = (224, 224, shape[2])
        seed = 100 + k
        img = synth.synth_image(synth.STORE_SEED + (1 if is_flow else 0), 7 * k + 3, shape)
        pil = Image.fromarray(img[:, :, 0] if is_flow else img)
        torch.manual_seed(seed)
        ref_out = (tr_flow if is_flow else tr_rgb)(pil)
        torch.manual_seed(seed)
        i, j, flip = ts.draw_transform_params(shape[0], shape[1])
        mean, std = (ts.flow_norm_constants(1) if is_flow else (ts.NORM_MEANS_TF, ts.NORM_STDS_TF))
        mine = ts.apply_transform(img, i, j, flip, mean, std)
        assert torch.equal(ref_out, mine), f"transform case {k} differs from the reference"
        cases.append(dict(seed=seed, store_seed=synth.STORE_SEED + (1 if is_flow else 0), image_id=7 * k + 3, shape=list(shape),
                          flow=is_flow, i=i, j=j, flip=flip, sha256=sha(mine), first8=[float(x) for x in mine.flatten()[:8]]))
    json.dump(cases, open(os.path.join(GOLD, "transform_cases.json"), "w"), indent=1)
    report["transform_cases_bit_exact"] = len(cases)

    # ---- (d) AverageMeter
    torch.manual_seed(5)
    vals = [torch.rand(256) for _ in range(25)]
    a, b = ref_utils.AverageMeter(), ts.AverageMeter()
    for v in vals:
        a.update(v)
        b.update(v)
    assert torch.equal(a.avg, b.avg) and a.count == b.count
    json.dump({"seed": 5, "n": 25, "dim": 256, "sha256_avg": sha(a.avg)}, open(os.path.join(GOLD, "average_meter.json"), "w"))
    report["average_meter"] = "bit-exact"

    # ---- (e) SpatialDataset.__getitem__: reference class source exec'd on a temp folder of JPEG frames
    ns = {}
    exec("from torch.utils.data import Dataset\nimport torchvision.transforms as transforms\nfrom PIL import Image\n"
         "import os, random\n", ns)
    ns.update({k: getattr(ref_params, k) for k in dir(ref_params) if k.isupper()})
    ns["videoInfo"] = ref_utils.videoInfo
    exec(ref_source_lines("spatialModel.py", 21, 81), ns)
    RefSpatialDataset = ns["SpatialDataset"]
    lay = make_layout(3)
    with tempfile.TemporaryDirectory() as tmp:
        frames_root = os.path.join(tmp, "frames")
        decoded = []
        for m in lay.videos:
            d = os.path.join(frames_root, m.category, m.name)
            os.makedirs(d)
            for k in range(m.n_frames):
                img = synth.synth_image(lay.seed, m.rgb_first + k, lay.rgb_shape)
                Image.fromarray(img).save(os.path.join(d, f"{k}.jpg"), quality=95)
                decoded.append(np.asarray(Image.open(os.path.join(d, f"{k}.jpg")).convert("RGB")))
        list_path = os.path.join(tmp, "list.txt")
        with open(list_path, "w") as f:
            for v in range(3):
                f.write(lay.list_line(v, "train"))
        cls_path = os.path.join(tmp, "classInd.txt")
        with open(cls_path, "w") as f:
            for m in lay.videos:
                f.write(f"{m.label} {m.category}\n")
        ref_ds = RefSpatialDataset(list_path, frames_root, ref_utils.getTransforms(), actionLabelLoc=cls_path)
        store = ts.OracleStore(lay, np.stack(decoded), np.zeros((0,) + tuple(lay.flow_shape), np.uint8))
        my_ds = ts.SpatialDataset(open(list_path).readlines(), store, actionLabelDict=ref_ds.actionLabelDict)
        recs = []
        for rep in range(4):
            for idx in range(3):
                random.seed(10 * rep + idx); torch.manual_seed(10 * rep + idx)
                ra = ref_ds[idx]
                random.seed(10 * rep + idx); torch.manual_seed(10 * rep + idx)
                rb = my_ds[idx]
                assert torch.equal(ra[0], rb[0]) and ra[1] == rb[1] and ra[2] == rb[2], (rep, idx)
                recs.append(dict(seed=10 * rep + idx, index=idx, **{k: v for k, v in my_ds.last_indices.items()}))
    json.dump(recs, open(os.path.join(GOLD, "spatial_dataset_draws.json"), "w"), indent=1)
    report["SpatialDataset_getitem_vs_reference_class"] = len(recs)

    # ---- (f) combineDescriptors: reference function source exec'd on CSVs written by the reference's own
    #          saveVideoDescriptors (utils.py:174-195)
    import pandas as pd
    ns2 = {"pd": pd, "np": np, "VIDEO_DESCRIPTOR_DIM": ref_params.VIDEO_DESCRIPTOR_DIM}
    exec(ref_source_lines("combinedModel.py", 9, 26), ns2)
    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(9)
        dicts = []
        for s in range(2):
            d = {}
            for v in (range(5) if s == 0 else (3, 1, 0, 4, 2, 7)):   # different orders; video 7 only in temporal
                m = ref_utils.AverageMeter()
                for _ in range(3):
                    m.update(torch.rand(256))
                d[f"v_Class{v:03d}_g01_c01"] = (m, torch.tensor(1 + v))
            dicts.append(d)
        ps, pt = os.path.join(tmp, "s.csv"), os.path.join(tmp, "t.csv")
        ref_utils.saveVideoDescriptors(dicts[0], ps)
        ref_utils.saveVideoDescriptors(dicts[1], pt)
        Xr, yr = ns2["combineDescriptors"](ps, pt)
        Xm, ym = ts.combineDescriptors(ps, pt)
        assert np.array_equal(Xr, Xm) and np.array_equal(yr, ym)
        # and the oracle's CSV writer reproduces the reference's bytes
        buf = io.StringIO()
        ts.saveVideoDescriptors(dicts[0], buf)
        assert buf.getvalue().replace("\r\n", "\n") == open(ps, newline="").read().replace("\r\n", "\n")
        json.dump({"seed": 9, "X_shape": list(Xr.shape), "labels": [int(v) for v in yr], "sha256_X": sha(Xr),
                   "csv_spatial": open(ps, newline="").read(), "csv_temporal": open(pt, newline="").read()},
                  open(os.path.join(GOLD, "combine_descriptors.json"), "w"))
    report["combineDescriptors_vs_reference_function"] = "equal"

    # ---- (h) training step: the reference's own loop-body lines (spatialModel.py train(), "op = self.features(ip)" ..
    #          "self.optimizer.step()") exec'd on a shim `self` with REAL nn.Dropout modules, vs ts.train_step with the
    #          masks ts.draw_dropout_masks draws from the same RNG state.  Two consecutive steps (momentum path).
    import copy
    import textwrap
    import types
    with open(os.path.join(REF, "spatialModel.py")) as f:
        lines = f.readlines()
    first = next(i for i, l in enumerate(lines) if "op = self.features(ip)" in l)
    last = next(i for i, l in enumerate(lines) if i > first and "self.optimizer.step()" in l)
    body = textwrap.dedent("".join(lines[first:last + 1]).expandtabs(4))
    model_r = ts.build_spatial_model(seed=5)
    model_o = copy.deepcopy(model_r)
    shim = types.SimpleNamespace(features=model_r.features, classifierList=list(model_r.classifier),
                                 classifierLen=len(list(model_r.classifier)), criterion=torch.nn.CrossEntropyLoss(),
                                 optimizer=torch.optim.SGD(model_r.parameters(), 0.1, momentum=0.9))
    opt_o = torch.optim.SGD(model_o.parameters(), 0.1, momentum=0.9)
    g = torch.Generator().manual_seed(77)
    steps = []
    for it in range(2):
        ip = torch.randn(2, 3, 224, 224, generator=g)
        labelVar = torch.randint(1, 101, (2,), generator=g)
        model_r.train()
        torch.manual_seed(1000 + it)
        ns3 = {"self": shim, "ip": ip, "labelVar": labelVar}
        exec(body, ns3)
        torch.manual_seed(1000 + it)
        masks = ts.draw_dropout_masks([(2, 4096), (2, 4096), (2, ref_params.VIDEO_DESCRIPTOR_DIM)])
        loss_o, fv_o, op_o = ts.train_step(model_o, opt_o, torch.nn.CrossEntropyLoss(), ip, labelVar, masks)
        assert torch.equal(ns3["loss"].detach(), loss_o), (it, float(ns3["loss"]), float(loss_o))
        assert torch.equal(ns3["featureVectors"].detach(), fv_o)
        for pr, po in zip(model_r.parameters(), model_o.parameters()):
            assert torch.equal(pr, po)
        steps.append(dict(loss=float(loss_o), fv_abs_mean=float(fv_o.abs().mean()), logits_abs_max=float(op_o.abs().max()),
                          w0_abs_mean=float(model_o.features[0].weight.abs().mean()),
                          w_last_abs_mean=float(model_o.classifier[9].weight.abs().mean())))
    json.dump({"model_seed": 5, "input_seed": 77, "mask_seed_base": 1000, "lr": 0.1, "momentum": 0.9, "batch": 2,
               "reference_lines": [first + 1, last + 1], "steps": steps},
              open(os.path.join(GOLD, "train_step.json"), "w"), indent=1)
    report["train_step_vs_reference_loop_body"] = [st["loss"] for st in steps]

    # ---- (i) model surgery: the reference's OWN __copyFirstLayer__ / __swapClassifier__ (temporalModel.py:149-181, py3-clean
    #          method bodies) exec'd as plain functions on a shim `self`, vs the restated builders, same seed -> every
    #          state_dict tensor bit-equal (the fresh conv1 bias and the classifier consume the RNG in the same order).
    import torchvision.models as tv_models
    with open(os.path.join(REF, "temporalModel.py")) as f:
        tlines = f.readlines()
    i0 = next(i for i, l in enumerate(tlines) if "def __copyFirstLayer__" in l)
    i1 = next(i for i, l in enumerate(tlines) if i > i0 and l.strip().startswith("def ") and "__swapClassifier__" not in l
              and "__copyFirstLayer__" not in l)
    surgery_src = textwrap.dedent("".join(tlines[i0:i1]).expandtabs(4))
    ns4 = {"nn": torch.nn}
    exec(surgery_src, ns4)
    surgery = {}
    for seed in (0, 3):
        torch.manual_seed(seed)
        shim_t = types.SimpleNamespace(model=tv_models.vgg16(weights=None), flowSampleSize=ref_params.VIDEO_INPUT_FLOW_COUNT,
                                       descriptorDim=ref_params.VIDEO_DESCRIPTOR_DIM, nActionClasses=ref_params.NACTION_CLASSES)
        ns4["__copyFirstLayer__"](shim_t)
        ns4["__swapClassifier__"](shim_t)
        mine_t = ts.build_temporal_model(seed=seed)
        sd_r, sd_m = shim_t.model.state_dict(), mine_t.state_dict()
        assert list(sd_r.keys()) == list(sd_m.keys())
        for k in sd_r:
            assert torch.equal(sd_r[k], sd_m[k]), ("temporal surgery", seed, k)
        # spatial: only the classifier swap (spatialModel.py:136-152 is the same method body)
        torch.manual_seed(seed)
        shim_s = types.SimpleNamespace(model=tv_models.vgg16(weights=None), descriptorDim=ref_params.VIDEO_DESCRIPTOR_DIM,
                                       nActionClasses=ref_params.NACTION_CLASSES)
        ns4["__swapClassifier__"](shim_s)
        mine_s = ts.build_spatial_model(seed=seed)
        for k, v in shim_s.model.state_dict().items():
            assert torch.equal(v, mine_s.state_dict()[k]), ("spatial surgery", seed, k)
        surgery[str(seed)] = {"temporal_conv1_w_sha256": sha(sd_m["features.0.weight"]), "temporal_conv1_b_sha256": sha(sd_m["features.0.bias"]),
                              "temporal_fc4_w_sha256": sha(sd_m["classifier.9.weight"]),
                              "spatial_fc1_w_sha256": sha(mine_s.state_dict()["classifier.0.weight"])}
    json.dump({"reference_lines": [i0 + 1, i1], "seeds": surgery}, open(os.path.join(GOLD, "model_surgery.json"), "w"), indent=1)
    report["model_surgery_vs_reference_methods"] = "bit-equal (seeds 0, 3)"

    # ---- (j) validate(): the reference's own loop-body lines (spatialModel.py:212-228: forward, summed loss, argmax, correct
    #          count, per-video AverageMeter update) exec'd on a shim, vs ts.forward_eval + ts.update_video_dict
    with open(os.path.join(REF, "spatialModel.py")) as f:
        lines = f.readlines()
    v0 = next(i for i, l in enumerate(lines) if "def validate" in l)
    first = next(i for i, l in enumerate(lines) if i > v0 and "op = self.features(ip)" in l)
    last = max(i for i, l in enumerate(lines) if i > first and i < first + 25 and "self.testDict[videoNames[i]][0].update(featureVectors[i])" in l)
    vbody = textwrap.dedent("".join(lines[first:last + 1]).expandtabs(4))
    model_v = ts.build_spatial_model(seed=6)
    model_v.eval()
    shim_v = types.SimpleNamespace(features=model_v.features, classifierList=list(model_v.classifier),
                                   classifierLen=len(list(model_v.classifier)), criterion=torch.nn.CrossEntropyLoss(), testDict={})
    my_dict = {}
    g = torch.Generator().manual_seed(78)
    val = {"loss": 0, "correct": 0}
    my_loss, my_correct = 0, 0
    names_all = [("v_A_g01_c01", "v_B_g01_c01"), ("v_B_g01_c01", "v_C_g01_c02")]
    for it in range(2):
        ip = torch.randn(2, 3, 224, 224, generator=g)
        labels = torch.randint(1, 101, (2,), generator=g)
        ns5 = {"self": shim_v, "ip": ip, "labelVar": labels, "labels": labels, "videoNames": names_all[it],
               "AverageMeter": ref_utils.AverageMeter, "loss": val["loss"], "correct": val["correct"]}
        with torch.no_grad():
            exec(vbody, ns5)
        val["loss"], val["correct"] = ns5["loss"], ns5["correct"]
        fv, op, pred = ts.forward_eval(model_v, ip)
        assert torch.equal(ns5["featureVectors"], fv) and torch.equal(ns5["op"], op) and torch.equal(ns5["pred"].view(-1), pred)
        my_loss = my_loss + torch.nn.CrossEntropyLoss()(op, labels)
        my_correct += int(pred.eq(labels).sum())
        ts.update_video_dict(my_dict, names_all[it], labels, fv)
    assert torch.equal(val["loss"], my_loss) and val["correct"] == my_correct
    assert sorted(shim_v.testDict) == sorted(my_dict)
    for k in my_dict:
        assert torch.equal(shim_v.testDict[k][0].avg, my_dict[k][0].avg) and shim_v.testDict[k][0].count == my_dict[k][0].count
    json.dump({"model_seed": 6, "input_seed": 78, "reference_lines": [first + 1, last + 1], "loss": float(my_loss), "correct": my_correct,
               "avg_sha256": {k: sha(v[0].avg) for k, v in sorted(my_dict.items())}},
              open(os.path.join(GOLD, "validate_body.json"), "w"), indent=1)
    report["validate_body_vs_reference_lines"] = "bit-equal (2 batches)"

    # ---- (g) oracle forward vectors (restated model; pins the oracle to itself across machines)
    lay = make_layout(2)
    rgb, flow = synth.build_store_numpy(lay)
    store = ts.OracleStore(lay, rgb, flow)
    for kind in ("spatial", "temporal"):
        model = ts.build_spatial_model(seed=0) if kind == "spatial" else ts.build_temporal_model(seed=0)
        name = lay.videos[0].name
        snips, recs = (ts.video_snippets_spatial if kind == "spatial" else ts.video_snippets_temporal)(store, name)
        sel = [0, 7, 131, 249]
        fv, logits, pred = ts.forward_eval(model, snips[sel])
        np.savez_compressed(os.path.join(GOLD, f"oracle_forward_{kind}.npz"), sel=np.array(sel), recs=np.array([recs[i] for i in sel]),
                            desc=fv.numpy(), logits=logits.numpy(), pred=pred.numpy(),
                            input_sha=np.array([sha(snips[i]) for i in sel]))
        report[f"oracle_forward_{kind}"] = [float(fv.abs().mean()), float(logits.abs().max())]
    json.dump(report, open(os.path.join(GOLD, "MANIFEST.json"), "w"), indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()

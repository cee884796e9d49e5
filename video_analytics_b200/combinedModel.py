"""Late fusion of the two streams -- API-compatible with the reference's `Sheet03/combinedModel.py`
(`combineDescriptors` :9-26, `main` :29-43) plus the explicit `CombinedModel` class named by the north star.

`combineDescriptors` keeps the CSV wire format (it is a one-off host-side join); per-video consensus, the 512-d
concatenation, LinearSVC scoring (argmax_c X.W_c + b_c) and the class-score average run in the CUDA fusion kernel
(`va_fuse`); fitting the SVM (reference :34-35) is `va_svm_fit`.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops
from .parameters import *  # noqa: F401,F403


def combineDescriptors(spatialCsv, temporalCsv):
    """reference combinedModel.py:9-26 -- inner-join the two descriptor CSVs on the video name and return
    (descriptors [V, 2*VIDEO_DESCRIPTOR_DIM] float64 = [spatial | temporal], labels [V] from the spatial file)."""
    import pandas as pd
    columns = ["vidname", "label"] + ["dim" + str(d) for d in range(VIDEO_DESCRIPTOR_DIM)]
    spatial = pd.read_csv(spatialCsv, names=columns)
    temporal = pd.read_csv(temporalCsv, names=columns)
    merged = pd.merge(spatial, temporal, on="vidname", how="inner", suffixes=("_s", "_t"))
    wanted = [c + "_s" for c in columns[2:]] + [c + "_t" for c in columns[2:]]
    return merged[wanted].values, merged["label_s"].values


class CombinedModel:
    """Two-stream late fusion on the GPU.

    * `fit(X, y)`: LinearSVC().fit of reference :34-35 on the device (`va_svm_fit`, fp64 dual coordinate descent).
    * `predict(X)`: reference :38 -- `classes_[argmax(X @ coef.T + intercept)]`, scored by `va_fuse` in fp64.
    * `fuse(...)`: the whole K4 step for a batch of videos straight from per-snippet network outputs.
    """

    def __init__(self, w_spatial: float = STREAM_WEIGHT_SPATIAL, w_temporal: float = STREAM_WEIGHT_TEMPORAL):
        self.w_s, self.w_t = float(w_spatial), float(w_temporal)
        self.classes_: Optional[np.ndarray] = None
        self.coef_ = self.intercept_ = None
        self._w_dev = self._b_dev = None

    # ---- SVM
    def fit(self, descriptors, labels, C: float = 1.0, intercept_scaling: float = 1.0, tol: float = 1e-4,
            max_iter: int = 1000):
        """reference :34-35 `svm.LinearSVC().fit(X, y)` -- same model (one-vs-rest, L2 penalty, squared hinge, regularised
        intercept) and the same solver family and stopping rule (LIBLINEAR's dual coordinate descent, projected-gradient
        range <= tol or max_iter epochs), run by `va_svm_fit` on the device in fp64.  `n_iter_` holds the epochs per class
        problem; reaching max_iter is scikit-learn's ConvergenceWarning case and is reported the same way."""
        y = np.asarray(labels.detach().cpu() if isinstance(labels, torch.Tensor) else labels)
        classes, index = np.unique(y, return_inverse=True)
        if len(classes) < 2:
            raise ValueError("This solver needs samples of at least 2 classes in the data, but the data contains only "
                             "one class: %r" % classes[0])
        X = descriptors if isinstance(descriptors, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(descriptors))
        X = X.to(device="cuda", dtype=torch.float64).contiguous()
        idx = torch.from_numpy(index.astype(np.int32)).cuda()
        coef, intercept, epochs = ops.svm_fit(X, idx, len(classes), C_reg=C, bias=intercept_scaling, tol=tol,
                                              max_iter=max_iter)
        self.n_iter_ = epochs.cpu().numpy()
        if int(self.n_iter_.max()) >= max_iter:
            import warnings
            warnings.warn("va_svm_fit reached max_iter before the tolerance (scikit-learn: 'Liblinear failed to "
                          "converge, increase the number of iterations.')", RuntimeWarning)
        return self.set_svm(coef.cpu().numpy(), intercept.cpu().numpy(), classes)

    def set_svm(self, coef, intercept, classes=None):
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        intercept = np.ascontiguousarray(intercept, dtype=np.float64)
        if coef.shape[0] == 1:
            # sklearn stores a single hyperplane for two classes: score > 0 -> classes_[1].  Expand to one-vs-rest
            # rows so that argmax reproduces predict().
            coef = np.concatenate([-coef, coef], axis=0)
            intercept = np.concatenate([-intercept, intercept], axis=0)
        self.coef_, self.intercept_ = coef, intercept
        self.classes_ = np.arange(coef.shape[0]) if classes is None else np.asarray(classes)
        self._w_dev = torch.from_numpy(coef).cuda()
        self._b_dev = torch.from_numpy(intercept).cuda()
        return self

    def as_sklearn(self):
        """The fitted model as a scikit-learn `LinearSVC` (coef_, intercept_, classes_ set as `fit` would leave them; the
        two-class case collapsed back to scikit-learn's single hyperplane), so that a file written by `main()` loads in a
        consumer of the reference's `joblib.dump(linearClassifier, SVM_FILE)` (combinedModel.py:36) and `.predict`s there."""
        if self.coef_ is None:
            raise ops.VAError("CombinedModel: no SVM set (call fit() or set_svm() first)")
        from sklearn import svm
        est = svm.LinearSVC()
        two = len(self.classes_) == 2
        est.coef_ = np.ascontiguousarray(self.coef_[1:2] if two else self.coef_)
        est.intercept_ = np.ascontiguousarray(self.intercept_[1:2] if two else self.intercept_)
        est.classes_ = np.asarray(self.classes_)
        est.n_features_in_ = int(self.coef_.shape[1])
        est.n_iter_ = int(np.max(getattr(self, "n_iter_", 0)))
        return est

    def decision_function(self, descriptors) -> torch.Tensor:
        """[V, C_svm] fp64 scores on the device for already-fused descriptors [V, 2D] (numpy or tensor), scored in fp64 on
        the fp64 values -- what `linearClassifier.predict(svmTestData)` (reference :38) sees after pandas parsed the CSVs."""
        if self._w_dev is None:
            raise ops.VAError("CombinedModel: no SVM set (call fit() or set_svm() first)")
        X = torch.as_tensor(np.asarray(descriptors) if not isinstance(descriptors, torch.Tensor) else descriptors)
        X = X.to(device="cuda", dtype=torch.float64).contiguous()
        scores, pred = ops.svm_decision(X, self._w_dev, self._b_dev)
        self._last = {"svm_scores": scores, "svm_pred": pred}
        return scores

    def predict(self, descriptors) -> np.ndarray:
        self.decision_function(descriptors)
        idx = self._last["svm_pred"].cpu().numpy()
        return self.classes_[idx]

    # ---- full K4 step
    def fuse(self, desc_s, desc_t, score_s, score_t, video_offsets, out=None):
        """Per-video consensus over each video's snippets + late fusion; see ops.fuse."""
        return ops.fuse(desc_s, desc_t, score_s, score_t, video_offsets, svm_w=self._w_dev, svm_b=self._b_dev,
                        w_s=self.w_s, w_t=self.w_t, out=out)


def main():
    """reference combinedModel.py:29-43 with the GPU scorer."""
    import joblib
    trainX, trainY = combineDescriptors(SPATIAL_TRAIN_CSV_LOC, TEMPORAL_TRAIN_CSV_LOC)
    testX, testY = combineDescriptors(SPATIAL_TEST_CSV_LOC, TEMPORAL_TEST_CSV_LOC)
    model = CombinedModel().fit(trainX, trainY)
    joblib.dump(model.as_sklearn(), SVM_FILE)                 # reference :36: a LinearSVC object, loadable by its consumers
    preds = model.predict(testX)
    acc = sum(int(p == a) for p, a in zip(preds, testY))
    print("accuracy = %f percent" % ((acc * 100.0) / len(testY)))


if __name__ == "__main__":
    main()

"""va_svm_fit at the fusion problem's full size (3783 videos x 512-d x 101 classes) next to scikit-learn's LinearSVC
(the reference's call, combinedModel.py:34-35) on the host cores.  `python tools/bench_svm.py [--no-sklearn]`."""
import argparse
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analytics_b200 import ops                          # noqa: E402
from video_analytics_b200.combinedModel import CombinedModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-sklearn", action="store_true")
    a = ap.parse_args()
    g = torch.Generator().manual_seed(7)
    V, F, K = 3783, 512, 101
    cent = torch.randn(K, F, generator=g, dtype=torch.float64).abs()
    lab = torch.randint(0, K, (V,), generator=g)
    X = (cent[lab] * 0.5 + 0.5 * torch.randn(V, F, generator=g, dtype=torch.float64)).clamp_min(0)
    Xd, idx = X.cuda(), lab.to(torch.int32).cuda()
    ops.svm_fit(Xd, idx, K, max_iter=2)
    torch.cuda.synchronize()
    for tol in (1e-4, 1e-6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        coef, ic, epochs = ops.svm_fit(Xd, idx, K, tol=tol)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ep = epochs.cpu().numpy()
        pred = (coef @ Xd.T + ic[:, None]).argmax(0)
        print(f"va_svm_fit tol={tol:g}: {ms:.1f} ms, epochs min/mean/max {ep.min()}/{ep.mean():.0f}/{ep.max()} "
              f"(an epoch visits the ACTIVE samples only), train acc {(pred == idx).double().mean():.4f}")
        if tol == 1e-4:
            c_gpu = coef.cpu().numpy()
    t0 = time.perf_counter()
    m = CombinedModel().fit(X.numpy(), lab.numpy())
    print(f"CombinedModel.fit (host arrays in, H2D + fit + D2H): {1e3 * (time.perf_counter() - t0):.1f} ms")
    if not a.no_sklearn:
        from sklearn import svm
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t0 = time.perf_counter()
            clf = svm.LinearSVC(dual=True).fit(X.numpy(), lab.numpy())
            dt = time.perf_counter() - t0
        print(f"sklearn LinearSVC() (LIBLINEAR dual CD, 1 host core of {os.cpu_count()}): {1e3 * dt:.0f} ms, n_iter {clf.n_iter_}, "
              f"coef max|gpu - sklearn| {np.abs(clf.coef_ - c_gpu).max():.2e}, "
              f"predictions equal: {bool((clf.predict(X.numpy()) == m.predict(X.numpy())).all())}")


if __name__ == "__main__":
    main()

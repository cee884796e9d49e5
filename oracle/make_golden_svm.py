"""TEST INFRASTRUCTURE -- pins oracle/svm_fit.py against scikit-learn's LinearSVC (the library the reference calls at
Sheet03/combinedModel.py:34-35) run in the build container, and writes tests/golden/svm_fit.npz.

    python oracle/make_golden_svm.py

The fixture holds a seeded 5-class problem of non-negative "descriptors" (the real ones are post-ReLU means), a
two-class relabelling of it, and scikit-learn's converged coefficients from LIBLINEAR's primal solver (dual=False,
tol 1e-12): its dual solver -- randomly ordered, with shrinking -- does not reach such a tolerance on this data within
2e6 epochs on harder data (it ends 5e-3 away); the coefficients of the reference's own default call (tol 1e-4) are
stored too, as the measure of how far the reference itself is from the optimum.
The one-vs-rest problems are strictly convex, so primal and dual solvers share one optimum.
"""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    import sklearn
    from sklearn import svm
    from oracle import svm_fit as sf
    rng = np.random.default_rng(5)
    V, F, K = 160, 24, 5
    centres = rng.normal(size=(K, F))
    labels = rng.integers(1, K + 1, V)                     # 1-based like the reference's action labels
    X = np.abs(centres[labels - 1] * 1.5 + rng.normal(size=(V, F)) * 0.5)
    labels2 = np.where(labels > 3, 7, 2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        multi = svm.LinearSVC(dual=False, tol=1e-12, max_iter=2_000_000).fit(X, labels)
        binary = svm.LinearSVC(dual=False, tol=1e-12, max_iter=2_000_000).fit(X, labels2)
        default = svm.LinearSVC(dual=True).fit(X, labels)   # the reference's call (random visiting order, shrinking)
    coef, ic, classes, epochs = sf.fit_linear_svc(X, labels, tol=1e-10, max_iter=20000)
    coef2, ic2, classes2, epochs2 = sf.fit_linear_svc(X, labels2, tol=1e-10, max_iter=20000)
    report = {
        "sklearn": sklearn.__version__,
        "oracle_vs_sklearn_multi_coef_maxabs": float(np.abs(coef - multi.coef_).max()),
        "oracle_vs_sklearn_multi_intercept_maxabs": float(np.abs(ic - multi.intercept_).max()),
        "oracle_vs_sklearn_binary_coef_maxabs": float(np.abs(coef2 - binary.coef_).max()),
        "oracle_vs_sklearn_binary_intercept_maxabs": float(np.abs(ic2 - binary.intercept_).max()),
        "oracle_epochs_multi": [int(e) for e in epochs], "oracle_epochs_binary": [int(e) for e in epochs2],
        "sklearn_default_call_n_iter": int(np.max(default.n_iter_)),
        "sklearn_default_vs_converged_coef_maxabs": float(np.abs(default.coef_ - multi.coef_).max()),
    }
    assert report["oracle_vs_sklearn_multi_coef_maxabs"] < 1e-6 and report["oracle_vs_sklearn_binary_coef_maxabs"] < 1e-6
    assert (sf.decision(X, coef, ic, classes)[1] == multi.predict(X)).all()
    assert (sf.decision(X, coef2, ic2, classes2)[1] == binary.predict(X)).all()
    np.savez_compressed(os.path.join(GOLD, "svm_fit.npz"), X=X, labels=labels, labels2=labels2,
                        sk_coef=multi.coef_, sk_intercept=multi.intercept_, sk_classes=multi.classes_,
                        sk_coef2=binary.coef_, sk_intercept2=binary.intercept_, sk_classes2=binary.classes_,
                        sk_default_coef=default.coef_, sk_default_intercept=default.intercept_)
    mpath = os.path.join(GOLD, "MANIFEST.json")
    manifest = json.load(open(mpath))
    manifest["svm_fit_vs_sklearn_LinearSVC"] = report
    json.dump(manifest, open(mpath, "w"), indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()

"""Shared synthetic inputs for the tests (seeded; same bytes on every box)."""
import numpy as np


def synth(n, m, npos, seed, shift_feats=5, discretize=False):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, m))
    y = np.zeros(n, np.uint32)
    y[rng.choice(n, npos, replace=False)] = 1
    X[y == 1, :shift_feats] += 1.0
    if discretize:
        X[:, ::5] = np.round(X[:, ::5] * 2) / 2   # heavy value ties in every 5th column
    return X, y


def folds_strat(y, nF, seed):
    rng = np.random.default_rng(seed)
    f = np.zeros(len(y), np.uint32)
    for c in (0, 1):
        idx = np.flatnonzero(y == c)
        rng.shuffle(idx)
        f[idx] = np.arange(len(idx)) % nF
    return f


def trees_equal(a, b):
    """bit-exact comparison of two trees in ranger's getter layout"""
    if len(a["var"]) != len(b["var"]):
        return False, "node count %d vs %d" % (len(a["var"]), len(b["var"]))
    for k in ("left", "right", "var"):
        if not np.array_equal(a[k], b[k]):
            i = int(np.flatnonzero(a[k] != b[k])[0])
            return False, "%s differs first at node %d: %s vs %s" % (k, i, a[k][i], b[k][i])
    if not np.array_equal(a["val"].view(np.uint64), b["val"].view(np.uint64)):
        i = int(np.flatnonzero(a["val"].view(np.uint64) != b["val"].view(np.uint64))[0])
        return False, "split value differs first at node %d: %r vs %r" % (i, a["val"][i], b["val"][i])
    fa, fb = a["frac"].copy(), b["frac"].copy()
    na, nb = np.isnan(fa), np.isnan(fb)
    if not np.array_equal(na, nb):
        return False, "terminal/internal pattern differs"
    fa[na] = 0; fb[nb] = 0
    if not np.array_equal(fa.view(np.uint64), fb.view(np.uint64)):
        return False, "leaf fractions differ"
    return True, ""


CASES = {
    # name: (n, m, npos, nFolds, nParts, fp, ratio, k, T, mtry, discretize)
    "fy_26": (3000, 26, 100, 5, 4, 2, 3, 5, 10, 5, False),
    "reject_300": (2000, 300, 150, 3, 3, 1, 1, 10, 4, 17, False),
    "ties_ratio0": (3000, 26, 120, 3, 7, 3, 0, 3, 5, 5, True),
}


def make_case(name):
    n, m, npos, nF, nParts, fp, ratio, k, T, mtry, disc = CASES[name]
    X, y = synth(n, m, npos, 1, discretize=disc)
    f = folds_strat(y, nF, 2)
    return dict(X=X, y=y, f=f, n=n, m=m, nF=nF, nParts=nParts, fp=fp, ratio=ratio, k=k, T=T, mtry=mtry)

"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libparsmurf_ref.so, built by oracle/Makefile from
/root/reference).  Run in the build container only:  python scripts/make_golden.py
Everything stored here is an OUTPUT of the reference classes at ensThrd = 1 (the only deterministic mode)
or an input needed to reproduce it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as O   # noqa: E402
from helpers import CASES, make_case   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"


def fold_metrics(ctx, y, c1, nF):
    fm = np.zeros((nF, 2))
    for f in range(nF):
        fd = ctx.fold(f)
        idx = np.concatenate([fd["testPos"], fd["testNeg"]])
        fm[f] = O.r_curves(y[idx], c1[idx])
    return fm


# ---- config 1: exampleDataset (data1.txt is stored 30 x 1000 = transposed) through cfgEx/dataFromFile.json's parameters
d = np.loadtxt("/root/reference/exampleDataset/data1.txt")
X = np.ascontiguousarray(d.T)
y = np.loadtxt("/root/reference/exampleDataset/labels1.txt").astype(np.uint32)
f = np.loadtxt("/root/reference/exampleDataset/folds1.txt").astype(np.uint32)
x = O.make_x(X, y)
params = dict(nFolds=10, nParts=10, fp=1, ratio=1, k=5, nTrees=10, mtry=5, seed=1)
secs, c1, c2 = O.r_cv(x, y, f, params["nFolds"], params["nParts"], params["fp"], params["ratio"], params["k"], params["nTrees"], params["mtry"], params["seed"])
ctx = O.RefCtx(x, y, f, params["nFolds"])
auroc, auprc = O.r_curves(y, c1)
np.savez_compressed(os.path.join(OUT, "cfg1_example_transposed.npz"), X=X.astype(np.float64), y=y, folds=f, class1=c1, class2=c2,
                    auroc=auroc, auprc=auprc, fold_metrics=fold_metrics(ctx, y, c1, 10), **{k: np.uint32(v) for k, v in params.items()})
print("cfg1: AUROC %.6f AUPRC %.6f" % (auroc, auprc))

# ---- synthetic cases of tests/helpers.py (inputs are regenerated from the seed; only reference outputs are stored)
for name in CASES:
    c = make_case(name)
    x = O.make_x(c["X"], c["y"])
    secs, c1, c2 = O.r_cv(x, c["y"], c["f"], c["nF"], c["nParts"], c["fp"], c["ratio"], c["k"], c["T"], c["mtry"], 7)
    ctx = O.RefCtx(x, c["y"], c["f"], c["nF"])
    fd = ctx.fold(0)
    # SamplerKNN output + the rand() draws it consumed, partition 0 and the last partition of fold 0
    tr0, sz0, dr0 = ctx.sample(0, 0, c["nParts"], c["fp"], c["ratio"], c["k"])
    trL, szL, drL = ctx.sample(0, c["nParts"] - 1, c["nParts"], c["fp"], c["ratio"], c["k"])
    # exact kNN table of fold 0 (ANN kd-tree and ANN brute force)
    pts = x[fd["trngPos"]]
    kk = min(c["k"], len(pts))
    nn_kd, dd_kd = O.r_knn(pts, kk, brute=False)
    nn_bf, _ = O.r_knn(pts, kk, brute=True)
    # forest of partition 0 (seed = 7 + 0 + 0*nParts), first and last tree, and its predictions on fold 0's test rows
    forest = O.r_forest_train(np.ascontiguousarray(tr0.T), c["T"], c["mtry"], 7)
    t0, tL = forest.tree(0), forest.tree(c["T"] - 1)
    test_idx = np.concatenate([fd["testPos"], fd["testNeg"]])
    test = x[test_idx].copy(); test[:, -1] = 0
    preds = O.r_forest_predict(forest, np.ascontiguousarray(test.T))
    auroc, auprc = O.r_curves(c["y"], c1)
    np.savez_compressed(os.path.join(OUT, "case_%s.npz" % name), class1=c1, class2=c2, auroc=auroc, auprc=auprc,
                        fold_metrics=fold_metrics(ctx, c["y"], c1, c["nF"]),
                        trngPos=fd["trngPos"], trngNeg_head=fd["trngNeg"][:64], trngNeg_crc=np.uint64(int(fd["trngNeg"].astype(np.uint64).dot(np.arange(1, len(fd["trngNeg"]) + 1, dtype=np.uint64)) % (1 << 61))),
                        train0=tr0, draws0=dr0, sizes0=np.array([sz0["trngPos"], sz0["trngNeg"], sz0["lineLen"]], np.uint32),
                        trainL_crc=np.float64(np.sum(trL * (1 + np.arange(trL.shape[1])))), sizesL=np.array([szL["trngPos"], szL["trngNeg"], szL["lineLen"]], np.uint32),
                        nn_kd=nn_kd, dd_kd=dd_kd, nn_bf=nn_bf,
                        **{"t0_" + k: v for k, v in t0.items()}, **{"tL_" + k: v for k, v in tL.items()}, preds0=preds)
    print(name, "AUROC %.6f AUPRC %.6f" % (auroc, auprc), "nodes", len(t0["var"]), len(tL["var"]), "kd==brute", np.array_equal(nn_kd, nn_bf))
# ---- grid search (cfgEx/gridTune.json shape: optimizer "grid") on the fy_26 case, three grid points
c = make_case("fy_26")
x = O.make_x(c["X"], c["y"])
grid = [(4, 1, 1, 5, 5, 5), (4, 2, 1, 5, 5, 5), (6, 1, 2, 5, 5, 7)]
c1, c2, gm = O.r_cv_grid(x, c["y"], c["f"], c["nF"], grid, 7)
au = O.r_curves(c["y"], c1)
np.savez_compressed(os.path.join(OUT, "grid_fy_26.npz"), class1=c1, class2=c2, grid=np.array(grid, np.uint32), grid_metrics=gm, auroc=au[0], auprc=au[1])
print("grid", gm.tolist(), au)
print(os.listdir(OUT))

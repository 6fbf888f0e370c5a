"""GPU parity at the BASELINE shapes (Missing #2): the CUDA path, through the C ABI, against the unmodified
reference classes (oracle/_ref) on the benchmarked partition geometry -- not only on toy sizes.

  cfg-2 shape  1M x 26, 1000 positives, 5 folds: P = 800, Ntr = 9600, Nte = 200 000; 4 partitions of fold 0
  cfg-3 shape  1.2M x 26, 400 positives, 20 folds, nParts 320: P = 380, Ntr = 4560, Nte = 60 000; 304 partitions under a
               batch budget that cuts them into several batches
  cfg-4 shape  60k x 1000, 10 000 positives, 5 folds: P = 8000 (tensor-core kNN), Ntr = 28 000 (8-byte list entries, radix
               feature sort, rejection-sampled mtry = 31); one partition, stage by stage

The reference runs multi-threaded; every partition's SMOTE draws are recorded and injected (oracle/ref_harness.cpp:
ref_run_partitions_ex, pinned on the CPU by tests/test_ref_export.py), so the comparison is bit for bit."""
import os

import numpy as np
import pytest

from helpers import trees_equal

pytestmark = pytest.mark.gpu


def _data(n, m, npos, nF, shift, seed=1):
    rng = np.random.default_rng(seed)
    x = np.empty((n, m + 1))
    x[:, :m] = rng.standard_normal((n, m))
    y = np.zeros(n, np.uint32)
    y[rng.choice(n, npos, replace=False)] = 1
    x[y == 1, :shift] += 1.0
    x[:, m] = np.where(y > 0, 1.0, 2.0)
    f = np.zeros(n, np.uint32)
    for c in (0, 1):
        idx = np.flatnonzero(y == c)
        rng.shuffle(idx)
        f[idx] = np.arange(len(idx)) % nF
    return x, y, f


def _threads():
    return max(1, min(os.cpu_count() or 1, 32))


def test_cfg2_shape_scores_equal_the_reference_bit_for_bit(ctx, oracle):
    O = oracle
    if not O.have_ref():
        pytest.skip("oracle/_ref did not travel")
    n, m, nF, nParts, fp, ratio, k, T, mtry, seed = 1_000_000, 26, 5, 100, 2, 3, 5, 10, 5, 1
    x, y, f = _data(n, m, 1000, nF, 5)
    ref = O.RefCtx(x, y, f, nF)
    secs, c1, c2, draws, preds, fd = ref.run_partitions_ex(0, 0, 4, nParts, fp, ratio, k, T, mtry, seed, 4)
    assert len(fd["trngPos"]) == 800 and preds.shape[1] == 200_000
    ctx.dataset_upload(x)
    r = O.scale_parity(ctx, fd, draws, preds, 0, 4, nParts, fp, ratio, k, T, mtry, seed, 0, ref_class1=c1)
    assert r["mismatching_rows"] == 0 and r["max_abs_diff"] == 0.0, r
    assert r["max_abs_diff_vs_reference_sum"] < 1e-15


def test_cfg3_shape_budget_split_batches_equal_the_reference(ctx, oracle):
    O = oracle
    if not O.have_ref():
        pytest.skip("oracle/_ref did not travel")
    n, m, nF, nParts, fp, ratio, k, T, mtry, seed = 1_200_000, 26, 20, 320, 2, 3, 5, 10, 5, 1
    x, y, f = _data(n, m, 400, nF, 5)
    ref = O.RefCtx(x, y, f, nF)
    p1 = 304
    secs, c1, c2, draws, preds, fd = ref.run_partitions_ex(0, 0, p1, nParts, fp, ratio, k, T, mtry, seed, _threads())
    assert len(fd["trngPos"]) == 380 and preds.shape[1] == 60_000
    ctx.dataset_upload(x)
    ctx.fold_begin(fd["trngPos"], fd["trngNeg"], fd["testPos"], fd["testNeg"])
    ctx.set_batch_budget(1 << 30)                       # 1 GiB: the 304 partitions need several batches
    try:
        cap = ctx.batch_capacity(nParts, fp, ratio, T, mtry, 0)
        assert 1 <= cap < p1 / 2, cap
        r = O.scale_parity(ctx, fd, draws, preds, 0, p1, nParts, fp, ratio, k, T, mtry, seed, 0, ref_class1=c1)
    finally:
        ctx.set_batch_budget(0)
    assert r["partitions"] == p1 and r["mismatching_rows"] == 0 and r["max_abs_diff"] == 0.0, r


def test_cfg4_shape_stage_by_stage(ctx, oracle, monkeypatch):
    O = oracle
    n, m, nF, nParts, fp, ratio, k, T, mtry, seed = 60_000, 1000, 5, 10, 2, 3, 10, 10, 31, 1
    x, y, f = _data(n, m, 10_000, nF, 50)
    rnd = O.GlibcRand(1)
    fd = O.o_ttd(y, f, nF, rnd)[0]
    P = len(fd["trngPos"])
    assert P == 8000
    ctx.dataset_upload(x)
    ctx.fold_begin(fd["trngPos"], fd["trngNeg"], fd["testPos"], fd["testNeg"])
    # (1) neighbour table: the tensor-core path must be taken and proven complete; the table equals the exact CUDA kernel's
    # (indices and fp64 distances) and, on sampled queries, the reference's own ANN kd-tree search
    monkeypatch.delenv("PSM_KNN_TENSOR", raising=False)
    ctx.fold_knn(k)
    path, bad = ctx.fold_knn_path()
    assert path == 1 and bad == 0
    nn, dist = ctx.fold_knn_get()
    monkeypatch.setenv("PSM_KNN_TENSOR", "0")
    ctx.fold_knn(k)
    nn0, dist0 = ctx.fold_knn_get()
    monkeypatch.delenv("PSM_KNN_TENSOR", raising=False)
    assert np.array_equal(nn, nn0) and np.array_equal(dist.view(np.uint64), dist0.view(np.uint64))
    if O.have_ref():
        pts = x[fd["trngPos"]]
        q = np.random.default_rng(3).choice(P, 48, replace=False)
        ri, rd = O.r_knn_queries(pts, k, q)
        assert np.array_equal(ri, nn[q]) and np.array_equal(rd.view(np.uint64), dist[q].view(np.uint64))
    # (2) partition 0: training matrix bit-exact against the oracle's SMOTE restatement with the same draws
    per = P * fp * (m + 1)
    draws = rnd.fill(per)
    ctx.batch_begin(nParts, fp, ratio, 0, 1)
    ctx.batch_assemble(draws)
    got, sz = ctx.batch_get_training(0)
    assert sz["Ntr"] == 28_000                     # > 16384 rows: 8-byte list entries and the radix feature sort
    rc, osz = O.o_partition_sizes(P, len(fd["trngNeg"]), nParts, fp, ratio, 0)
    rc, want = O.o_assemble(x, fd["trngPos"], fd["trngNeg"][osz["negOffset"]:], osz["nNegOut"], nn, fp, draws)
    assert rc == 0 and np.array_equal(got.view(np.uint64), want.view(np.uint64))
    # (3) trees bit-exact against the reference's ranger grown on that matrix (rfThr threads: every tree has its own seed)
    ctx.batch_train(T, mtry, np.array([seed + 0], np.uint32))
    col = np.ascontiguousarray(got.T)
    forest = O.r_forest_train(col, T, mtry, seed, rfThr=min(_threads(), T)) if O.have_ref() else O.o_forest_train(col, T, mtry, seed)
    for t in range(T):
        ok, why = trees_equal(ctx.batch_export_tree(0, t), forest.tree(t))
        assert ok, "tree %d: %s" % (t, why)
    # (4) scores of the partition bit-exact against the reference's predict on the fold's test rows
    ctx.batch_predict_accumulate(nParts)
    g1, g2 = ctx.fold_scores_get()
    test = np.concatenate([fd["testPos"], fd["testNeg"]])
    if O.have_ref():
        tcol = np.ascontiguousarray(np.c_[x[test, :m], np.zeros(len(test))].T)
        pr = O.r_forest_predict(forest, tcol)
    else:
        pr = O.o_forest_predict(forest, x[test])
    inv = 1.0 / nParts
    assert np.array_equal(g1.view(np.uint64), (pr[:, 0] * inv).view(np.uint64))
    assert np.array_equal(g2.view(np.uint64), (pr[:, 1] * inv).view(np.uint64))

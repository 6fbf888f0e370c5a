"""exec.mode "train" / "predict" and ranger's binary forest file: the files our host writes are the
bytes the reference writes (Forest::saveToFile, src/ranger/Forest/Forest.cpp:390-420), the reference reads them back, and
predict mode reproduces the reference's predict mode (src/hyperSMURF.cpp:401-446) bit for bit."""
import numpy as np
import pytest

from helpers import synth


def _col(X, y12):
    return np.ascontiguousarray(np.c_[X, y12].T)


def test_forest_file_reader_and_writer_reproduce_the_reference_file(oracle, tmp_path):
    from parsmurf_b200 import host_api as H
    rng = np.random.default_rng(4)
    for (ntr, m, T, mtry, seed, ties) in [(120, 6, 3, 2, 11, False), (400, 12, 5, 4, 3, True)]:
        X = rng.standard_normal((ntr, m))
        if ties:
            X[:, ::3] = np.round(X[:, ::3])
        y12 = np.r_[np.ones(ntr // 3), 2 * np.ones(ntr - ntr // 3)]
        f = oracle.r_forest_train(_col(X, y12), T, mtry, seed)
        ref_fn = oracle.r_forest_save(f, 7, tmp_path)
        out_fn = tmp_path / ("rewritten_%d.forest" % ntr)
        assert H.forest_file_rewrite(ref_fn, out_fn) == T
        assert open(ref_fn, "rb").read() == open(out_fn, "rb").read()


def test_forest_file_errors(tmp_path):
    from parsmurf_b200 import host_api as H
    from parsmurf_b200._capi import PsmError
    with pytest.raises(PsmError):
        H.forest_file_rewrite(tmp_path / "missing.forest", tmp_path / "o")
    bad = tmp_path / "short.forest"
    bad.write_bytes(b"\x05\x00\x00\x00\x00\x00\x00\x00\x02\x00")
    with pytest.raises(PsmError):
        H.forest_file_rewrite(bad, tmp_path / "o")


@pytest.mark.gpu
def test_train_mode_files_equal_the_reference_and_predict_mode_reads_them(oracle, ctx, tmp_path):
    from parsmurf_b200 import host_api as H
    n, m, npos, nParts, fp, ratio, k, T, mtry, seed = 1500, 8, 60, 3, 2, 2, 3, 4, 3, 5
    X, y = synth(n, m, npos, 9, shift_feats=3)
    x = oracle.make_x(X, y)
    ours = tmp_path / "ours"; ref = tmp_path / "ref"
    ours.mkdir(); ref.mkdir()
    H.train_run(x, y, ours, nParts, fp, ratio, k, T, mtry, seed)
    # the reference's own exec.mode "train", driven like its main() (one randomly generated fold, the rand() draws its
    # std::random_shuffle calls burn before the first SMOTE draw included): the files must be the same bytes
    oracle.r_mode_train(x, y, ref, nParts, fp, ratio, k, T, mtry, seed)
    pos, neg = np.flatnonzero(y == 1).astype(np.uint32), np.flatnonzero(y == 0).astype(np.uint32)
    for p in range(nParts):
        assert open(ref / ("%d.out.forest" % p), "rb").read() == open(ours / ("%d.out.forest" % p), "rb").read(), "partition %d" % p
    # predict mode on our files == the reference's predict mode on our files, accumulated in ascending partition order
    out = H.predict_run(x, y, ours, nParts, T, mtry, seed, tie_mode=0)
    order = np.r_[pos, neg]                                   # copyTestSet: positives first, then negatives
    test_col = np.ascontiguousarray(np.c_[X[order], np.zeros(n)].T)
    c1 = np.zeros(n); c2 = np.zeros(n)
    for p in range(nParts):
        pr = oracle.r_forest_file_predict(ours / ("%d.out.forest" % p), test_col, T, mtry, seed + p)
        c1[order] += pr[:, 0] * (1.0 / nParts)
        c2[order] += pr[:, 1] * (1.0 / nParts)
    assert np.array_equal(out["class1"].view(np.uint64), c1.view(np.uint64))
    assert np.array_equal(out["class2"].view(np.uint64), c2.view(np.uint64))
    want = oracle.o_curves(y, c1)
    assert abs(out["auroc"] - want[0]) < 1e-12 and abs(out["auprc"] - want[1]) < 1e-12
    # a forest directory with a missing partition is an error, not a silent skip
    from parsmurf_b200._capi import PsmError
    with pytest.raises(PsmError):
        H.predict_run(x, y, ours, nParts + 1, T, mtry, seed)

"""Pins the CPU oracle (oracle/psm_oracle.cpp) against the UNMODIFIED reference classes compiled into
oracle/_ref/libparsmurf_ref.so.  Skipped where the reference build did not travel (the golden tests still run)."""
import numpy as np
import pytest

from helpers import CASES, make_case, trees_equal
from oracle import pyoracle as O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("name", list(CASES))
def test_cv_byte_identical(name):
    c = make_case(name)
    x = O.make_x(c["X"], c["y"])
    rc, c1, c2 = O.o_cv(x, c["y"], c["f"], c["nF"], c["nParts"], c["fp"], c["ratio"], c["k"], c["T"], c["mtry"], 7)
    secs, r1, r2 = O.r_cv(x, c["y"], c["f"], c["nF"], c["nParts"], c["fp"], c["ratio"], c["k"], c["T"], c["mtry"], 7)
    assert rc == 0
    assert np.array_equal(c1.view(np.uint64), r1.view(np.uint64))
    assert np.array_equal(c2.view(np.uint64), r2.view(np.uint64))
    assert O.o_curves(c["y"], c1) == O.r_curves(c["y"], r1)


@pytest.mark.parametrize("name", list(CASES))
def test_ttd_sampler_knn_forest_stagewise(name):
    c = make_case(name)
    x = O.make_x(c["X"], c["y"])
    ctx = O.RefCtx(x, c["y"], c["f"], c["nF"])
    rnd = O.GlibcRand(1)
    ttd = O.o_ttd(c["y"], c["f"], c["nF"], rnd)
    for f in range(c["nF"]):
        rf = ctx.fold(f)
        for k_ in rf:
            assert np.array_equal(rf[k_], ttd[f][k_]), (f, k_)
    fd = ttd[0]
    pts = x[fd["trngPos"]]
    kk = min(c["k"], len(pts))
    oi, od = O.o_knn(pts, kk)
    ri, rd = O.r_knn(pts, kk, brute=False)
    bi, bd = O.r_knn(pts, kk, brute=True)
    assert np.array_equal(oi, ri) and np.array_equal(oi, bi)
    assert np.array_equal(od.view(np.uint64), rd.view(np.uint64))
    for p in (0, c["nParts"] - 1):
        want, sz, draws = ctx.sample(0, p, c["nParts"], c["fp"], c["ratio"], c["k"])
        rc, osz = O.o_partition_sizes(len(fd["trngPos"]), len(fd["trngNeg"]), c["nParts"], c["fp"], c["ratio"], p)
        assert rc == 0 and osz["Ntr"] == sz["lineLen"] and osz["nPosOut"] == sz["trngPos"] and osz["nNegOut"] == sz["trngNeg"]
        assert len(draws) == len(fd["trngPos"]) * c["fp"] * (c["m"] + 1)
        rc, got = O.o_assemble(x, fd["trngPos"], fd["trngNeg"][osz["negOffset"]:], osz["nNegOut"], oi, c["fp"], draws)
        assert rc == 0 and np.array_equal(got.view(np.uint64), want.view(np.uint64))
        col = np.ascontiguousarray(want.T)
        of = O.o_forest_train(col, c["T"], c["mtry"], 7 + p)
        rf_ = O.r_forest_train(col, c["T"], c["mtry"], 7 + p)
        for t in range(c["T"]):
            ok, why = trees_equal(of.tree(t), rf_.tree(t))
            assert ok, why
        test_idx = np.concatenate([fd["testPos"], fd["testNeg"]])
        rows = x[test_idx].copy(); rows[:, -1] = 0
        op = O.o_forest_predict(of, rows)
        rp = O.r_forest_predict(rf_, np.ascontiguousarray(rows.T))
        assert np.array_equal(op.view(np.uint64), rp.view(np.uint64))


def test_curves_equal_reference_with_ties():
    rng = np.random.default_rng(5)
    for N, levels in ((1000, None), (200000, 21), (300000, None)):
        s = rng.random(N)
        if levels:
            s = np.round(s * (levels - 1)) / (levels - 1)
        y = (rng.random(N) < 0.02 + 0.3 * s).astype(np.uint32)
        assert O.o_curves(y, s) == O.r_curves(y, s)

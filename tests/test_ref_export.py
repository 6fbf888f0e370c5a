"""The checker's export path (oracle/ref_harness.cpp: ref_run_partitions_ex) is itself checked on the CPU: a MULTI-THREADED run
of the unmodified reference classes -- whose threads interleave on the one global rand() stream -- must be
reproduced bit for bit by the oracle's restatement when every partition is given the draws it recorded.  bench.py's `parity`
block and tests/test_gpu_scale_parity.py rest on this."""
import numpy as np
import pytest

from helpers import make_case


@pytest.mark.parametrize("name,threads", [("fy_26", 4), ("ties_ratio0", 3)])
def test_recorded_draws_reproduce_a_multithreaded_reference_run(oracle, name, threads):
    O = oracle
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    c = make_case(name)
    x = O.make_x(c["X"], c["y"])
    ctx = O.RefCtx(x, c["y"], c["f"], c["nF"])
    secs, c1, c2, draws, preds, fd = ctx.run_partitions_ex(0, 0, c["nParts"], c["nParts"], c["fp"], c["ratio"], c["k"], c["T"], c["mtry"], 1, threads)
    P = len(fd["trngPos"])
    assert draws.shape == (c["nParts"], P * c["fp"] * (c["m"] + 1)) and (draws >= 0).all()
    test = np.concatenate([fd["testPos"], fd["testNeg"]])
    assert preds.shape == (c["nParts"], len(test), 2)
    nn, _ = O.o_knn(x[fd["trngPos"]], min(c["k"], P))
    for p in range(c["nParts"]):
        rc, sz = O.o_partition_sizes(P, len(fd["trngNeg"]), c["nParts"], c["fp"], c["ratio"], p)
        assert rc == 0
        rc, rows = O.o_assemble(x, fd["trngPos"], fd["trngNeg"][sz["negOffset"]:], sz["nNegOut"], nn, c["fp"], draws[p])
        assert rc == 0
        forest = O.o_forest_train(np.ascontiguousarray(rows.T), c["T"], c["mtry"], 1 + p)      # seed + p + fold*nParts, fold 0
        want = O.o_forest_predict(forest, x[test])
        assert np.array_equal(want.view(np.uint64), preds[p].view(np.uint64)), "partition %d" % p
    # and the reference's own accumulation is the ascending sum up to the fp64 addition order of the completion order
    e = np.zeros(len(test))
    for p in range(c["nParts"]):
        e = e + preds[p, :, 0] * (1.0 / c["nParts"])
    assert np.abs(e - c1[test]).max() < 1e-15


def test_simulated_set_is_the_references():
    """cfgEx/simulCV.json's "simulate" block: parsmurf::generateRandomSet == the reference's (src/HyperSMURFUtils.cpp:40-63), value for value"""
    from oracle import pyoracle as O
    from parsmurf_b200 import host_api as H
    if not O.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for n, m, prob, seed in ((1000, 10, 0.01, 1), (777, 3, 0.3, 4242)):
        xr, yr = O.ref_generate_random_set(n, m, prob, seed)
        x, y = H.generate_random_set(n, m, prob, seed)
        assert np.array_equal(y, yr)
        assert np.array_equal(x.view(np.uint64), xr.view(np.uint64))

"""GPU parity tests: every CUDA stage, called through the C ABI, against the CPU oracle on the same seeded inputs.
Bars: neighbour tables equal; training matrices, bootstrap counts, trees and scores BIT-exact; AUROC/AUPRC within
1e-12 (tie_mode 0; the trapezoid terms are the reference's, only the fp64 summation is parallel)."""
import numpy as np
import pytest

from helpers import CASES, make_case, trees_equal

pytestmark = pytest.mark.gpu


def _fold_setup(ctx, oracle, case, fold=0):
    x = oracle.make_x(case["X"], case["y"])
    rnd = oracle.GlibcRand(1)
    ttd = oracle.o_ttd(case["y"], case["f"], case["nF"], rnd)
    fd = ttd[fold]
    ctx.dataset_upload(x)
    ctx.fold_begin(fd["trngPos"], fd["trngNeg"], fd["testPos"], fd["testNeg"])
    return x, fd, rnd


@pytest.mark.parametrize("name", list(CASES))
def test_knn_matches_oracle(ctx, oracle, name):
    case = make_case(name)
    x, fd, _ = _fold_setup(ctx, oracle, case)
    ctx.fold_knn(case["k"])
    idx, dist = ctx.fold_knn_get()
    pts = x[fd["trngPos"]]                      # all m+1 columns, as the reference feeds ANN (label column adds 0)
    oi, od = oracle.o_knn(pts, min(case["k"], len(pts)))
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist.view(np.uint64), od.view(np.uint64))


def test_knn_duplicates_and_missing(ctx, oracle):
    # duplicates are excluded like self matches; missing neighbours come back as -1 / inf
    pts = np.array([[0, 0], [1, 0], [0, 2], [3, 0], [1, 0], [0, 0]], float)
    x = np.concatenate([pts, np.ones((6, 1))], axis=1)
    x = np.concatenate([x, np.array([[9.0, 9.0, 2.0]])])
    ctx.dataset_upload(np.ascontiguousarray(x))
    ctx.fold_begin(np.arange(6), np.array([6]), np.array([0]), np.array([6]))
    ctx.fold_knn(6)
    idx, dist = ctx.fold_knn_get()
    oi, od = oracle.o_knn(x[:6], 6)
    assert np.array_equal(idx, oi)
    assert list(idx[0]) == [1, 4, 2, 3, -1, -1]
    assert np.isinf(dist[0, 4]) and np.isinf(dist[0, 5])


def test_tensor_core_knn_gives_the_exact_table(ctx, oracle, monkeypatch):
    """knn_tc.cu: candidates from the tcgen05 3xTF32 distance GEMM, exact fp64 rerank, completeness proof -- the neighbour
    table (indices AND fp64 distances) must be the exact kernel's bit for bit, on small folds (every point a candidate),
    on folds with several 128-row tiles and a ragged feature count, with duplicates, and with near-degenerate data"""
    rng = np.random.default_rng(17)
    shapes = [(90, 26, 5), (700, 26, 5), (1500, 200, 10), (1030, 77, 12), (2100, 333, 20), (300, 1000, 10)]
    paths = []
    for P, m, k in shapes:
        n = P + 50
        X = rng.standard_normal((n, m))
        if P == 1030:
            X[:40] = X[40:80]                     # exact duplicates among the positives (zero distances are excluded)
            X[100:140] = X[140:180] + 1e-9        # near duplicates: distances far below the fp32 error of the candidate search
        y = np.zeros(n, np.uint32); y[:P] = 1
        x = oracle.make_x(X, y)
        ctx.dataset_upload(x)
        pos = np.arange(P, dtype=np.uint32); neg = np.arange(P, n, dtype=np.uint32)
        ctx.fold_begin(pos, neg, pos[:1], neg[:1])
        monkeypatch.setenv("PSM_KNN_TENSOR", "0")
        ctx.fold_knn(k)
        assert ctx.fold_knn_path()[0] == 0
        i0, d0 = ctx.fold_knn_get()
        monkeypatch.setenv("PSM_KNN_TENSOR", "1")
        ctx.fold_knn(k)
        path, bad = ctx.fold_knn_path()
        i1, d1 = ctx.fold_knn_get()
        assert path in (1, 2)
        assert np.array_equal(i0, i1), (P, m, k, path, bad)
        assert np.array_equal(d0.view(np.uint64), d1.view(np.uint64)), (P, m, k, path, bad)
        paths.append((P, m, path, bad))
    # well-separated random data must be proven complete by the tensor path itself (no fallback)
    assert all(path == 1 for P, m, path, bad in paths if P != 1030), paths


@pytest.mark.parametrize("name", list(CASES))
def test_assemble_bit_exact(ctx, oracle, name):
    case = make_case(name)
    x, fd, rnd = _fold_setup(ctx, oracle, case)
    ctx.fold_knn(case["k"])
    nn, _ = ctx.fold_knn_get()
    P = len(fd["trngPos"])
    per = P * case["fp"] * (case["m"] + 1)
    draws = rnd.fill(per * case["nParts"])
    ctx.batch_begin(case["nParts"], case["fp"], case["ratio"], 0, case["nParts"])
    ctx.batch_assemble(draws)
    for p in range(case["nParts"]):
        rc, sz = oracle.o_partition_sizes(P, len(fd["trngNeg"]), case["nParts"], case["fp"], case["ratio"], p)
        assert rc == 0
        rc, want = oracle.o_assemble(x, fd["trngPos"], fd["trngNeg"][sz["negOffset"]:], sz["nNegOut"], nn, case["fp"], draws[p * per:(p + 1) * per])
        assert rc == 0
        got, gsz = ctx.batch_get_training(p)
        assert gsz["Ntr"] == sz["Ntr"] and gsz["trngPos"] == sz["nPosOut"] and gsz["trngNeg"] == sz["nNegOut"]
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), "partition %d differs" % p


def test_device_rand_replay_equals_host_draws(ctx, oracle):
    """SMOTE draws replayed on the GPU from glibc's generator state == the same draws generated on the host"""
    import parsmurf_b200._capi as capi
    case = make_case("fy_26")
    x, fd, rnd = _fold_setup(ctx, oracle, case)
    ctx.fold_knn(case["k"])
    P = len(fd["trngPos"])
    per = P * case["fp"] * (case["m"] + 1)
    # stream position after the folds' training-negative shuffles (one rand() per element but the first)
    consumed = sum(len(t["trngNeg"]) - 1 for t in oracle.o_ttd(case["y"], case["f"], case["nF"], None))
    w = capi.glibc_jump(capi.glibc_initial_window(1), consumed)
    draws = rnd.fill(per * case["nParts"])
    ctx.batch_begin(case["nParts"], case["fp"], case["ratio"], 0, case["nParts"])
    ctx.batch_assemble(draws)
    want = [ctx.batch_get_training(p)[0] for p in range(case["nParts"])]
    ctx.batch_begin(case["nParts"], case["fp"], case["ratio"], 0, case["nParts"])
    ctx.batch_assemble_glibc(w)
    for p in range(case["nParts"]):
        got = ctx.batch_get_training(p)[0]
        assert np.array_equal(got.view(np.uint64), want[p].view(np.uint64))


def test_device_rand_replay_far_into_the_stream(ctx):
    """psm_glibc_fill_device == the host generator at every digit of the radix-256 jump tables: the first chunk, chunk
    indices with a second digit (>= 256 chunks of 248 draws) and with a third one (>= 65536 chunks = 16.25 M draws)"""
    import parsmurf_b200._capi as capi
    from parsmurf_b200 import host_api as H
    w0 = capi.glibc_initial_window(1)
    head = H.rand_fill(1, 70000)
    got = ctx.glibc_fill_device(w0, 70000)
    assert np.array_equal(got, head)                                  # chunks 0..282: digits 0 and 1
    count = 248 * 65536 * 2 + 1000                                     # third digit 0, 1 and 2
    for skip in (248 * 65536 - 500, 248 * 65536 * 2 - 300, count - 1000):
        got = ctx.glibc_fill_device(w0, count, skip, 1000)
        wj = capi.glibc_jump(w0, skip)                                # host: O(log n) jump, then the plain recurrence
        r = list(wj.astype(np.uint64))
        want = []
        for _ in range(1000):
            v = (int(r[-3]) + int(r[-31])) & 0xFFFFFFFF
            r.append(v); want.append(v >> 1)
        assert np.array_equal(got, np.array(want, np.int32)), skip
    # a stream entered in the middle (the window of a later position) continues identically
    w1 = capi.glibc_jump(w0, 12345)
    assert np.array_equal(ctx.glibc_fill_device(w1, 3000), head[12345:15345])


def test_device_folds_equal_the_host_test_train_divider(ctx, oracle):
    """psm_fold_begin_device: the four index arrays of every fold, built from the per-class fold lists on the device, with the
    training negatives shuffled by the device replay of std::random_shuffle on glibc rand() (csrc/shuffle.cu) == the host
    TestTrainDivider == the reference's arrays (tests/test_host_logic.py::test_ttd_matches_oracle pins the host ones)"""
    from parsmurf_b200 import _capi as psm
    from parsmurf_b200 import host_api as H
    rng = np.random.default_rng(5)
    for n, npos, nF in ((3000, 100, 5), (20011, 37, 3), (257, 9, 4), (70000, 1200, 10)):
        y = np.zeros(n, np.uint32); y[rng.choice(n, npos, replace=False)] = 1
        f = rng.integers(0, nF, n).astype(np.uint32)
        x = oracle.make_x(rng.standard_normal((n, 3)), y)
        ctx.dataset_upload(x)
        pos = [np.flatnonzero((f == k) & (y == 1)).astype(np.uint32) for k in range(nF)]
        neg = [np.flatnonzero((f == k) & (y == 0)).astype(np.uint32) for k in range(nF)]
        off = lambda parts: np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
        ctx.folds_upload(np.concatenate(pos), off(pos), np.concatenate(neg), off(neg))
        w = psm.glibc_initial_window(1)
        for k in range(nF):
            want = H.ttd_fold(y, f, nF, k)
            ctx.fold_begin_device(k, w)
            got = ctx.fold_indices_get()
            for key in ("trngPos", "trngNeg", "testPos", "testNeg"):
                assert np.array_equal(got[key], want[key]), (n, k, key)
            assert not np.array_equal(got["trngNeg"], np.sort(got["trngNeg"]))          # really shuffled
            w = psm.glibc_jump(w, max(len(want["trngNeg"]) - 1, 0))                      # fold k's shuffle consumed trngNeg - 1 draws
        ctx.fold_begin_device(0, None)                                                   # no window: training negatives in list order
        got = ctx.fold_indices_get()
        assert np.array_equal(got["trngNeg"], np.concatenate(neg[1:]))


def test_host_and_device_folds_give_the_same_cv(oracle, monkeypatch):
    from parsmurf_b200 import host_api as H
    case = make_case("fy_26")
    x = oracle.make_x(case["X"], case["y"])
    args = (x, case["y"], case["f"], case["nF"], case["nParts"], case["fp"], case["ratio"], case["k"], case["T"], case["mtry"], 7)
    a = H.cv_run(*args, tie_mode=0)
    monkeypatch.setenv("PSM_HOST_FOLDS", "1")
    c = H.cv_run(*args, tie_mode=0)
    assert np.array_equal(a["class1"].view(np.uint64), c["class1"].view(np.uint64)) and np.array_equal(a["fold_metrics"], c["fold_metrics"])


def test_cv_through_host_classes_bit_exact(oracle):
    """the whole CV through the C++ host mirror (hyperSMURF::smurfIt) == the oracle == the reference at ensThrd 1"""
    from parsmurf_b200 import host_api as H
    for name in ("fy_26", "reject_300", "ties_ratio0"):
        case = make_case(name)
        x = oracle.make_x(case["X"], case["y"])
        rc, o1, o2 = oracle.o_cv(x, case["y"], case["f"], case["nF"], case["nParts"], case["fp"], case["ratio"], case["k"], case["T"], case["mtry"], 7)
        out = H.cv_run(x, case["y"], case["f"], case["nF"], case["nParts"], case["fp"], case["ratio"], case["k"], case["T"], case["mtry"], 7, tie_mode=0)
        assert np.array_equal(out["class1"].view(np.uint64), o1.view(np.uint64)), name
        assert np.array_equal(out["class2"].view(np.uint64), o2.view(np.uint64)), name
        want = oracle.o_curves(case["y"], o1)
        assert abs(out["auroc"] - want[0]) < 1e-12 and abs(out["auprc"] - want[1]) < 1e-12
        # tiny batches (one partition per launch group) give the same bits
        out2 = H.cv_run(x, case["y"], case["f"], case["nF"], case["nParts"], case["fp"], case["ratio"], case["k"], case["T"], case["mtry"], 7, tie_mode=0, batch_budget=1)
        assert np.array_equal(out2["class1"].view(np.uint64), o1.view(np.uint64)), name


def test_fold_lanes_give_identical_bits(oracle):
    """folds in flight concurrently (one context + stream + host thread per lane) == folds one after the other, bit for bit:
    scores, per-fold and global curves; also for a rank's share of a sharded run (the gated reduce path)"""
    from parsmurf_b200 import host_api as H
    for name in ("fy_26", "ties_ratio0"):
        case = make_case(name)
        x = oracle.make_x(case["X"], case["y"])
        args = (x, case["y"], case["f"], case["nF"], case["nParts"], case["fp"], case["ratio"], case["k"], case["T"], case["mtry"], 7)
        one = H.cv_run(*args, tie_mode=0, fold_lanes=1)
        for lanes in (2, 3, 8):
            out = H.cv_run(*args, tie_mode=0, fold_lanes=lanes)
            assert np.array_equal(out["class1"].view(np.uint64), one["class1"].view(np.uint64)), (name, lanes)
            assert np.array_equal(out["class2"].view(np.uint64), one["class2"].view(np.uint64)), (name, lanes)
            assert np.array_equal(out["fold_metrics"], one["fold_metrics"]) and out["auroc"] == one["auroc"] and out["auprc"] == one["auprc"]
            assert out["stats"] == one["stats"] and out["kernel_launches"]["split_search"] == one["kernel_launches"]["split_search"]
        calls = []
        a1 = H.cv_run(*args, tie_mode=0, rank=1, world=2, reduce=lambda ptr, count: None, eval_fold_curves=False, fold_lanes=1, shard="fold")
        a3 = H.cv_run(*args, tie_mode=0, rank=1, world=2, reduce=lambda ptr, count: calls.append(count), eval_fold_curves=False, fold_lanes=3, shard="fold")
        assert np.array_equal(a1["class1"].view(np.uint64), a3["class1"].view(np.uint64)), name
        assert len(calls) == case["nF"]            # one reduce per fold, issued from the calling thread in fold order


def test_size_classes_do_not_change_the_trees(oracle, monkeypatch):
    """split search over the size-class lists (sub-warp groups for short segments; one merged launch per level, largest class
    first, at every CTA size) vs one launch per class vs one generic class"""
    from parsmurf_b200 import host_api as H
    case = make_case("fy_26")
    x = oracle.make_x(case["X"], case["y"])
    args = (x, case["y"], case["f"], case["nF"], case["nParts"], case["fp"], case["ratio"], case["k"], case["T"], case["mtry"], 7)
    a = H.cv_run(*args, tie_mode=0)
    assert a["kernel_launches"]["split_search"] == a["stats"]["levels"]        # merged: one launch per level
    for block in ("32", "128"):
        monkeypatch.setenv("PSM_SS_BLOCK", block)
        c = H.cv_run(*args, tie_mode=0)
        assert np.array_equal(a["class1"].view(np.uint64), c["class1"].view(np.uint64)) and a["stats"] == c["stats"], block
    monkeypatch.delenv("PSM_SS_BLOCK")
    monkeypatch.setenv("PSM_SS_SEPARATE", "1")
    sep = H.cv_run(*args, tie_mode=0)
    assert np.array_equal(a["class1"].view(np.uint64), sep["class1"].view(np.uint64)) and a["stats"] == sep["stats"]
    assert sep["kernel_launches"]["split_search"] > sep["stats"]["levels"]
    monkeypatch.setenv("PSM_NO_SIZE_CLASSES", "1")
    b = H.cv_run(*args, tie_mode=0)
    assert np.array_equal(a["class1"].view(np.uint64), b["class1"].view(np.uint64))
    assert a["stats"] == b["stats"]
    assert b["kernel_launches"]["split_search"] == b["stats"]["levels"]
    monkeypatch.delenv("PSM_SS_SEPARATE")
    b2 = H.cv_run(*args, tie_mode=0)                                           # one generic class through the merged kernel
    assert np.array_equal(a["class1"].view(np.uint64), b2["class1"].view(np.uint64)) and a["stats"] == b2["stats"]


def test_two_rank_shares_on_one_gpu(oracle):
    """partition sharding (rank r of 2 handles its contiguous chunk and skips the other rank's SMOTE draws): the two
    partial results add up to the single-rank CV; emulated as two sequential runs on one GPU, no collective needed"""
    from parsmurf_b200 import host_api as H
    case = make_case("fy_26")
    x = oracle.make_x(case["X"], case["y"])
    args = (x, case["y"], case["f"], case["nF"], case["nParts"], case["fp"], case["ratio"], case["k"], case["T"], case["mtry"], 7)
    full = H.cv_run(*args, tie_mode=0)
    noop = lambda ptr, count: None
    a = H.cv_run(*args, tie_mode=0, rank=0, world=2, reduce=noop, eval_fold_curves=False, shard="fold")
    b = H.cv_run(*args, tie_mode=0, rank=1, world=2, reduce=noop, eval_fold_curves=False, shard="fold")
    assert np.allclose(a["class1"] + b["class1"], full["class1"], rtol=0, atol=1e-15)
    assert np.allclose(a["class2"] + b["class2"], full["class2"], rtol=0, atol=1e-15)
    assert np.abs(a["class1"]).sum() > 0 and np.abs(b["class1"]).sum() > 0
    rc, o1, o2 = oracle.o_cv_parts(*args, 0, 2)
    assert np.array_equal(a["class1"].view(np.uint64), o1.view(np.uint64))


def test_unit_sharded_ranks_reproduce_the_single_rank_cv(oracle):
    """unit sharding (contiguous ranges of the nFolds x nParts units; ONE sum of the score arrays at the end; per-fold curves
    from the summed arrays, fold f on rank f % world), three ranks emulated one after the other on one GPU: the reduce
    callback of a rank adds the other ranks' contributions, captured from earlier passes, into the device buffer"""
    import torch
    from parsmurf_b200 import host_api as H

    class _Arr:
        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}

    # (world, nParts, nTrees): the second shape gives a rank more than 32 units of one fold, so its share is cut into pieces
    for W, nParts, T in ((3, 4, 10), (2, 80, 2)):
      case = make_case("fy_26")
      x = oracle.make_x(case["X"], case["y"])
      args = (x, case["y"], case["f"], case["nF"], nParts, case["fp"], case["ratio"], case["k"], T, case["mtry"], 7)
      n, nF = len(case["y"]), case["nF"]
      full = H.cv_run(*args, tie_mode=0, fold_lanes=1)
      for lanes in (1, 3):
          def run(rank, add_scores=None, add_metrics=None):
              got = []
              def reduce(ptr, count):
                  t = torch.as_tensor(_Arr(ptr, count), device="cuda")
                  got.append(t.cpu().numpy().copy())
                  extra = add_scores if count == 2 * n else add_metrics
                  assert count in (2 * n, 2 * nF + 2)                         # the scores; the per-fold metrics + the global pair
                  if extra is not None:
                      t += torch.from_numpy(extra).cuda()
                  torch.cuda.synchronize()
              out = H.cv_run(*args, tie_mode=0, rank=rank, world=W, reduce=reduce, shard="unit", fold_lanes=lanes)
              return out, got
          partial = [run(r)[1][0] for r in range(W)]                          # pass 1: every rank's partial class1Prob/class2Prob
          assert all(np.abs(p).sum() > 0 for p in partial)
          if lanes == 1 and nParts == 4:                                      # a whole-fold share per lane: the rank's partial sums are the oracle's, bit for bit
              for r in range(W):
                  rng_ = [H.part_range(True, nF, nParts, W, r, f) for f in range(nF)]
                  rc, o1, o2 = oracle.o_cv_ranges(*args, [a for a, b in rng_], [b for a, b in rng_])
                  assert np.array_equal(partial[r][:n].view(np.uint64), o1.view(np.uint64)) and np.array_equal(partial[r][n:].view(np.uint64), o2.view(np.uint64)), r
          tot = sum(partial)
          assert np.allclose(tot[:n], full["class1"], rtol=0, atol=1e-13) and np.allclose(tot[n:], full["class2"], rtol=0, atol=1e-13)   # fp64 addition order across ranks
          others = lambda r, xs: sum(xs[q] for q in range(W) if q != r)
          metrics = [run(r, add_scores=others(r, partial))[1][1] for r in range(W)]   # pass 2: own folds' metrics from the summed scores
          for r in range(W):
              out, _ = run(r, add_scores=others(r, partial), add_metrics=others(r, metrics))   # pass 3: what rank r returns in a real run
              assert np.allclose(out["class1"], full["class1"], rtol=0, atol=1e-13) and np.allclose(out["class2"], full["class2"], rtol=0, atol=1e-13)
              assert np.allclose(out["fold_metrics"], full["fold_metrics"], rtol=0, atol=1e-4), (r, out["fold_metrics"], full["fold_metrics"])   # 1-ulp score differences reorder exact ties (no tie grouping)
              assert abs(out["auroc"] - full["auroc"]) < 1e-4 and abs(out["auprc"] - full["auprc"]) < 1e-4


def test_cuda_path_against_committed_golden_fixtures(oracle):
    """CUDA path vs the reference's own outputs stored in tests/golden (no oracle in between)"""
    import os
    from parsmurf_b200 import host_api as H
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "cfg1_example_transposed.npz"))
    x = oracle.make_x(g["X"], g["y"])
    out = H.cv_run(x, g["y"], g["folds"], int(g["nFolds"]), int(g["nParts"]), int(g["fp"]), int(g["ratio"]), int(g["k"]), int(g["nTrees"]), int(g["mtry"]), int(g["seed"]), tie_mode=0)
    assert np.array_equal(out["class1"].view(np.uint64), g["class1"].view(np.uint64))
    assert np.array_equal(out["class2"].view(np.uint64), g["class2"].view(np.uint64))
    assert abs(out["auroc"] - float(g["auroc"])) < 1e-12 and abs(out["auprc"] - float(g["auprc"])) < 1e-12
    assert np.allclose(out["fold_metrics"], g["fold_metrics"], rtol=0, atol=1e-12)
    for name in CASES:
        c = make_case(name)
        g = np.load(os.path.join(gold, "case_%s.npz" % name))
        x = oracle.make_x(c["X"], c["y"])
        out = H.cv_run(x, c["y"], c["f"], c["nF"], c["nParts"], c["fp"], c["ratio"], c["k"], c["T"], c["mtry"], 7, tie_mode=0)
        assert np.array_equal(out["class1"].view(np.uint64), g["class1"].view(np.uint64)), name
        assert abs(out["auroc"] - float(g["auroc"])) < 1e-12 and abs(out["auprc"] - float(g["auprc"])) < 1e-12
        assert np.allclose(out["fold_metrics"], g["fold_metrics"], rtol=0, atol=1e-12)


def test_grid_search_matches_reference_golden():
    """optimizer "grid" (inner CV per grid point, arg-max AUPRC, outer fold with the winner) == the reference's own run"""
    import os
    from parsmurf_b200 import host_api as H
    from oracle import pyoracle as O
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid_fy_26.npz"))
    c = make_case("fy_26")
    x = O.make_x(c["X"], c["y"])
    out = H.cv_grid_run(x, c["y"], c["f"], c["nF"], [tuple(int(v) for v in r) for r in g["grid"]], 7, tie_mode=0)
    assert np.allclose(out["grid_metrics"], g["grid_metrics"], rtol=0, atol=1e-12)
    assert np.array_equal(out["class1"].view(np.uint64), g["class1"].view(np.uint64))
    assert np.array_equal(out["class2"].view(np.uint64), g["class2"].view(np.uint64))
    assert abs(out["auroc"] - float(g["auroc"])) < 1e-12 and abs(out["auprc"] - float(g["auprc"])) < 1e-12


def _threaded_ranks(W, fn):
    """W ranks of one multi-rank run as W host threads on one GPU: reduce(ptr, count) is a real all-reduce between the threads
    (every rank's buffer is read, summed in rank order, written back), so every collective the run issues is exercised in order"""
    import threading
    import torch

    class _Arr:
        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    bar = threading.Barrier(W)
    bufs, outs, errs, counts = [None] * W, [None] * W, [], [[] for _ in range(W)]

    def reducer(r):
        def reduce(ptr, count):
            t = torch.as_tensor(_Arr(ptr, count), device="cuda")
            torch.cuda.synchronize()
            bufs[r] = t.clone()
            counts[r].append(count)
            bar.wait(timeout=120)
            tot = bufs[0].clone()
            for q in range(1, W):
                tot += bufs[q]
            torch.cuda.synchronize()
            bar.wait(timeout=120)
            t.copy_(tot)
            torch.cuda.synchronize()
        return reduce

    def work(r):
        try:
            outs[r] = fn(r, reducer(r))
        except Exception as e:   # a dead rank must not leave the others waiting
            errs.append((r, e))
            bar.abort()
    th = [threading.Thread(target=work, args=(r,)) for r in range(W)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    assert all(c == counts[0] for c in counts), "the ranks issued different collectives"
    return outs, counts[0]


def test_grid_search_sharded_by_grid_point_equals_the_single_rank_run():
    """multi-rank grid search: every rank runs whole inner CVs (its share of the grid points, SMOTE stream entered at each point's
    offset) and the [G][2] metric table is summed once per outer fold -- the table, the winners and therefore the outer folds
    must be those of one rank; the outer folds themselves are chunked over the ranks (sum order differs: 1e-13)"""
    import os
    from parsmurf_b200 import host_api as H
    from oracle import pyoracle as O
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid_fy_26.npz"))
    c = make_case("fy_26")
    x = O.make_x(c["X"], c["y"])
    grid = [tuple(int(v) for v in r) for r in g["grid"]]
    n, G, nF = len(c["y"]), len(grid), c["nF"]
    for W in (2, 3):
        devs = [H.HostDevice(0) for _ in range(W)]
        try:
            outs, counts = _threaded_ranks(W, lambda r, red: devs[r].cv_grid_run(x, c["y"], c["f"], nF, grid, 7, tie_mode=0, rank=r, world=W, reduce=red))
        finally:
            for d in devs:
                d.close()
        assert counts.count(2 * G) == nF, counts                                # one metric-table reduce per outer fold, none per inner fold
        assert len(counts) == 2 * nF, counts                                   # + one score reduce per outer fold
        for out in outs:
            assert np.allclose(out["grid_metrics"], g["grid_metrics"], rtol=0, atol=1e-12)
            assert np.allclose(out["class1"], g["class1"], rtol=0, atol=1e-13) and np.allclose(out["class2"], g["class2"], rtol=0, atol=1e-13)
            assert abs(out["auroc"] - float(g["auroc"])) < 1e-4 and abs(out["auprc"] - float(g["auprc"])) < 1e-4


def _train_case(ctx, oracle, case, fold=0, seed=7):
    x, fd, rnd = _fold_setup(ctx, oracle, case, fold)
    ctx.fold_knn(case["k"])
    P = len(fd["trngPos"])
    per = P * case["fp"] * (case["m"] + 1)
    draws = rnd.fill(per * case["nParts"])
    ctx.batch_begin(case["nParts"], case["fp"], case["ratio"], 0, case["nParts"])
    ctx.batch_assemble(draws)
    seeds = np.array([seed + p + fold * case["nParts"] for p in range(case["nParts"])], np.uint32)
    ctx.batch_train(case["T"], case["mtry"], seeds)
    return x, fd, seeds


@pytest.mark.parametrize("name", list(CASES))
def test_bootstrap_replay_matches_std_mt19937_64(ctx, oracle, name):
    case = make_case(name)
    x, fd, seeds = _train_case(ctx, oracle, case)
    for p in (0, case["nParts"] - 1):
        tr, sz = ctx.batch_get_training(p)
        for t in (0, case["T"] - 1):
            tree_seed = (int(seeds[p]) * (t + 1)) & 0xFFFFFFFF
            boot, _ = oracle.o_tree_draws(tree_seed, sz["Ntr"], case["m"], case["mtry"], 1)
            want = np.bincount(boot.astype(np.int64), minlength=sz["Ntr"]).astype(np.uint32)
            got = ctx.batch_get_inbag(p, t, sz["Ntr"])
            assert np.array_equal(got, want)


@pytest.mark.parametrize("name", list(CASES))
def test_trees_bit_exact(ctx, oracle, name):
    case = make_case(name)
    x, fd, seeds = _train_case(ctx, oracle, case)
    total_nodes = 0
    for p in range(case["nParts"]):
        tr, sz = ctx.batch_get_training(p)
        of = oracle.o_forest_train(np.ascontiguousarray(tr.T), case["T"], case["mtry"], int(seeds[p]))
        for t in range(case["T"]):
            ok, why = trees_equal(ctx.batch_export_tree(p, t), of.tree(t))
            assert ok, "partition %d tree %d: %s" % (p, t, why)
            total_nodes += len(of.tree(t)["var"])
    assert total_nodes > case["nParts"] * case["T"]


def test_trees_bit_exact_wide_entries(ctx, oracle, monkeypatch):
    """the general paths used when Ntr > 16384 (8-byte list entries, global-memory radix sort) grow the same trees"""
    monkeypatch.setenv("PSM_FORCE_ENTRY64", "1")
    monkeypatch.setenv("PSM_FORCE_RADIX_SORT", "1")
    case = make_case("ties_ratio0")
    x, fd, seeds = _train_case(ctx, oracle, case)
    for p in range(case["nParts"]):
        tr, sz = ctx.batch_get_training(p)
        of = oracle.o_forest_train(np.ascontiguousarray(tr.T), case["T"], case["mtry"], int(seeds[p]))
        for t in range(case["T"]):
            ok, why = trees_equal(ctx.batch_export_tree(p, t), of.tree(t))
            assert ok, "partition %d tree %d: %s" % (p, t, why)


# feature-sort edge sizes: the split bitonic network sorts run A = largest power of two <= Ntr and run B = the rest
# (padded to its own power of two) and merges them; cover Ntr a power of two (no run B), run B of one element, run B padded
# to the size of run A, and a mid-size segment -- each with and without heavy value ties, plus the padded single network.
SORT_SIZES = [
    # (n, npos, nFolds, nParts, fp, ratio, expected Ntr)
    (400, 20, 5, 2, 1, 1, 64),      # Ntr = 2^6
    (400, 25, 5, 2, 1, 1, 80),      # A = 64, run B = 16
    (400, 35, 5, 2, 1, 1, 112),     # A = 64, run B = 48 -> padded to 64 == A
    (330, 20, 5, 4, 0, 0, None),    # ratio 0: Ntr = P + chunk (odd sizes, last partition shorter)
    (12000, 310, 5, 3, 2, 3, 2976),  # A = 2048, run B = 928 -> 1024
]


@pytest.mark.parametrize("spec", SORT_SIZES)
@pytest.mark.parametrize("disc", [False, True])
def test_trees_bit_exact_over_sort_sizes(ctx, oracle, spec, disc):
    from helpers import synth, folds_strat
    n, npos, nF, nParts, fp, ratio, want_ntr = spec
    X, y = synth(n, 9, npos, 5, shift_feats=3, discretize=disc)
    case = dict(X=X, y=y, f=folds_strat(y, nF, 3), n=n, m=9, nF=nF, nParts=nParts, fp=fp, ratio=ratio, k=3, T=3, mtry=3)
    x, fd, seeds = _train_case(ctx, oracle, case)
    sizes = set()
    for p in range(nParts):
        tr, sz = ctx.batch_get_training(p)
        sizes.add(sz["Ntr"])
        of = oracle.o_forest_train(np.ascontiguousarray(tr.T), case["T"], case["mtry"], int(seeds[p]))
        for t in range(case["T"]):
            ok, why = trees_equal(ctx.batch_export_tree(p, t), of.tree(t))
            assert ok, "Ntr %d partition %d tree %d: %s" % (sz["Ntr"], p, t, why)
    if want_ntr is not None:
        assert sizes == {want_ntr}


def test_trees_bit_exact_padded_sort_network(ctx, oracle, monkeypatch):
    """PSM_SORT_POW2 keeps the single padded bitonic network selectable; it must grow the same trees"""
    monkeypatch.setenv("PSM_SORT_POW2", "1")
    case = make_case("fy_26")
    x, fd, seeds = _train_case(ctx, oracle, case)
    for p in range(case["nParts"]):
        tr, sz = ctx.batch_get_training(p)
        of = oracle.o_forest_train(np.ascontiguousarray(tr.T), case["T"], case["mtry"], int(seeds[p]))
        for t in range(case["T"]):
            ok, why = trees_equal(ctx.batch_export_tree(p, t), of.tree(t))
            assert ok, "partition %d tree %d: %s" % (p, t, why)


def test_trees_with_injected_draws(ctx, oracle):
    """reference-exported draws injected instead of the device replay: same trees"""
    case = make_case("fy_26")
    x, fd, seeds = _train_case(ctx, oracle, case)
    cap = 4096
    boots, cands = [], []
    for p in range(case["nParts"]):
        _, sz = ctx.batch_get_training(p)
        for t in range(case["T"]):
            tree_seed = (int(seeds[p]) * (t + 1)) & 0xFFFFFFFF
            b, c = oracle.o_tree_draws(tree_seed, sz["Ntr"], case["m"], case["mtry"], cap)
            boots.append(b); cands.append(c)
    ref = [[ctx.batch_export_tree(p, t) for t in range(case["T"])] for p in range(case["nParts"])]
    ctx.batch_train(case["T"], case["mtry"], None, inj_bootstrap=np.concatenate(boots), inj_cand=np.stack(cands))
    for p in range(case["nParts"]):
        for t in range(case["T"]):
            ok, why = trees_equal(ctx.batch_export_tree(p, t), ref[p][t])
            assert ok, why


@pytest.mark.parametrize("name", list(CASES))
def test_fold_scores_bit_exact_and_curves(ctx, oracle, name):
    case = make_case(name)
    x = oracle.make_x(case["X"], case["y"])
    rc, o1, o2 = oracle.o_cv(x, case["y"], case["f"], case["nF"], case["nParts"], case["fp"], case["ratio"], case["k"], case["T"], case["mtry"], 7)
    assert rc == 0
    rnd = oracle.GlibcRand(1)
    ttd = oracle.o_ttd(case["y"], case["f"], case["nF"], rnd)
    ctx.dataset_upload(x)
    c1 = np.zeros(case["n"]); c2 = np.zeros(case["n"])
    for fold in range(case["nF"]):
        fd = ttd[fold]
        ctx.fold_begin(fd["trngPos"], fd["trngNeg"], fd["testPos"], fd["testNeg"])
        ctx.fold_knn(case["k"])
        per = len(fd["trngPos"]) * case["fp"] * (case["m"] + 1)
        draws = rnd.fill(per * case["nParts"])
        ctx.fold_run_partitions(case["nParts"], case["fp"], case["ratio"], case["T"], case["mtry"], 7, fold, 0, case["nParts"], draws)
        ctx.fold_scatter_host(c1, c2)
    assert np.array_equal(c1.view(np.uint64), o1.view(np.uint64))
    assert np.array_equal(c2.view(np.uint64), o2.view(np.uint64))
    want = oracle.o_curves(case["y"], o1)
    got = ctx.curves(case["y"], c1, tie_mode=0)
    assert abs(got[0] - want[0]) < 1e-12 and abs(got[1] - want[1]) < 1e-12
    got1 = ctx.curves(case["y"], c1, tie_mode=1)
    assert abs(got1[0] - want[0]) < 5e-3 and abs(got1[1] - want[1]) < 5e-3   # stable tie order may differ


def test_curves_auto_tie_mode(ctx, oracle):
    """tie_mode 2: exact when no run of tied scores holds both labels; otherwise bracketed by the positives-first / negatives-first
    tie orders -- the reference's value (oracle: the identical std::sort call) lies inside the bracket, the midpoint is returned
    when the bracket is narrower than 1e-6, and the host std::sort decides when it is not"""
    rng = np.random.default_rng(11)
    n = 200_000
    y = (rng.random(n) < 0.01).astype(np.uint32)
    # (a) continuous scores: no ties at all
    s = rng.random(n) + 0.3 * y
    want = oracle.o_curves(y, s)
    got = ctx.curves(y, s, tie_mode=2)
    mixed, host, spread = ctx.curves_info()
    assert mixed == 0 and host == 0 and spread == (0.0, 0.0)
    assert abs(got[0] - want[0]) < 1e-12 and abs(got[1] - want[1]) < 1e-12
    # (b) a huge run of ties that is single-label (most negatives score exactly 0), ties elsewhere only among equal labels
    s = np.where(y == 1, 0.5 + np.round(rng.random(n), 2), np.where(rng.random(n) < 0.9, 0.0, np.round(0.4 * rng.random(n), 3)))
    want = oracle.o_curves(y, s)
    got = ctx.curves(y, s, tie_mode=2)
    mixed, host, spread = ctx.curves_info()
    assert mixed == 0 and host == 0
    assert abs(got[0] - want[0]) < 1e-12 and abs(got[1] - want[1]) < 1e-12
    # (c) a few positives inside the big run of zeros: the tie order moves the curves by less than the tolerance -> midpoint
    s2 = s.copy()
    pos = np.flatnonzero(y == 1)
    s2[pos[:3]] = 0.0
    want = oracle.o_curves(y, s2)
    hi = ctx.curves(y, s2, tie_mode=2)
    mixed, host, spread = ctx.curves_info()
    assert mixed > 0
    if host == 0:
        assert 0 <= spread[0] <= 1e-6 and 0 <= spread[1] <= 1e-6
        assert abs(hi[0] - want[0]) <= spread[0] / 2 + 1e-8 and abs(hi[1] - want[1]) <= spread[1] / 2 + 1e-12   # AUROC terms are float: 1e-8 of slack
    else:
        assert abs(hi[0] - want[0]) < 1e-12 and abs(hi[1] - want[1]) < 1e-12
    # (d) heavy mixed ties (21 score levels): the bracket is wide, the host sort must decide, result == tie_mode 0 == oracle
    s3 = np.round(rng.random(n) * 20) / 20 + 0.1 * y
    s3 = np.round(s3 * 20) / 20
    want = oracle.o_curves(y, s3)
    got = ctx.curves(y, s3, tie_mode=2)
    mixed, host, spread = ctx.curves_info()
    assert mixed > 0 and host == 1 and (spread[0] > 1e-6 or spread[1] > 1e-6)
    assert abs(got[0] - want[0]) < 1e-12 and abs(got[1] - want[1]) < 1e-12
    # the bracket really brackets: tie_mode 1 (ties in index order) and the reference's order both lie inside [lo, hi]
    t1 = ctx.curves(y, s3, tie_mode=1)
    assert want[0] - spread[0] - 1e-7 <= t1[0] <= want[0] + spread[0] + 1e-7


def test_predict_paths_agree_bitwise(oracle, monkeypatch):
    """compact predict (fp32 tile + 8-byte nodes + exact fp64 tie-break) == 16-byte-node TMA path == oracle, incl. data with
    values that collide in fp32 (features rounded to a coarse grid so float(v) == float(threshold) happens constantly)"""
    from parsmurf_b200 import host_api as H
    case = make_case("ties_ratio0")
    X = case["X"].copy()
    X[:, 1::5] = np.round(X[:, 1::5] * 4) / 4 + 1e-12 * np.arange(len(X))[:, None]   # equal in fp32, distinct in fp64
    x = oracle.make_x(X, case["y"])
    args = (x, case["y"], case["f"], case["nF"], case["nParts"], case["fp"], case["ratio"], case["k"], case["T"], case["mtry"], 7)
    rc, o1, o2 = oracle.o_cv(*args)
    a = H.cv_run(*args, tie_mode=0)
    assert np.array_equal(a["class1"].view(np.uint64), o1.view(np.uint64)) and np.array_equal(a["class2"].view(np.uint64), o2.view(np.uint64))
    # one tree per stage (walk_one), two walked side by side (walk_two; with T odd a pair straddles two partitions), and the
    # sequential multi-tree stages
    for tps, T in (("1", case["T"]), ("2", case["T"]), ("2", 3), ("4", case["T"])):
        monkeypatch.setenv("PSM_PREDICT_TPS", tps)
        args_t = args[:8] + (T,) + args[9:]
        want1, want2 = (o1, o2) if T == case["T"] else oracle.o_cv(*args_t)[1:]
        c = H.cv_run(*args_t, tie_mode=0)
        assert np.array_equal(c["class1"].view(np.uint64), want1.view(np.uint64)) and np.array_equal(c["class2"].view(np.uint64), want2.view(np.uint64)), (tps, T)
    monkeypatch.delenv("PSM_PREDICT_TPS")
    monkeypatch.setenv("PSM_PREDICT_WIDE", "1")
    b = H.cv_run(*args, tie_mode=0)
    assert np.array_equal(b["class1"].view(np.uint64), o1.view(np.uint64)) and np.array_equal(b["class2"].view(np.uint64), o2.view(np.uint64))


def test_small_batches_equal_one_batch(ctx, oracle):
    """the memory-budget batching must not change results"""
    case = make_case("fy_26")
    x, fd, rnd = _fold_setup(ctx, oracle, case)
    ctx.fold_knn(case["k"])
    per = len(fd["trngPos"]) * case["fp"] * (case["m"] + 1)
    draws = rnd.fill(per * case["nParts"])
    ctx.fold_run_partitions(case["nParts"], case["fp"], case["ratio"], case["T"], case["mtry"], 7, 0, 0, case["nParts"], draws)
    a1, a2 = ctx.fold_scores_get()
    ctx.set_batch_budget(1)   # forces one partition per batch
    ctx.fold_begin(fd["trngPos"], fd["trngNeg"], fd["testPos"], fd["testNeg"])
    ctx.fold_knn(case["k"])
    ctx.fold_run_partitions(case["nParts"], case["fp"], case["ratio"], case["T"], case["mtry"], 7, 0, 0, case["nParts"], draws)
    b1, b2 = ctx.fold_scores_get()
    ctx.set_batch_budget(0)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)


def test_curves_heavy_ties_and_sizes(ctx, oracle):
    rng = np.random.default_rng(5)
    for N, levels in ((1000, None), (200000, 21), (1 << 20, None)):
        s = rng.random(N)
        if levels:
            s = np.round(s * (levels - 1)) / (levels - 1)
        y = (rng.random(N) < 0.02 + 0.3 * s).astype(np.uint32)
        want = oracle.o_curves(y, s)
        got = ctx.curves(y, s, tie_mode=0)
        assert abs(got[0] - want[0]) < 1e-12 and abs(got[1] - want[1]) < 1e-12
        if not levels:   # no ties: the device sort must give the same order
            got1 = ctx.curves(y, s, tie_mode=1)
            assert abs(got1[0] - want[0]) < 1e-12 and abs(got1[1] - want[1]) < 1e-12


def test_error_codes(ctx, oracle):
    import parsmurf_b200 as psm
    case = make_case("fy_26")
    x, fd, rnd = _fold_setup(ctx, oracle, case)
    with pytest.raises(psm.PsmError) as e:
        ctx.batch_begin(4, 2, 3, 3, 2)
    assert e.value.code == -2
    ctx.fold_begin(fd["trngPos"][:1], fd["trngNeg"], fd["testPos"], fd["testNeg"])
    with pytest.raises(psm.PsmError) as e:
        ctx.fold_knn(5)
    assert e.value.code == -3     # the reference abort()s with < 2 positives
    ctx.fold_begin(fd["trngPos"], fd["trngNeg"][:41], fd["testPos"], fd["testNeg"])
    with pytest.raises(psm.PsmError) as e:
        ctx.batch_begin(10, 1, 1, 0, 10)   # 41 negatives / 10 parts: last chunk underflows in the reference
    assert e.value.code == -6

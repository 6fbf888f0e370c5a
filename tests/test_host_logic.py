"""Host-side logic of the product (no GPU): the C ABI library loads and exports every symbol the header declares, fails
loudly without a device, and the C++ host mirror's folds / train-test division / rand replica / JSON surface behave like
the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

import parsmurf_b200 as psm
from parsmurf_b200 import _capi, host_api as H
from helpers import make_case
from oracle import pyoracle as O


def test_capi_exports_every_declared_symbol():
    hdr = open(_capi.HEADER_PATH).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(psm_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    L = psm.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.psm_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(psm.PsmError) as e:
        psm.Context(0)
    assert e.value.code == -1
    x = np.zeros((10, 3)); y = np.zeros(10, np.uint32); y[:4] = 1
    with pytest.raises(psm.PsmError):
        H.cv_run(x, y, np.arange(10) % 2, 2, 2, 1, 1, 2, 2, 1, 1)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, dn, fn in os.walk(os.path.join(root, "parsmurf_b200")):
        if "_build" in dp:
            continue
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in txt and "psm_oracle" not in txt and "libparsmurf_ref" not in txt, f


def test_glibc_rand_replica_and_jump_ahead():
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    want = np.array([libc.rand() for _ in range(5000)], np.int32)
    assert np.array_equal(H.rand_fill(1, 5000), want)
    libc.srand(12345)
    want2 = np.array([libc.rand() for _ in range(100)], np.int32)
    assert np.array_equal(H.rand_fill(12345, 100), want2)
    w0 = _capi.glibc_initial_window(1)
    for n in (0, 1, 2, 30, 31, 32, 63, 64, 1000, 4999 - 40):
        w = _capi.glibc_jump(w0, n)
        seq = [int(v) for v in w]
        outs = []
        for _ in range(40):
            v = (seq[-31] + seq[-3]) & 0xFFFFFFFF
            seq.append(v); outs.append(v >> 1)
        assert np.array_equal(np.array(outs, np.int32), want[n:n + 40]), n


@pytest.mark.parametrize("name", ["fy_26", "ties_ratio0"])
def test_ttd_matches_oracle(name):
    c = make_case(name)
    o = O.o_ttd(c["y"], c["f"], c["nF"], O.GlibcRand(1))
    for f in range(c["nF"]):
        t = H.ttd_fold(c["y"], c["f"], c["nF"], f)
        for k_ in t:
            assert np.array_equal(t[k_], o[f][k_]), (f, k_)


def test_inner_cv_ttd_skips_fold_and_does_not_shuffle():
    c = make_case("fy_26")
    jump = 1
    full = [H.ttd_fold(c["y"], c["f"], c["nF"], f) for f in range(c["nF"])]
    for i in range(c["nF"] - 1):
        t = H.ttd_fold(c["y"], c["f"], c["nF"], i, fold_to_jump=jump)
        cur = i if i < jump else i + 1
        assert np.array_equal(t["testPos"], full[cur]["testPos"]) and np.array_equal(t["testNeg"], full[cur]["testNeg"])
        other = np.flatnonzero((c["f"] != cur) & (c["f"] != jump) & (c["y"] == 0))
        # fold order of the fold file, not shuffled (src/testtraindivider.cpp:214)
        want = np.concatenate([np.flatnonzero((c["f"] == k_) & (c["y"] == 0)) for k_ in range(c["nF"]) if k_ not in (cur, jump)])
        assert np.array_equal(t["trngNeg"], want) and len(other) == len(want)


def test_random_folds_are_stratified_round_robin():
    c = make_case("fy_26")
    t = [H.ttd_fold(c["y"], None, 5, f) for f in range(5)]
    npos = [len(v["testPos"]) for v in t]
    assert sum(npos) == int(c["y"].sum()) and max(npos) - min(npos) <= 1
    allidx = np.sort(np.concatenate([np.concatenate([v["testPos"], v["testNeg"]]) for v in t]))
    assert np.array_equal(allidx, np.arange(c["n"]))


CFG_GRID = """{
  "name": "cfgEx_gridTune",
  "exec": {"name": "parSMURF", "nProcs": 1, "ensThrd": 6, "rfThrd": 4, "seed": 18, "verboseLevel": 1, "saveTime": true,
           "timeFile": "timeout.txt", "printCfg": true, "mode": "cv", "optimizer": "grid"},
  "data": {"dataFile": "d.txt", "labelFile": "l.txt", "outFile": "predictions.txt",},
  "folds": {"nFolds": 5},
  "params": {"nParts": [10, 20, 30], "fp": [1, 2], "ratio": [1], "k": [5], "nTrees": [10], "mtry": [5, 7, 10]}
}"""


def test_config_surface():
    d = H.config_parse(CFG_GRID)   # note the trailing comma in "data": jsoncons accepts it, so must we
    assert d["nFolds"] == 5 and d["seed"] == 18 and d["woptimiz"] == 32 and d["wmode"] == 1 and d["ensThrd"] == 6 and d["rfThrd"] == 4
    assert d["gridSize"] == 18
    assert d["grid"][0] == dict(nParts=10, fp=1, ratio=1, k=5, nTrees=10, mtry=5)
    assert d["grid"][1]["mtry"] == 7 and d["grid"][3] == dict(nParts=10, fp=2, ratio=1, k=5, nTrees=10, mtry=5) and d["grid"][-1]["nParts"] == 30
    d = H.config_parse('{"exec": {"mode": "cv"}, "simulate": {"simulation": true, "prob": 0.02, "n": 1200, "m": 25}, "folds": {"nFolds": 10}, "params": {}}')
    assert d["simulate"] == 1 and d["n"] == 1200 and d["m"] == 25
    assert d["grid"] == [dict(nParts=3, fp=1, ratio=1, k=5, nTrees=50, mtry=0)]      # defaults of ArgHandle::checkConfig
    with pytest.raises(psm.PsmError):
        H.config_parse('{"exec": {"mode": "bogus"}}')
    with pytest.raises(psm.PsmError):
        H.config_parse('{"exec": ')


def test_rank_sharding_rules_cover_every_partition_once():
    """both rules of hyperSMURF::partRange: per-fold chunks (src/MPIHelper.cpp:29-37) and contiguous (fold, partition) units"""
    from parsmurf_b200 import host_api as H
    for nF, nP in ((5, 100), (10, 1000), (3, 7), (4, 1), (1, 5)):
        for world in (1, 2, 3, 4, 8, 16):
            for units in (False, True):
                sizes = []
                for f in range(nF):
                    r = [H.part_range(units, nF, nP, world, k, f) for k in range(world)]
                    owned = [(a, b) for a, b in r if b > a]
                    assert owned[0][0] == 0 and owned[-1][1] == nP and all(owned[i][1] == owned[i + 1][0] for i in range(len(owned) - 1)), (nF, nP, world, units, f, r)
                    sizes.append([b - a for a, b in r])
                tot = np.array(sizes).sum(axis=0)
                assert tot.sum() == nF * nP
                if units:
                    assert tot.max() - tot.min() <= 1                      # units balanced over the ranks
                    assert all(sum(1 for f in range(nF) if sizes[f][k] > 0) <= (nF + world - 1) // world + 1 for k in range(world))   # a rank opens few folds
                else:
                    assert all(max(srow) - min(srow) <= 1 for srow in sizes)   # every fold balanced over the ranks


def test_folds_from_assignment_large_input_threaded_pass():
    """n >= 2^18 takes the multi-threaded count/place passes of Folds::computeFolds: same arrays as the definition
    (every fold in ascending example order, src/folds.cpp:93-110)"""
    from parsmurf_b200 import host_api as H
    rng = np.random.default_rng(11)
    n, nF = 2_300_017, 3
    y = (rng.random(n) < 0.01).astype(np.uint32)
    f = rng.integers(0, nF, n).astype(np.uint32)
    for k in range(nF):
        got = H.ttd_fold(y, f, nF, k)
        assert np.array_equal(got["testPos"], np.flatnonzero((f == k) & (y == 1)))
        assert np.array_equal(got["testNeg"], np.flatnonzero((f == k) & (y == 0)))
        others = [q for q in range(nF) if q != k]
        assert np.array_equal(got["trngPos"], np.concatenate([np.flatnonzero((f == q) & (y == 1)) for q in others]))
        assert np.array_equal(np.sort(got["trngNeg"]), np.sort(np.concatenate([np.flatnonzero((f == q) & (y == 0)) for q in others])))


def test_backward_resolution_of_random_shuffle_equals_the_sequential_walk():
    """the algorithm of csrc/shuffle.cu, restated in numpy: the result of  for i in 1..n-1: swap(a[i], a[rand() % (i+1)])  is
    recovered per output slot by walking the steps backwards (largest later step that targets the slot ends the walk, else the
    slot's own step redirects), with 'largest step <= t with j_i == s' as one binary search in the sorted (j, i) pairs"""
    import bisect
    from parsmurf_b200 import host_api as H
    for n in (1, 2, 3, 17, 1000, 20000):
        draws = H.rand_fill(1, max(n - 1, 1)).astype(np.int64)
        a = np.arange(100, 100 + n)
        seq = a.copy()
        j = np.zeros(n, np.int64)
        for i in range(1, n):
            j[i] = draws[i - 1] % (i + 1)
            seq[i], seq[j[i]] = seq[j[i]], seq[i]
        shift = 24
        keys = sorted((int(j[i]) << shift) | i for i in range(1, n))
        out = np.empty(n, a.dtype)
        hops = 0
        for k in range(n):
            s, t = k, n - 1
            while True:
                hops += 1
                lo = bisect.bisect_right(keys, (s << shift) | t)
                if lo > 0 and (keys[lo - 1] >> shift) == s and (keys[lo - 1] & ((1 << shift) - 1)) > s:
                    s = keys[lo - 1] & ((1 << shift) - 1)
                    break
                if s == 0 or j[s] == s:
                    break
                t, s = s - 1, int(j[s])
            out[k] = a[s]
        assert np.array_equal(out, seq), n
        assert hops <= 3 * n + 3                                          # ~2 hops per slot


def test_parallel_sort_replica_equals_std_sort(oracle):
    """psm_host_sort_permutation (parallel replica of the reference's std::sort call, csrc/api.cu) == the permutation the oracle's
    identical std::sort call produces (oracle psmo_curves exports it), for continuous scores, heavy ties, all-equal and sorted
    inputs, on 1 and on several threads"""
    import ctypes as C
    import parsmurf_b200 as psm
    L = psm.lib()
    L.psm_host_sort_permutation.argtypes = [C.POINTER(C.c_double), C.c_uint64, C.POINTER(C.c_uint32), C.c_int]
    rng = np.random.default_rng(5)
    cases = []
    for n in (1, 2, 15, 16, 17, 33, 1000, 70_001, 400_000):
        cases.append(rng.random(n))
        cases.append(np.round(rng.random(n) * 20) / 20)               # 21 score levels
        cases.append(np.where(rng.random(n) < 0.9, 0.0, rng.random(n)))  # one huge run of ties
    cases += [np.zeros(100_000), np.arange(100_000, dtype=float), -np.arange(100_000, dtype=float)]
    for s in cases:
        s = np.ascontiguousarray(s, np.float64)
        labels = np.zeros(len(s), np.uint32); labels[:1] = 1
        _, _, want = oracle.o_curves(labels, s, want_perm=True)
        for threads in (1, 6):
            perm = np.zeros(len(s), np.uint32)
            assert L.psm_host_sort_permutation(s.ctypes.data_as(C.POINTER(C.c_double)), len(s), perm.ctypes.data_as(C.POINTER(C.c_uint32)), threads) == 0
            assert np.array_equal(perm.astype(np.uint64), want), (len(s), threads)


def test_grid_points_are_dealt_to_the_ranks_by_cost():
    """multi-rank grid search (hyperSMURF::gridOwners): every grid point has exactly one owner, the same table on every rank, and the
    longest-processing-time deal stays within 4/3 of the best possible balance"""
    import itertools
    from parsmurf_b200 import host_api as H
    grid = [(a, b, 1, 5, 10, c) for a, b, c in itertools.product([10, 20, 30], [1, 2], [5, 7, 10])]      # cfgEx/gridTune.json
    cost = np.array([g[0] * (g[1] + 1) * (g[2] + 1) * g[4] for g in grid], np.float64)
    for world in (1, 2, 3, 8, 32):
        own = H.grid_owners(grid, world)
        assert own.shape == (len(grid),) and own.max() < world
        assert np.array_equal(own, H.grid_owners(grid, world))
        load = np.bincount(own, weights=cost, minlength=world)
        assert load.sum() == cost.sum()
        assert load.max() <= max(cost.max(), 4.0 / 3.0 * cost.sum() / world + cost.max() / 3.0) + 1e-9
        if world <= len(grid):
            assert (load > 0).all()


def test_lane_pieces_cover_the_fold_shares():
    """unit sharding with fold lanes (hyperSMURF::laneCuts): every non-empty fold share gets at least one piece, the pieces never
    outnumber the target of one or two per lane (unless there are more shares than that), and tiny shares are not crumbled"""
    from parsmurf_b200 import host_api as H
    for shares, lanes in (([0, 11, 52, 0, 0], 4), ([63, 0, 0, 0, 0], 4), ([100, 100, 50, 0, 0], 4), ([1000, 250] + [0] * 8, 4), ([5, 5, 5, 5, 5], 2),
                          ([0, 0, 0], 4), ([37, 26, 0, 0, 0], 1), ([3, 0, 0], 4)):
        cuts = H.lane_cuts(shares, lanes)
        mine = sum(shares)
        for s_, c in zip(shares, cuts):
            assert (c == 0) == (s_ == 0)
            assert c <= max(1, s_)
        if mine:
            per_lane = -(-mine // lanes)
            target = lanes * (2 if per_lane >= 64 else 1)
            assert cuts.sum() <= max(target, sum(1 for s_ in shares if s_))
    assert list(H.lane_cuts([0, 11, 52, 0, 0], 4)) == [0, 1, 3, 0, 0]          # BASELINE configs[1] on 8 GPUs, rank 3: four pieces for four lanes

"""TEST INFRASTRUCTURE: ctypes bindings for the CPU oracle (oracle/_build/libpsm_oracle.so) and, where it
was built, the unmodified-reference harness (oracle/_ref/libparsmurf_ref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package `parsmurf_b200` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libpsm_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libparsmurf_ref.so")

dp = C.POINTER(C.c_double)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)
u32, u64, i32, vp = C.c_uint32, C.c_uint64, C.c_int32, C.c_void_p


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def build(target="oracle"):
    subprocess.check_call(["make", "-s", "-C", HERE, target])


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build("oracle")
        L = C.CDLL(ORACLE_SO)
        L.psmo_rand_create.restype = vp
        L.psmo_rand_create.argtypes = [u32]
        L.psmo_rand_destroy.argtypes = [vp]
        L.psmo_rand_next.restype = i32
        L.psmo_rand_next.argtypes = [vp]
        L.psmo_rand_fill.argtypes = [vp, i32p, u64]
        L.psmo_ttd_create.restype = vp
        L.psmo_ttd_create.argtypes = [u32p, u32p, u32, u32, vp]
        L.psmo_ttd_destroy.argtypes = [vp]
        L.psmo_ttd_fold_sizes.argtypes = [vp, u32, u32p]
        L.psmo_ttd_fold_indices.argtypes = [vp, u32, u32p, u32p, u32p, u32p]
        L.psmo_partition_sizes.argtypes = [u32] * 6 + [u32p]
        L.psmo_knn.argtypes = [dp, u32, u32, u32, i32p, dp]
        L.psmo_assemble.argtypes = [dp, u32, u32p, u32, u32p, u32, i32p, u32, u32, i32p, dp]
        L.psmo_forest_train.restype = vp
        L.psmo_forest_train.argtypes = [dp, u32, u32, u32, u32, u32]
        L.psmo_forest_destroy.argtypes = [vp]
        L.psmo_forest_tree_nodes.restype = u64
        L.psmo_forest_tree_nodes.argtypes = [vp, u32]
        L.psmo_forest_export_tree.argtypes = [vp, u32, u64p, u64p, u64p, dp, dp]
        L.psmo_forest_split_bytes.restype = u64
        L.psmo_forest_split_bytes.argtypes = [vp]
        L.psmo_tree_draws.argtypes = [u32, u32, u32, u32, u64, u64p, u32p]
        L.psmo_forest_predict.argtypes = [vp, dp, u32, u32, dp]
        L.psmo_curves.argtypes = [u32p, dp, u64, dp, u64p]
        L.psmo_cv.argtypes = [dp, u32, u32, u32p, u32p] + [u32] * 11 + [dp, dp]
        L.psmo_cv_parts.argtypes = [dp, u32, u32, u32p, u32p] + [u32] * 10 + [dp, dp]
        L.psmo_cv_ranges.argtypes = [dp, u32, u32, u32p, u32p] + [u32] * 8 + [u32p, u32p, dp, dp]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.ref_ctx_create.restype = vp
        L.ref_ctx_create.argtypes = [dp, u32, u32, u32p, u32p, u32, C.c_int]
        L.ref_ctx_destroy.argtypes = [vp]
        L.ref_ctx_fold_sizes.argtypes = [vp, u32, u32p]
        L.ref_ctx_fold_indices.argtypes = [vp, u32, u32p, u32p, u32p, u32p]
        L.ref_partition_maxsize.restype = u32
        L.ref_partition_maxsize.argtypes = [vp, u32, u32, u32]
        L.ref_sample_partition.restype = C.c_int64
        L.ref_sample_partition.argtypes = [vp, u32, u32, u32, u32, u32, u32, dp, u64, u32p, i32p, u64]
        L.ref_import.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, u64p, dp, u32p, u32p]
        L.ref_save.argtypes = [dp, dp, u32, u32p, C.c_char_p]
        L.ref_save.restype = None
        L.ref_knn.argtypes = [dp, C.c_int, C.c_int, C.c_int, i32p, dp, C.c_int]
        L.ref_knn_queries.argtypes = [dp, C.c_int, C.c_int, C.c_int, i32p, C.c_int, i32p, dp]
        L.ref_forest_train.restype = vp
        L.ref_forest_train.argtypes = [dp, u32, u32, u32, u32, u32, u32]
        L.ref_forest_destroy.argtypes = [vp]
        L.ref_forest_tree_nodes.restype = u64
        L.ref_forest_tree_nodes.argtypes = [vp, u32]
        L.ref_forest_export_tree.argtypes = [vp, u32, u64p, u64p, u64p, dp, dp]
        L.ref_forest_predict.argtypes = [vp, dp, u32, dp]
        L.ref_forest_save.argtypes = [vp, u32, C.c_char_p]
        L.ref_forest_file_predict.argtypes = [C.c_char_p, dp, u32, u32, u32, u32, u32, u32, dp]
        L.ref_curves.argtypes = [u32p, dp, u64, dp]
        L.ref_cv.restype = C.c_double
        L.ref_cv.argtypes = [dp, u32, u32, u32p, u32p] + [u32] * 10 + [i32, i32, C.c_int, dp, dp]
        L.ref_run_partitions.restype = C.c_double
        L.ref_run_partitions.argtypes = [vp] + [u32] * 12 + [dp, dp]
        L.ref_run_partitions_ex.restype = C.c_double
        L.ref_run_partitions_ex.argtypes = [vp] + [u32] * 12 + [dp, dp, i32p, u64, dp]
        L.ref_cv_grid.restype = C.c_double
        L.ref_cv_grid.argtypes = [dp, u32, u32, u32p, u32p, u32, u32p, u32, u32, u32, u32, C.c_int, dp, dp, dp]
        L.ref_mode_train.restype = C.c_double
        L.ref_mode_train.argtypes = [dp, u32, u32, u32p] + [u32] * 9 + [C.c_char_p, C.c_int]
        L.ref_config_parse.argtypes = [C.c_char_p, u32p, u32p, u32]
        L.ref_num_procs.restype = C.c_int
        L.ref_generate_random_set.argtypes = [u32, u32, C.c_double, u32, dp, u32p]
        L.ref_generate_random_set.restype = None
        _ref = L
    return _ref


# --------------------------------------------------------------------------- helpers shared by both
def make_x(X, y):
    """Row-major [n][m+1] with the reference's label column: 1.0 positives / 2.0 negatives (fileImport.cpp:88-91)."""
    n, m = X.shape
    x = np.empty((n, m + 1), dtype=np.float64)
    x[:, :m] = X
    x[:, m] = np.where(np.asarray(y) > 0, 1.0, 2.0)
    return np.ascontiguousarray(x)


class _Forest:
    def __init__(self, lib, prefix, handle, T):
        self.lib, self.prefix, self.h, self.T = lib, prefix, handle, T

    def tree(self, t):
        nn = getattr(self.lib, self.prefix + "forest_tree_nodes")(self.h, t)
        left = np.zeros(nn, np.uint64); right = np.zeros(nn, np.uint64); var = np.zeros(nn, np.uint64)
        val = np.zeros(nn); frac = np.zeros((nn, 2))
        getattr(self.lib, self.prefix + "forest_export_tree")(self.h, t, _p(left, u64p), _p(right, u64p), _p(var, u64p), _p(val, dp), _p(frac, dp))
        return dict(left=left, right=right, var=var, val=val, frac=frac)

    def trees(self):
        return [self.tree(t) for t in range(self.T)]

    def close(self):
        if self.h:
            getattr(self.lib, self.prefix + "forest_destroy")(self.h)
            self.h = None

    def __del__(self):
        self.close()


# --------------------------------------------------------------------------- oracle API
class GlibcRand:
    def __init__(self, seed=1):
        self.L = oracle_lib(); self.h = self.L.psmo_rand_create(seed)

    def fill(self, count):
        out = np.empty(count, np.int32); self.L.psmo_rand_fill(self.h, _p(out, i32p), count); return out

    def __del__(self):
        self.L.psmo_rand_destroy(self.h)


def o_ttd(labels, folds, nFolds, rand=None):
    """Returns list of per-fold dicts of index arrays; shuffles training negatives with `rand` (GlibcRand) if given."""
    L = oracle_lib()
    labels = np.ascontiguousarray(labels, np.uint32); folds = np.ascontiguousarray(folds, np.uint32)
    h = L.psmo_ttd_create(_p(labels, u32p), _p(folds, u32p), len(labels), nFolds, rand.h if rand else None)
    out = []
    for f in range(nFolds):
        s = np.zeros(4, np.uint32); L.psmo_ttd_fold_sizes(h, f, _p(s, u32p))
        a = [np.zeros(int(v), np.uint32) for v in s]
        L.psmo_ttd_fold_indices(h, f, *[_p(v, u32p) for v in a])
        out.append(dict(trngPos=a[0], trngNeg=a[1], testPos=a[2], testNeg=a[3]))
    L.psmo_ttd_destroy(h)
    return out


def o_partition_sizes(P, Nneg, nParts, fp, ratio, p):
    out = np.zeros(6, np.uint32)
    rc = oracle_lib().psmo_partition_sizes(P, Nneg, nParts, fp, ratio, p, _p(out, u32p))
    return rc, dict(zip(["chunk", "negAvail", "negOffset", "nPosOut", "nNegOut", "Ntr"], [int(v) for v in out]))


def o_knn(pts, k):
    pts = np.ascontiguousarray(pts, np.float64); P, dim = pts.shape
    idx = np.zeros((P, k), np.int32); dist = np.zeros((P, k))
    oracle_lib().psmo_knn(_p(pts, dp), P, dim, k, _p(idx, i32p), _p(dist, dp))
    return idx, dist


def o_assemble(x, posIdx, negIdx, nNegOut, nn, fp, draws):
    n, ld = x.shape; m = ld - 1; P = len(posIdx); localk = nn.shape[1]
    posIdx = np.ascontiguousarray(posIdx, np.uint32); negIdx = np.ascontiguousarray(negIdx, np.uint32)
    nn = np.ascontiguousarray(nn, np.int32); draws = np.ascontiguousarray(draws, np.int32)
    out = np.zeros(((fp + 1) * P + nNegOut, ld))
    rc = oracle_lib().psmo_assemble(_p(x, dp), m, _p(posIdx, u32p), P, _p(negIdx, u32p), nNegOut, _p(nn, i32p), localk, fp, _p(draws, i32p), _p(out, dp))
    return rc, out


def o_forest_train(colmajor, T, mtry, seed):
    colmajor = np.ascontiguousarray(colmajor, np.float64); ld, Ntr = colmajor.shape
    L = oracle_lib()
    f = _Forest(L, "psmo_", L.psmo_forest_train(_p(colmajor, dp), Ntr, ld - 1, T, mtry, seed), T)
    f.split_bytes = L.psmo_forest_split_bytes(f.h)
    return f


def o_forest_predict(forest, rows, m=None):
    rows = np.ascontiguousarray(rows, np.float64); Nte, ld = rows.shape
    preds = np.zeros((Nte, 2))
    oracle_lib().psmo_forest_predict(forest.h, _p(rows, dp), Nte, ld, _p(preds, dp))
    return preds


def o_tree_draws(tree_seed, Ntr, m, mtry, nNodes):
    boot = np.zeros(Ntr, np.uint64); cand = np.zeros((nNodes, mtry), np.uint32)
    oracle_lib().psmo_tree_draws(tree_seed, Ntr, m, mtry, nNodes, _p(boot, u64p), _p(cand, u32p))
    return boot, cand


def o_curves(labels, preds, want_perm=False):
    labels = np.ascontiguousarray(labels, np.uint32); preds = np.ascontiguousarray(preds, np.float64)
    out = np.zeros(2); perm = np.zeros(len(labels), np.uint64) if want_perm else None
    oracle_lib().psmo_curves(_p(labels, u32p), _p(preds, dp), len(labels), _p(out, dp), _p(perm, u64p))
    return (out[0], out[1], perm) if want_perm else (out[0], out[1])


def o_cv(x, labels, folds, nFolds, nParts, fp, ratio, k, nTrees, mtry, seed, fold0=0, fold1=None, nthreads=1):
    n, ld = x.shape
    labels = np.ascontiguousarray(labels, np.uint32); folds = np.ascontiguousarray(folds, np.uint32)
    c1 = np.zeros(n); c2 = np.zeros(n)
    rc = oracle_lib().psmo_cv(_p(x, dp), n, ld - 1, _p(labels, u32p), _p(folds, u32p), nFolds, nParts, fp, ratio, k, nTrees, mtry, seed,
                              fold0, nFolds if fold1 is None else fold1, nthreads, _p(c1, dp), _p(c2, dp))
    return rc, c1, c2


def o_cv_parts(x, labels, folds, nFolds, nParts, fp, ratio, k, nTrees, mtry, seed, part0, part1):
    n, ld = x.shape
    labels = np.ascontiguousarray(labels, np.uint32); folds = np.ascontiguousarray(folds, np.uint32)
    c1 = np.zeros(n); c2 = np.zeros(n)
    rc = oracle_lib().psmo_cv_parts(_p(x, dp), n, ld - 1, _p(labels, u32p), _p(folds, u32p), nFolds, nParts, fp, ratio, k, nTrees, mtry, seed,
                                    part0, part1, _p(c1, dp), _p(c2, dp))
    return rc, c1, c2


def o_cv_ranges(x, labels, folds, nFolds, nParts, fp, ratio, k, nTrees, mtry, seed, first, last):
    """only partitions [first[f], last[f]) of fold f contribute (a unit-sharded rank's share)"""
    n, ld = x.shape
    labels = np.ascontiguousarray(labels, np.uint32); folds = np.ascontiguousarray(folds, np.uint32)
    first = np.ascontiguousarray(first, np.uint32); last = np.ascontiguousarray(last, np.uint32)
    c1 = np.zeros(n); c2 = np.zeros(n)
    rc = oracle_lib().psmo_cv_ranges(_p(x, dp), n, ld - 1, _p(labels, u32p), _p(folds, u32p), nFolds, nParts, fp, ratio, k, nTrees, mtry, seed,
                                     _p(first, u32p), _p(last, u32p), _p(c1, dp), _p(c2, dp))
    return rc, c1, c2


# --------------------------------------------------------------------------- reference API
def ref_generate_random_set(n, m, prob, seed):
    """the reference's generateRandomSet (src/HyperSMURFUtils.cpp:40-63)"""
    L = ref_lib()
    x = np.empty((n, m + 1)); y = np.empty(n, np.uint32)
    L.ref_generate_random_set(n, m, prob, seed, _p(x, dp), _p(y, u32p))
    return x, y


def ref_import(data_file, label_file, fold_file=None):
    """the reference's Importer::import -> (x flat with the label column appended per row, y, f, nFolds)"""
    L = ref_lib()
    sz = np.zeros(4, np.uint64)
    enc = lambda s: None if s is None else str(s).encode()
    if L.ref_import(enc(data_file), enc(label_file), enc(fold_file), _p(sz, u64p), None, None, None) != 0:
        raise RuntimeError("reference importer failed")
    x = np.zeros(int(sz[0])); y = np.zeros(int(sz[1]), np.uint32); f = np.zeros(int(sz[2]), np.uint32)
    L.ref_import(enc(data_file), enc(label_file), enc(fold_file), _p(sz, u64p), _p(x, dp), _p(y, u32p), _p(f, u32p))
    return x, y, f, int(sz[3])


def ref_save(c1, c2, labels, path):
    c1 = np.ascontiguousarray(c1, np.float64); c2 = np.ascontiguousarray(c2, np.float64)
    lab = None if labels is None else np.ascontiguousarray(labels, np.uint32)
    ref_lib().ref_save(_p(c1, dp), _p(c2, dp), len(c1), _p(lab, u32p), str(path).encode())


class RefCtx:
    def __init__(self, x, labels, folds, nFolds, reset_rand=True):
        self.L = ref_lib()
        self.x = np.ascontiguousarray(x, np.float64)
        self.labels = np.ascontiguousarray(labels, np.uint32); self.folds = np.ascontiguousarray(folds, np.uint32)
        self.n, ld = self.x.shape; self.m = ld - 1; self.nFolds = nFolds
        self.h = self.L.ref_ctx_create(_p(self.x, dp), self.n, self.m, _p(self.labels, u32p), _p(self.folds, u32p), nFolds, int(reset_rand))

    def fold(self, f):
        s = np.zeros(4, np.uint32); self.L.ref_ctx_fold_sizes(self.h, f, _p(s, u32p))
        a = [np.zeros(int(v), np.uint32) for v in s]
        self.L.ref_ctx_fold_indices(self.h, f, *[_p(v, u32p) for v in a])
        return dict(trngPos=a[0], trngNeg=a[1], testPos=a[2], testNeg=a[3])

    def sample(self, fold, part, nParts, fp, ratio, k):
        cap = self.L.ref_partition_maxsize(self.h, nParts, fp, ratio)
        buf = np.zeros((cap, self.m + 1)); sizes = np.zeros(3, np.uint32)
        P = len(self.fold(fold)["trngPos"])
        dcap = P * fp * (self.m + 1) + 16
        draws = np.zeros(dcap, np.int32)
        nd = self.L.ref_sample_partition(self.h, fold, part, nParts, fp, ratio, k, _p(buf, dp), buf.size, _p(sizes, u32p), _p(draws, i32p), dcap)
        assert nd >= 0
        return buf[: int(sizes[2])].copy(), dict(trngPos=int(sizes[0]), trngNeg=int(sizes[1]), lineLen=int(sizes[2])), draws[:nd].copy()

    def run_partitions(self, fold, p0, p1, nParts, fp, ratio, k, nTrees, mtry, seed, nThr, rfThr=1):
        c1 = np.zeros(self.n); c2 = np.zeros(self.n)
        secs = self.L.ref_run_partitions(self.h, fold, p0, p1, nParts, fp, ratio, k, nTrees, mtry, seed, nThr, rfThr, _p(c1, dp), _p(c2, dp))
        return secs, c1, c2

    def run_partitions_ex(self, fold, p0, p1, nParts, fp, ratio, k, nTrees, mtry, seed, nThr, rfThr=1):
        """run_partitions that also exports, per partition, the rand() draws its SamplerKNN::sample() consumed ([p1-p0][P*fp*(m+1)])
        and the reference's raw predictions ([p1-p0][Nte][2]) -- what makes a multi-threaded reference run comparable bit for bit"""
        fd = self.fold(fold)
        P, Nte = len(fd["trngPos"]), len(fd["testPos"]) + len(fd["testNeg"])
        per = P * fp * (self.m + 1)
        c1 = np.zeros(self.n); c2 = np.zeros(self.n)
        draws = np.zeros((p1 - p0, max(per, 1)), np.int32); preds = np.zeros((p1 - p0, Nte, 2))
        secs = self.L.ref_run_partitions_ex(self.h, fold, p0, p1, nParts, fp, ratio, k, nTrees, mtry, seed, nThr, rfThr, _p(c1, dp), _p(c2, dp),
                                            _p(draws, i32p), per, _p(preds, dp))
        if per and (draws[:, 0] == -2).any():
            raise RuntimeError("reference consumed a different number of rand() draws than P*fp*(m+1)")
        return secs, c1, c2, draws[:, :per], preds, fd

    def __del__(self):
        self.L.ref_ctx_destroy(self.h)


def scale_parity(gpu_ctx, fd, draws, preds, p0, p1, nParts, fp, ratio, k, nTrees, mtry, seed, fold, ref_class1=None):
    """CHECKER.  `gpu_ctx` is a product context (parsmurf_b200.Context, dataset already uploaded / attached); `fd`, `draws`,
    `preds` come from RefCtx.run_partitions_ex for partitions [p0, p1) of `fold`.  The CUDA path runs the same partitions with
    the reference's draws injected; its partial fold scores must equal, BIT FOR BIT, the reference's per-partition
    predictions times 1/nParts summed in ascending partition order from zero (src/sampler.cpp:116-133 at ensThrd 1).
    Returns dict(partitions, rows, max_abs_diff, mismatching_rows[, max_abs_diff_vs_reference_sum])."""
    gpu_ctx.fold_begin(fd["trngPos"], fd["trngNeg"], fd["testPos"], fd["testNeg"])
    gpu_ctx.fold_knn(k)
    gpu_ctx.fold_run_partitions(nParts, fp, ratio, nTrees, mtry, seed, fold, p0, p1, np.ascontiguousarray(draws).reshape(-1) if fp else None)
    g1, g2 = gpu_ctx.fold_scores_get()
    inv = 1.0 / nParts
    e1 = np.zeros(preds.shape[1]); e2 = np.zeros(preds.shape[1])
    for i in range(preds.shape[0]):                       # ascending partition order, one rounding per multiply and per add
        e1 = e1 + preds[i, :, 0] * inv
        e2 = e2 + preds[i, :, 1] * inv
    d = max(float(np.abs(g1 - e1).max(initial=0.0)), float(np.abs(g2 - e2).max(initial=0.0)))
    bad = int(np.count_nonzero((g1.view(np.uint64) != e1.view(np.uint64)) | (g2.view(np.uint64) != e2.view(np.uint64))))
    out = dict(partitions=int(p1 - p0), rows=int(preds.shape[1]), max_abs_diff=d, mismatching_rows=bad)
    if ref_class1 is not None:   # the reference's own accumulation (completion order under the lock when multi-threaded)
        test = np.concatenate([fd["testPos"], fd["testNeg"]])
        out["max_abs_diff_vs_reference_sum"] = float(np.abs(g1 - ref_class1[test]).max(initial=0.0))
    return out


def r_config_parse(path):
    """the reference's ArgHandle on `--cfg path` (exit()s the process on an invalid configuration: call it in a child process)"""
    out = np.zeros(10, np.uint32); grid = np.zeros((4096, 6), np.uint32)
    ref_lib().ref_config_parse(str(path).encode(), _p(out, u32p), _p(grid, u32p), 4096)
    keys = ["nFolds", "seed", "wmode", "woptimiz", "gridSize", "ensThrd", "rfThrd", "simulate", "n", "m"]
    d = {k_: int(out[i]) for i, k_ in enumerate(keys)}
    d["grid"] = [dict(zip(["nParts", "fp", "ratio", "k", "nTrees", "mtry"], [int(v) for v in row])) for row in grid[: d["gridSize"]]]
    return d


def r_knn(pts, k, brute=False):
    pts = np.ascontiguousarray(pts, np.float64); P, dim = pts.shape
    idx = np.zeros((P, k), np.int32); dist = np.zeros((P, k))
    ref_lib().ref_knn(_p(pts, dp), P, dim, k, _p(idx, i32p), _p(dist, dp), int(brute))
    return idx, dist


def r_knn_queries(pts, k, queries):
    """the reference's ANN kd-tree over all of `pts`, searched for the query rows `queries` only"""
    pts = np.ascontiguousarray(pts, np.float64); P, dim = pts.shape
    q = np.ascontiguousarray(queries, np.int32)
    idx = np.zeros((len(q), k), np.int32); dist = np.zeros((len(q), k))
    ref_lib().ref_knn_queries(_p(pts, dp), P, dim, k, _p(q, i32p), len(q), _p(idx, i32p), _p(dist, dp))
    return idx, dist


def r_forest_train(colmajor, T, mtry, seed, rfThr=1):
    colmajor = np.ascontiguousarray(colmajor, np.float64); ld, Ntr = colmajor.shape
    L = ref_lib()
    return _Forest(L, "ref_", L.ref_forest_train(_p(colmajor, dp), Ntr, ld - 1, T, mtry, seed, rfThr), T)


def r_forest_predict(forest, test_colmajor):
    test_colmajor = np.ascontiguousarray(test_colmajor, np.float64); ld, Nte = test_colmajor.shape
    preds = np.zeros((Nte, 2))
    ref_lib().ref_forest_predict(forest.h, _p(test_colmajor, dp), Nte, _p(preds, dp))
    return preds


def r_forest_save(forest, part, dirname):
    """reference rfRanger::saveForest -> <dirname>/<part>.out.forest"""
    rc = ref_lib().ref_forest_save(forest.h, part, str(dirname).encode())
    assert rc == 0
    return "%s/%d.out.forest" % (dirname, part)


def r_forest_file_predict(filename, test_colmajor, T, mtry, seed, rfThr=1):
    """reference predict mode: load a .forest file and predict a column-major test matrix [m+1][Nte]"""
    test_colmajor = np.ascontiguousarray(test_colmajor, np.float64); ld, Nte = test_colmajor.shape
    preds = np.zeros((Nte, 2))
    rc = ref_lib().ref_forest_file_predict(str(filename).encode(), _p(test_colmajor, dp), Nte, ld - 1, T, mtry, seed, rfThr, _p(preds, dp))
    assert rc == 0
    return preds


def r_curves(labels, preds):
    labels = np.ascontiguousarray(labels, np.uint32); preds = np.ascontiguousarray(preds, np.float64)
    out = np.zeros(2)
    ref_lib().ref_curves(_p(labels, u32p), _p(preds, dp), len(labels), _p(out, dp))
    return out[0], out[1]


def r_cv(x, labels, folds, nFolds, nParts, fp, ratio, k, nTrees, mtry, seed, nThr=1, rfThr=1, minFold=-1, maxFold=-1, reset_rand=True):
    n, ld = x.shape
    x = np.ascontiguousarray(x, np.float64)
    labels = np.ascontiguousarray(labels, np.uint32); folds = np.ascontiguousarray(folds, np.uint32)
    c1 = np.zeros(n); c2 = np.zeros(n)
    secs = ref_lib().ref_cv(_p(x, dp), n, ld - 1, _p(labels, u32p), _p(folds, u32p), nFolds, nParts, fp, ratio, k, nTrees, mtry, seed,
                            nThr, rfThr, minFold, maxFold, int(reset_rand), _p(c1, dp), _p(c2, dp))
    return secs, c1, c2


def r_mode_train(x, labels, forest_dir, nParts, fp, ratio, k, nTrees, mtry, seed, nThr=1, rfThr=1):
    """the reference's exec.mode "train" as its main() drives it (one randomly generated fold holding every example): writes
    <forest_dir>/<p>.out.forest for every partition"""
    n, ld = x.shape
    x = np.ascontiguousarray(x, np.float64); labels = np.ascontiguousarray(labels, np.uint32)
    return ref_lib().ref_mode_train(_p(x, dp), n, ld - 1, _p(labels, u32p), nParts, fp, ratio, k, nTrees, mtry, seed, nThr, rfThr, str(forest_dir).encode(), 1)


def r_cv_grid(x, labels, folds, nFolds, grid, seed, nThr=1, rfThr=1):
    """reference grid search (optimizer "grid"); grid = list of (nParts, fp, ratio, k, nTrees, mtry).  Runs in a scratch cwd
    because the reference appends fold<k>.dat files.  Returns (class1, class2, gridMetrics[G][2] of the last outer fold)."""
    import tempfile
    n, ld = x.shape
    x = np.ascontiguousarray(x, np.float64)
    labels = np.ascontiguousarray(labels, np.uint32); folds = np.ascontiguousarray(folds, np.uint32)
    g = np.ascontiguousarray(np.array(grid, np.uint32).reshape(-1, 6))
    c1 = np.zeros(n); c2 = np.zeros(n); gm = np.zeros((len(g), 2))
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            ref_lib().ref_cv_grid(_p(x, dp), n, ld - 1, _p(labels, u32p), _p(folds, u32p), nFolds, _p(g, u32p), len(g), seed, nThr, rfThr, 1,
                                  _p(c1, dp), _p(c2, dp), _p(gm, dp))
        finally:
            os.chdir(cwd)
    return c1, c2, gm

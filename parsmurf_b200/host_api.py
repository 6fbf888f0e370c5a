"""ctypes binding of the C++ host layer (libparsmurf_host.so: parsmurf::hyperSMURF and friends, parsmurf_b200/host/).
`cv_run` is the call a user makes: the whole cross-validation through the reference-shaped host classes."""
import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import KERNEL_CLASSES, PsmError, dp, u32, u32p, u64, u64p, vp, i32p, _p

PHASES = ["setup_upload", "folds_ttd", "fold_begin", "smote_draws", "assemble", "train", "predict", "reduce", "scatter", "curves", "teardown"]
REDUCE_FN = C.CFUNCTYPE(C.c_int, vp, vp, u64)
_hlib = None


def hlib():
    global _hlib
    if _hlib is None:
        _capi.lib()   # RTLD_GLOBAL: resolves libparsmurf_b200.so for the host library
        if not os.path.exists(_capi.HOST_LIB_PATH):
            raise ImportError("parsmurf_b200: %s is missing -- run __graft_entry__.build()" % _capi.HOST_LIB_PATH)
        L = C.CDLL(_capi.HOST_LIB_PATH)
        L.psmh_cv_run.argtypes = [vp, C.c_int, u32, u32, u32p, u32p] + [u32] * 8 + [C.c_int, vp, C.c_int, C.c_int, u32, u32, REDUCE_FN, vp, u64,
                                  C.c_int, dp, dp, dp, dp, dp, u64p, u64p, dp, vp, C.c_char_p, C.c_size_t]
        L.psmh_set_fold_lanes.argtypes = [u32]
        L.psmh_set_fold_lanes.restype = None
        L.psmh_get_fold_lanes.restype = u32
        L.psmh_set_shard_units.argtypes = [C.c_int]
        L.psmh_set_shard_units.restype = None
        L.psmh_part_range.argtypes = [C.c_int, u32, u32, u32, u32, u32, u32p, u32p]
        L.psmh_part_range.restype = None
        L.psmh_device_create.restype = vp
        L.psmh_device_create.argtypes = [C.c_int, vp]
        L.psmh_device_destroy.argtypes = [vp]
        L.psmh_set_scores_root_only.argtypes = [C.c_int]
        L.psmh_set_scores_root_only.restype = None
        L.psmh_comm_unique_id.argtypes = [vp]
        L.psmh_device_comm_init.argtypes = [vp, vp, u32, u32, C.c_char_p, C.c_size_t]
        L.psmh_device_comm_info.argtypes = [vp, u32p, u32p, C.POINTER(C.c_int)]
        L.psmh_cv_grid_run.argtypes = [dp, u32, u32, u32p, u32p, u32, u32p, u32, u32, C.c_int, C.c_int, dp, dp, dp, dp, C.c_char_p, C.c_size_t]
        L.psmh_rand_fill.argtypes = [u32, i32p, u64]
        L.psmh_ttd_fold.argtypes = [u32p, u32p, u32, u32, u32, C.c_int64, u32p, u32p, u32p, u32p, u32p]
        L.psmh_config_parse.argtypes = [C.c_char_p, u32p, u32p, u32, C.c_char_p, C.c_size_t]
        _hlib = L
    return _hlib


def cv_run(x, labels, folds, nFolds, nParts, fp, ratio, k, nTrees, mtry, seed, device=0, stream=None, tie_mode=0, eval_fold_curves=True,
           rank=0, world=1, reduce=None, batch_budget=0, profile=False, x_device_ptr=None, n=None, m=None, dev_handle=None, out=None,
           fold_lanes=None, shard=None, scores_root_only=False):
    """Run the whole CV on one GPU (or this rank's share of the partitions when world > 1).

    x: host row-major float64 [n][m+1] (label column 1.0/2.0), or None with x_device_ptr/n/m for a dataset already in HBM.
    folds: per-example fold ids, or None for the reference's random stratified folds.
    reduce(dev_ptr, count): sums `count` float64 at device pointer `dev_ptr` across ranks in place (world > 1).
    out: optional (class1, class2) float64 arrays of length n to reuse; they are overwritten (a fresh 8 MB array costs
    more in first-touch page faults than the copy into it).
    fold_lanes: how many folds are in flight at once (one context + stream + host thread each, same results); None keeps the
    library default (PSM_FOLD_LANES or 4).  Kernel times / launch counts are summed over the lanes.
    shard (world > 1): "unit" -- the nFolds x nParts (fold, partition) units are cut into contiguous ranges, one per rank, and
    `reduce` is called once for the summed score arrays [2n] and once for the per-fold metrics [2 nFolds]; "fold" -- the
    reference's MPI rule: every rank takes its chunk of every fold's partitions, `reduce` is called once per fold [2 Nte].
    None keeps the library default (PSM_SHARD or "unit").
    reduce=None with world > 1: the sum goes through the library's own NCCL communicator (HostDevice.comm_init first).
    scores_root_only (world > 1): only rank 0 copies the final class1/class2 arrays to the host, like the master of the
    reference's MPI build; on the other ranks the returned arrays are not written.
    Returns dict(class1, class2, auroc, auprc, fold_metrics, kernel_ms, kernel_launches, stats)."""
    L = hlib()
    L.psmh_set_fold_lanes(0 if fold_lanes is None else int(fold_lanes))
    L.psmh_set_shard_units(-1 if shard is None else {"unit": 1, "fold": 0}[shard])
    L.psmh_set_scores_root_only(1 if scores_root_only else 0)
    if x is not None:
        assert x.dtype == np.float64 and x.flags.c_contiguous
        n, m = x.shape[0], x.shape[1] - 1
        xptr, isdev = x.ctypes.data, 0
    else:
        xptr, isdev = x_device_ptr, 1
    labels = np.ascontiguousarray(labels, np.uint32)
    folds = None if folds is None else np.ascontiguousarray(folds, np.uint32)
    if out is not None:
        c1, c2 = out
        assert c1.dtype == np.float64 and c2.dtype == np.float64 and c1.shape == (n,) and c2.shape == (n,) and c1.flags.c_contiguous and c2.flags.c_contiguous
        # (no zero-fill: psmh_cv_run overwrites both arrays)
    else:
        c1 = np.empty(n); c2 = np.empty(n)
    met = np.zeros(2); fm = np.zeros((nFolds, 2))
    kms = np.zeros(len(KERNEL_CLASSES)); kl = np.zeros(len(KERNEL_CLASSES), np.uint64); st = np.zeros(6, np.uint64)
    err = C.create_string_buffer(512)
    ph = np.zeros(len(PHASES))

    def _cb(user, ptr, count):
        try:
            reduce(ptr, count)
            return 0
        except Exception as e:   # never unwind through C
            print("reduce callback failed:", e)
            return 1
    cb = REDUCE_FN(_cb) if reduce is not None else REDUCE_FN()
    rc = L.psmh_cv_run(vp(xptr), isdev, n, m, _p(labels, u32p), _p(folds, u32p), nFolds, nParts, fp, ratio, k, nTrees, mtry, seed, device,
                       vp(stream) if stream else None, tie_mode, int(eval_fold_curves), rank, world, cb, None, batch_budget, int(profile),
                       _p(c1, dp), _p(c2, dp), _p(met, dp), _p(fm, dp), _p(kms, dp), _p(kl, u64p), _p(st, u64p), _p(ph, dp), vp(dev_handle) if dev_handle else None, err, 512)
    if rc != 0:
        raise PsmError(rc, err.value.decode())
    return dict(class1=c1, class2=c2, auroc=float(met[0]), auprc=float(met[1]), fold_metrics=fm,
                kernel_ms={k_: float(kms[i]) for i, k_ in enumerate(KERNEL_CLASSES)},
                kernel_launches={k_: int(kl[i]) for i, k_ in enumerate(KERNEL_CLASSES)},
                stats=dict(split_bytes=int(st[0]), nodes=int(st[1]), levels=int(st[2]), mixed_tie_pairs=int(st[3]), curves_host_sorts=int(st[4]), curves_max_tie_spread=float(st[5]) * 1e-15), fold_lanes=min(int(L.psmh_get_fold_lanes()), max(int(nFolds), 1)),
                host_phases_s={k_: float(ph[i]) for i, k_ in enumerate(PHASES)})


def _mode_run(mode, forest_dir, x, labels, nParts, fp, ratio, k, nTrees, mtry, seed, device, tie_mode, rank, world, reduce):
    L = hlib()
    L.psmh_mode_run.argtypes = [C.c_int, C.c_char_p, vp, u32, u32, u32p] + [u32] * 7 + [C.c_int, C.c_int, u32, u32, REDUCE_FN, vp, dp, dp, dp, C.c_char_p, C.c_size_t]
    assert x.dtype == np.float64 and x.flags.c_contiguous
    n, m = x.shape[0], x.shape[1] - 1
    labels = np.ascontiguousarray(labels, np.uint32)
    c1 = np.zeros(n); c2 = np.zeros(n); met = np.zeros(2)
    err = C.create_string_buffer(512)

    def _cb(user, ptr, count):
        try:
            reduce(ptr, count)
            return 0
        except Exception as e:   # never unwind through C
            print("reduce callback failed:", e)
            return 1
    cb = REDUCE_FN(_cb) if reduce is not None else REDUCE_FN()
    rc = L.psmh_mode_run(mode, str(forest_dir).encode(), vp(x.ctypes.data), n, m, _p(labels, u32p), nParts, fp, ratio, k, nTrees, mtry, seed, device,
                         tie_mode, rank, world, cb, None, _p(c1, dp), _p(c2, dp), _p(met, dp), err, 512)
    if rc != 0:
        raise PsmError(rc, err.value.decode())
    return dict(class1=c1, class2=c2, auroc=float(met[0]), auprc=float(met[1]))


def train_run(x, labels, forest_dir, nParts, fp, ratio, k, nTrees, mtry, seed, device=0, rank=0, world=1):
    """exec.mode "train" (src/hyperSMURF.cpp:349-360): one forest per partition grown on all the data, written as
    <forest_dir>/<p>.out.forest in ranger's binary forest layout.  With world > 1 every rank writes its own partitions."""
    _mode_run(2, forest_dir, x, labels, nParts, fp, ratio, k, nTrees, mtry, seed, device, 0, rank, world, None)


def predict_run(x, labels, forest_dir, nParts, nTrees, mtry, seed, device=0, tie_mode=0, rank=0, world=1, reduce=None):
    """exec.mode "predict" (src/hyperSMURF.cpp:401-446): score every example with the forests saved in forest_dir."""
    return _mode_run(4, forest_dir, x, labels, nParts, 0, 0, 0, nTrees, mtry, seed, device, tie_mode, rank, world, reduce)


def forest_file_rewrite(in_path, out_path):
    """read a ranger .forest file with the host library's reader and write it back with its writer; returns the tree count"""
    L = hlib()
    L.psmh_forest_file_rewrite.argtypes = [C.c_char_p, C.c_char_p, u64p, C.c_char_p, C.c_size_t]
    nt = np.zeros(1, np.uint64); err = C.create_string_buffer(512)
    rc = L.psmh_forest_file_rewrite(str(in_path).encode(), str(out_path).encode(), _p(nt, u64p), err, 512)
    if rc != 0:
        raise PsmError(rc, err.value.decode())
    return int(nt[0])


class _ScratchCwd:
    """the grid search appends fold<k>.dat to the working directory like the reference (src/hyperSMURF.cpp:191-199): run it in a
    scratch directory.  The working directory is process-wide, so concurrent callers (ranks emulated as threads) share one:
    the first one in creates and enters it, the last one out restores the old directory and removes it."""

    def __init__(self):
        import threading
        self.lock, self.users, self.tmp, self.cwd = threading.Lock(), 0, None, None

    def __enter__(self):
        import tempfile
        with self.lock:
            if self.users == 0:
                self.cwd = os.getcwd()
                self.tmp = tempfile.TemporaryDirectory()
                os.chdir(self.tmp.name)
            self.users += 1

    def __exit__(self, *exc):
        with self.lock:
            self.users -= 1
            if self.users == 0:
                os.chdir(self.cwd)
                self.tmp.cleanup()
                self.tmp = None
        return False


_scratch_cwd = _ScratchCwd()


def cv_grid_run(x, labels, folds, nFolds, grid, seed, device=0, tie_mode=0, stream=None, rank=0, world=1, reduce=None, x_device_ptr=None, n=None, m=None,
                dev_handle=None):
    """Grid search like cfgEx/gridTune.json (exec.optimizer "grid"): grid = list of (nParts, fp, ratio, k, nTrees, mtry).
    Appends fold<k>.dat files to a scratch directory.  x / x_device_ptr, rank / world / reduce and dev_handle as in cv_run.
    Returns dict(class1, class2, auroc, auprc, grid_metrics, kernel_launches, stats)."""
    L = hlib()
    L.psmh_cv_grid_run_ex.argtypes = [vp, C.c_int, u32, u32, u32p, u32p, u32, u32p, u32, u32, C.c_int, vp, C.c_int, u32, u32, REDUCE_FN, vp, dp, dp, dp, dp,
                                      u64p, u64p, vp, C.c_char_p, C.c_size_t]
    if x is not None:
        assert x.dtype == np.float64 and x.flags.c_contiguous
        n, m = x.shape[0], x.shape[1] - 1
        xptr, isdev = x.ctypes.data, 0
    else:
        xptr, isdev = x_device_ptr, 1
    labels = np.ascontiguousarray(labels, np.uint32)
    folds = None if folds is None else np.ascontiguousarray(folds, np.uint32)
    g = np.ascontiguousarray(np.array(grid, np.uint32).reshape(-1, 6))
    c1 = np.zeros(n); c2 = np.zeros(n); met = np.zeros(2); gm = np.zeros((len(g), 2))
    kl = np.zeros(len(KERNEL_CLASSES), np.uint64); st = np.zeros(6, np.uint64)
    err = C.create_string_buffer(512)

    def _cb(user, ptr, count):
        try:
            reduce(ptr, count)
            return 0
        except Exception as e:   # never unwind through C
            print("reduce callback failed:", e)
            return 1
    cb = REDUCE_FN(_cb) if reduce is not None else REDUCE_FN()
    with _scratch_cwd:
        rc = L.psmh_cv_grid_run_ex(vp(xptr), isdev, n, m, _p(labels, u32p), _p(folds, u32p), nFolds, _p(g, u32p), len(g), seed, device,
                                   vp(stream) if stream else None, tie_mode, rank, world, cb, None, _p(c1, dp), _p(c2, dp), _p(met, dp), _p(gm, dp),
                                   _p(kl, u64p), _p(st, u64p), vp(dev_handle) if dev_handle else None, err, 512)
    if rc != 0:
        raise PsmError(rc, err.value.decode())
    return dict(class1=c1, class2=c2, auroc=float(met[0]), auprc=float(met[1]), grid_metrics=gm,
                kernel_launches={k_: int(kl[i]) for i, k_ in enumerate(KERNEL_CLASSES)},
                stats=dict(split_bytes=int(st[0]), nodes=int(st[1]), levels=int(st[2]), mixed_tie_pairs=int(st[3]), curves_host_sorts=int(st[4]), curves_max_tie_spread=float(st[5]) * 1e-15))


class HostDevice:
    """Persistent GPU context for repeated cv_run calls (device buffers are reused, like a long-lived process)."""

    def __init__(self, device=0, stream=None):
        self.h = hlib().psmh_device_create(device, vp(stream) if stream else None)
        if not self.h:
            raise PsmError(-1, "no usable CUDA device %d (parsmurf_b200 has no CPU fallback)" % device)

    def cv_run(self, *a, **kw):
        return cv_run(*a, dev_handle=self.h, **kw)

    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id (bytes): call on one rank, hand to the others (torch.distributed broadcast, MPI, a file ...)"""
        buf = C.create_string_buffer(128)
        if hlib().psmh_comm_unique_id(C.cast(buf, vp)) != 0:
            raise PsmError(-1, "NCCL is not available (libnccl.so.2 could not be loaded)")
        return buf.raw

    def comm_init(self, id128, rank, world):
        """collective: joins this GPU to the library's NCCL communicator; cv_run(..., world > 1, reduce=None) then sums through it"""
        assert len(id128) == 128
        err = C.create_string_buffer(512)
        buf = C.create_string_buffer(bytes(id128), 128)
        if hlib().psmh_device_comm_init(self.h, C.cast(buf, vp), rank, world, err, 512) != 0:
            raise PsmError(-1, err.value.decode())

    def comm_info(self):
        r = np.zeros(1, np.uint32); w = np.zeros(1, np.uint32); v = C.c_int(0)
        hlib().psmh_device_comm_info(self.h, _p(r, u32p), _p(w, u32p), C.byref(v))
        return dict(rank=int(r[0]), world=int(w[0]), nccl_version=int(v.value))

    def cv_grid_run(self, *a, **kw):
        return cv_grid_run(*a, dev_handle=self.h, **kw)

    def close(self):
        if self.h:
            hlib().psmh_device_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate_random_set(n, m, prob, seed):
    """the simulated data set of cfgEx/simulCV.json (parsmurf::generateRandomSet): x [n][m+1] with the label column, y [n]"""
    L = hlib()
    L.psmh_generate_random_set.argtypes = [u32, u32, C.c_double, u32, dp, u32p]
    L.psmh_generate_random_set.restype = None
    x = np.empty((n, m + 1)); y = np.empty(n, np.uint32)
    L.psmh_generate_random_set(n, m, prob, seed, _p(x, dp), _p(y, u32p))
    return x, y


def ttd_fold(labels, folds, nFolds, fold, fold_to_jump=-1):
    L = hlib()
    labels = np.ascontiguousarray(labels, np.uint32)
    folds = None if folds is None else np.ascontiguousarray(folds, np.uint32)
    s = np.zeros(4, np.uint32)
    rc = L.psmh_ttd_fold(_p(labels, u32p), _p(folds, u32p), len(labels), nFolds, fold, fold_to_jump, _p(s, u32p), None, None, None, None)
    if rc:
        raise PsmError(rc, "ttd_fold")
    a = [np.zeros(int(v), np.uint32) for v in s]
    L.psmh_ttd_fold(_p(labels, u32p), _p(folds, u32p), len(labels), nFolds, fold, fold_to_jump, _p(s, u32p), *[_p(v, u32p) for v in a])
    return dict(trngPos=a[0], trngNeg=a[1], testPos=a[2], testNeg=a[3])


def part_range(units, n_folds, n_parts, world, rank, fold):
    """[first, last) of `fold`'s partitions owned by `rank` (hyperSMURF::partRange)"""
    a = np.zeros(1, np.uint32); b = np.zeros(1, np.uint32)
    hlib().psmh_part_range(int(bool(units)), n_folds, n_parts, world, rank, fold, _p(a, u32p), _p(b, u32p))
    return int(a[0]), int(b[0])


def grid_owners(grid, world):
    """rank that runs the inner CV of each grid point in a multi-rank grid search (hyperSMURF::gridOwners)"""
    L = hlib()
    L.psmh_grid_owners.argtypes = [u32p, u32, u32, u32p]
    L.psmh_grid_owners.restype = None
    g = np.ascontiguousarray(np.array(grid, np.uint32).reshape(-1, 6))
    out = np.zeros(len(g), np.uint32)
    L.psmh_grid_owners(_p(g, u32p), len(g), world, _p(out, u32p))
    return out


def lane_cuts(shares, lanes):
    """pieces per fold share of a unit-sharded rank (hyperSMURF::laneCuts)"""
    L = hlib()
    L.psmh_lane_cuts.argtypes = [u32p, u32, u32, u32p]
    L.psmh_lane_cuts.restype = None
    s_ = np.ascontiguousarray(shares, np.uint32)
    out = np.zeros(len(s_), np.uint32)
    L.psmh_lane_cuts(_p(s_, u32p), len(s_), lanes, _p(out, u32p))
    return out


def rand_fill(seed, count):
    out = np.zeros(count, np.int32)
    hlib().psmh_rand_fill(seed, _p(out, i32p), count)
    return out


def config_parse(text):
    out = np.zeros(10, np.uint32); grid = np.zeros((4096, 6), np.uint32); err = C.create_string_buffer(512)
    rc = hlib().psmh_config_parse(text.encode(), _p(out, u32p), _p(grid, u32p), 4096, err, 512)
    if rc:
        raise PsmError(rc, err.value.decode())
    keys = ["nFolds", "seed", "wmode", "woptimiz", "gridSize", "ensThrd", "rfThrd", "simulate", "n", "m"]
    d = {k_: int(out[i]) for i, k_ in enumerate(keys)}
    d["grid"] = [dict(zip(["nParts", "fp", "ratio", "k", "nTrees", "mtry"], [int(v) for v in row])) for row in grid[: d["gridSize"]]]
    return d


def import_files(data_file, label_file=None, fold_file=None, threads=0):
    """parsmurf::Importer::import (text files in the reference's format, or a PSMB1 binary data file) ->
    dict(x [n][m+1] with the label column appended, y, f, nFolds, m)"""
    L = hlib()
    L.psmh_import.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint, u64p, dp, u32p, u32p, C.c_char_p, C.c_size_t]
    enc = lambda s: None if s is None else str(s).encode()
    sz = np.zeros(5, np.uint64); err = C.create_string_buffer(512)
    if L.psmh_import(enc(data_file), enc(label_file), enc(fold_file), threads, _p(sz, u64p), None, None, None, err, 512) != 0:
        raise PsmError(-2, err.value.decode())
    x = np.zeros(int(sz[0])); y = np.zeros(int(sz[1]), np.uint32); f = np.zeros(int(sz[2]), np.uint32)
    if L.psmh_import(enc(data_file), enc(label_file), enc(fold_file), threads, _p(sz, u64p), _p(x, dp), _p(y, u32p), _p(f, u32p), err, 512) != 0:
        raise PsmError(-2, err.value.decode())
    m = int(sz[4])
    return dict(x=x.reshape(len(y), m + 1), y=y, f=f, nFolds=int(sz[3]), m=m)


def save_predictions(path, class1, class2, labels=None, threads=0):
    """parsmurf::saveToFile: the reference's text output (byte-identical), or raw f64 arrays when the name ends in .bin"""
    L = hlib()
    L.psmh_save.argtypes = [dp, dp, u32, u32p, C.c_char_p, C.c_uint]
    c1 = np.ascontiguousarray(class1, np.float64); c2 = np.ascontiguousarray(class2, np.float64)
    lab = None if labels is None else np.ascontiguousarray(labels, np.uint32)
    if L.psmh_save(_p(c1, dp), _p(c2, dp), len(c1), _p(lab, u32p), str(path).encode(), threads) != 0:
        raise PsmError(-2, "cannot write %s" % path)


def export_binary(path, x_with_label, y, folds=None):
    L = hlib()
    L.psmh_export_binary.argtypes = [C.c_char_p, dp, u32, u32, u32p, u32p]
    x = np.ascontiguousarray(x_with_label, np.float64); y = np.ascontiguousarray(y, np.uint32)
    f = None if folds is None else np.ascontiguousarray(folds, np.uint32)
    if L.psmh_export_binary(str(path).encode(), _p(x, dp), x.shape[0], x.shape[1] - 1, _p(y, u32p), _p(f, u32p)) != 0:
        raise PsmError(-2, "cannot write %s" % path)
